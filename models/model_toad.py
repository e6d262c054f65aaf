"""Same import path as the reference's models/model_toad.py; implementation in toad_b200."""
from toad_b200.model_toad import Attn_Net_Gated, TOAD_fc_mtl_concat, initialize_weights  # noqa: F401
