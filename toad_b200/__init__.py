"""toad_b200 -- B200 (sm_100a) implementation of the attention-MIL hot path of mahmoodlab/TOAD.

Layout:
  csrc/            hand-written CUDA (tcgen05 split-bf16 GEMM, pooling tail, backward, top-k) + C ABI
  _lib.py          ctypes binding of libtoad_b200.so (include/toad_b200.h)
  ops.py           tensor-level wrappers over the C ABI
  model_toad.py    nn.Module mirror of the reference's models/model_toad.py
  distributed.py   one-slide-per-GPU sharding and the flat-gradient all-reduce
"""
from .model_toad import Attn_Net_Gated, TOAD_fc_mtl_concat, initialize_weights  # noqa: F401

__all__ = ["Attn_Net_Gated", "TOAD_fc_mtl_concat", "initialize_weights"]
