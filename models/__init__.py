"""Drop-in `models` package: put this repo's root ahead of the reference on sys.path and the
reference's own `from models.model_toad import TOAD_fc_mtl_concat`
(utils/core_utils_mtl_concat.py:8, utils/eval_utils_mtl_concat.py:6) resolves here."""
