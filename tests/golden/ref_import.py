"""Load the UNMODIFIED reference modules (mahmoodlab/TOAD under /root/reference) by FILE PATH.

`import models.model_toad` cannot be used for this: the repo's own `models/` is a regular package and always
wins over the reference's namespace directory, whatever the order of sys.path.  The reference's
`models/model_toad.py:5` does `from utils.utils import initialize_weights`, so `utils.utils` is loaded from
the reference by path as well and registered in sys.modules only while model_toad executes.

Authoring-container / test infrastructure only (nothing here exists on the GPU box).
"""
import importlib.util
import os
import sys
import types

REF = "/root/reference"


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "model_toad.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_model_toad():
    """The reference's models/model_toad.py as a module object (classes Attn_Net_Gated, TOAD_fc_mtl_concat)."""
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.utils")}
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True      # /root/reference is read-only
    try:
        pkg = types.ModuleType("utils")
        pkg.__path__ = [os.path.join(REF, "utils")]
        sys.modules["utils"] = pkg
        uu = _load("utils.utils", os.path.join(REF, "utils", "utils.py"))
        sys.modules["utils.utils"] = uu
        pkg.utils = uu
        mt = _load("ref_model_toad", os.path.join(REF, "models", "model_toad.py"))
        assert mt.__file__.startswith(REF), mt.__file__
        return mt
    finally:
        sys.dont_write_bytecode = dont
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
