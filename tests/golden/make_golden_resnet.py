"""Golden vectors for resnet50_baseline from the UNMODIFIED reference (models/resnet_custom.py).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_resnet.py

`torchsummary` (an unused import at resnet_custom.py:5) is absent here, so an empty stub module is
put in sys.modules before importing the reference -- the reference file itself is untouched.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import resnet_oracle as RO  # noqa: E402

REF = "/root/reference"


def import_reference_resnet():
    stub = types.ModuleType("torchsummary")
    stub.summary = lambda *a, **k: None
    sys.modules.setdefault("torchsummary", stub)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_resnet_custom", os.path.join(REF, "models", "resnet_custom.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(mod, name, pseed, xseed, batch, size, width=None):
    """size = image height; width defaults to the height (square patches)."""
    params = RO.make_params(pseed)
    x = RO.make_images(xseed, batch, size, width)
    model = mod.resnet50_baseline(pretrained=False)
    sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in params.items()}
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        y32 = model(torch.from_numpy(x)).numpy()
        y64 = model.double()(torch.from_numpy(x).double()).numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta_pseed=pseed, meta_xseed=xseed, meta_batch=batch, meta_size=size,
                        meta_width=width or size, f32_out=y32, f64_out=y64)
    print("wrote", path, y32.shape, os.path.getsize(path) // 1024, "KiB")


def main():
    mod = import_reference_resnet()
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    cases = {"resnet_b2_s64": (1, 2, 2, 64), "resnet_b3_s128": (1, 3, 3, 128), "resnet_b2_s256": (1, 4, 2, 256),
             # sizes that are not powers of two: the usual 224 x 224 patch, and a non-square one whose late layers
             # hold several images per M tile (96 x 160 -> 6 x 10 pixels in layer3)
             "resnet_b2_s224": (1, 5, 2, 224), "resnet_b3_s96x160": (1, 6, 3, 96, 160)}
    for name, args in cases.items():
        if not only or name in only:
            run(mod, name, *args)


if __name__ == "__main__":
    main()
