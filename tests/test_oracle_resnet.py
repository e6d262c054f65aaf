"""Pin the ResNet oracle against reference-generated golden vectors (tests/golden/resnet_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import resnet_oracle as RO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "resnet_*.npz")))


def test_param_spec_counts():
    spec = RO.param_spec()
    n_conv = sum(1 for k, s in spec if len(s) == 4)
    n_param = sum(int(np.prod(s)) for k, s in spec if len(s) == 4 or k.endswith(".weight") or k.endswith(".bias"))
    assert n_conv == 43 and n_param == 8543296            # 43 convs, 8,543,296 params


@pytest.mark.parametrize("name", CASES)
def test_resnet_oracle_matches_reference(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    params = RO.make_params(int(z["meta_pseed"]))
    x = RO.make_images(int(z["meta_xseed"]), int(z["meta_batch"]), int(z["meta_size"]),
                       int(z["meta_width"]) if "meta_width" in z.files else None)
    if int(z["meta_size"]) <= 128:
        y64 = RO.resnet50_baseline_forward(torch.from_numpy(x).double(), params).numpy()
        np.testing.assert_allclose(y64, z["f64_out"], rtol=1e-9, atol=1e-11)
    y32 = RO.resnet50_baseline_forward(torch.from_numpy(x), params).numpy()
    np.testing.assert_allclose(y32, z["f32_out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y32, z["f64_out"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("name", [c for c in CASES if "s256" not in c and "s224" not in c])
def test_f16_activation_restatement_meets_the_parity_bar(name):
    """The CUDA trunk's default mode rounds stored activations to fp16 (include/toad_b200.h, toad_resnet_fwd flags = 0).
    Its CPU restatement stays within the 1e-3 bar of the reference's fp64 golden (relative to the feature scale) --
    with margin -- so the GPU tests can hold the kernels to that bar."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    params = RO.make_params(int(z["meta_pseed"]))
    x = RO.make_images(int(z["meta_xseed"]), int(z["meta_batch"]), int(z["meta_size"]),
                       int(z["meta_width"]) if "meta_width" in z.files else None)
    y = RO.resnet50_baseline_forward_f16act(torch.from_numpy(x), params).numpy()
    ref = z["f64_out"]
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= 6e-4 * scale, (np.abs(y - ref).max() / scale)
