"""GPU parity of resnet50_baseline (models/resnet_custom.py) against reference-generated goldens.

Tolerance: the 1e-3 relative bar of the TOAD head, relative to the feature scale (individual averaged features can
be ~0).  Both arithmetic modes are held to it: the default "f16x2" (fp16 activation planes between layers, fp16
(hi, lo) weights; restated on the CPU in oracle/resnet_oracle.py) and "bf16x3" (3-pass split-bf16, fp32-class,
held to a 4x tighter bound)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import resnet_oracle as RO
from tests.helpers import to_np

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "resnet_*.npz")))


TOL = {"f16x2": 1e-3, "bf16x3": 2.5e-4}


def build(params, precision="f16x2"):
    from models.resnet_custom import resnet50_baseline
    m = resnet50_baseline(pretrained=False)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v).copy()) for k, v in params.items()}, strict=True)
    m.precision = precision
    return m.cuda().eval()


@pytest.mark.parametrize("precision", ["f16x2", "bf16x3"])
@pytest.mark.parametrize("name", CASES)
def test_resnet_matches_reference(name, precision):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    params = RO.make_params(int(z["meta_pseed"]))
    x = RO.make_images(int(z["meta_xseed"]), int(z["meta_batch"]), int(z["meta_size"]),
                       int(z["meta_width"]) if "meta_width" in z.files else None)
    model = build(params, precision)
    with torch.no_grad():
        y = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    y = to_np(y)
    ref = z["f64_out"]
    assert y.shape == ref.shape == (int(z["meta_batch"]), 1024)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= TOL[precision] * scale, (np.abs(y - ref).max() / scale, precision)
    np.testing.assert_allclose(y, ref, rtol=1e-2, atol=1e-3 * scale)


@pytest.mark.parametrize("shape", [(4, 96, 160), (3, 256, 256), (5, 128, 128)])
def test_f16_mode_matches_its_cpu_restatement(shape):
    """The default mode against oracle.resnet50_baseline_forward_f16act (same rounding points): far tighter than the
    parity bar, so a wrong tile / halo / residual would show even where fp16 rounding dominates the reference error.
    256 x 256 and 128 x 128 run layer1 / layer2's 3x3 convolutions through the halo-reuse kernel (conv3x3_halo.cuh:
    zero-bordered input plane, tiles straddling image borders), 96 x 160 through the tap-by-tap implicit GEMM."""
    b, h, w = shape
    params = RO.make_params(1)
    x = RO.make_images(12, b, h, w)
    model = build(params, "f16x2")
    with torch.no_grad():
        y = to_np(model(torch.from_numpy(x).cuda()))
    ref = RO.resnet50_baseline_forward_f16act(torch.from_numpy(x), params).numpy()
    scale = np.abs(ref).max()
    # (fp32 summation order differs from the CPU's, which now and then flips an fp16 rounding: 1e-4 .. 2e-4 of the
    # feature scale is that noise -- measured 2.0e-4 at 128 x 128 --, a wrong tap or border shows at 1e-2 and more)
    assert np.abs(y - ref).max() <= 3e-4 * scale, np.abs(y - ref).max() / scale


def test_halo_convs_equal_the_tap_by_tap_convs():
    """conv3x3_halo.cuh and the A_CONV implicit GEMM contract the same fp16 operands (same K order per tap, same fp32
    accumulator), so the features agree to fp16-rounding noise; the workspace is reused across calls and shapes, so
    the second shape also proves the border cells are re-zeroed."""
    shapes = [(256, 256), (128, 128), (64, 256), (16, 128), (32, 256)]   # (the last two: image heights of 4 / 8 rows)
    a = _features_in_child({"TOAD_RESNET_HALO": "1"}, shapes)
    b = _features_in_child({"TOAD_RESNET_HALO": "0"}, shapes)
    for shp, ya, yb in zip(shapes, a, b):
        scale = np.abs(yb).max()
        # (features of the tiny inputs average only 8 / 32 positions: the same fp16 rounding flips weigh more --
        # measured 5.4e-4 at 16 x 128; a wrong tap, border or image boundary shows at 1e-2 and more)
        tol = 5e-4 if shp[0] * shp[1] >= 128 * 128 else 1e-3
        assert np.abs(ya - yb).max() <= tol * scale, (shp, np.abs(ya - yb).max() / scale)


def test_widest_supported_patch_two_stem_tiles_per_row():
    """W = 512: the stem's output rows are 256 pixels = two 128-pixel tiles of the fused implicit-GEMM stem, and
    layer1's 3x3 tiles span full 128-pixel rows; both modes against the oracle."""
    params = RO.make_params(1)
    x = RO.make_images(31, 2, 32, 512)
    ref = RO.resnet50_baseline_forward(torch.from_numpy(x).double(), params).numpy()
    scale = np.abs(ref).max()
    for precision in ("f16x2", "bf16x3"):
        model = build(params, precision)
        with torch.no_grad():
            y = to_np(model(torch.from_numpy(x).cuda()))
        assert np.abs(y - ref).max() <= TOL[precision] * scale, (precision, np.abs(y - ref).max() / scale)


def _features_in_child(env_extra, shapes):
    """Features of make_images(8, 3, H, W) for each (H, W), computed in a child process (the stem switches are read once
    per process)."""
    import os
    import subprocess
    import sys
    import tempfile
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r); from oracle import resnet_oracle as RO; "
            "from models.resnet_custom import resnet50_baseline; m = resnet50_baseline(); "
            "m.load_state_dict({k: torch.from_numpy(np.asarray(v).copy()) for k, v in RO.make_params(1).items()}); m = m.cuda().eval(); "
            "shapes = %r; "
            "np.savez(sys.argv[1], *[m(torch.from_numpy(RO.make_images(8, 3, h, w)).cuda()).cpu().numpy() for h, w in shapes])")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile(suffix=".npz", delete=False) as f:
        path = f.name
    r = subprocess.run([sys.executable, "-c", code % (root, list(shapes)), path], env=dict(os.environ, **env_extra),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    with np.load(path) as z:
        outs = [z["arr_%d" % i] for i in range(len(shapes))]
    os.remove(path)
    return outs


def test_fused_stem_equals_the_im2col_stem():
    """The implicit-GEMM stem (stem.cuh) and the explicit im2col + GEMM stem contract the same fp16 operands in a
    different K order (kw-fastest vs channel-fastest): their fp32 sums differ in the last bit, which now and then flips
    the fp16 rounding of a stem activation; downstream that stays at the level of the mode's own rounding noise
    (the CPU restatement test allows the same 2e-4 .. 6e-4), far from what a wrong tap / halo / weight would give."""
    shapes = [(96, 160)]
    a = _features_in_child({"TOAD_RESNET_STEM_IM2COL": "0"}, shapes)[0]
    b = _features_in_child({"TOAD_RESNET_STEM_IM2COL": "1"}, shapes)[0]
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= 5e-4 * scale, np.abs(a - b).max() / scale


def test_stem_with_fused_maxpool_is_bit_identical():
    """MaxPool2d(3, 2, 1) folded into the stem's epilogue (stem.cuh, POOL = true) takes the maximum of the same fp16 stem
    activations the separate pooling kernel reads: the features must agree bit for bit.  Shapes: whole bands, a ragged
    last band (H/4 = 20 = 8 + 8 + 4), a single half band, one-tile-wide rows (W = 256)."""
    shapes = [(96, 160), (80, 48), (16, 16), (64, 256)]
    fused = _features_in_child({"TOAD_RESNET_STEM_POOL": "1"}, shapes)
    split = _features_in_child({"TOAD_RESNET_STEM_POOL": "0"}, shapes)
    for shp, a, b in zip(shapes, fused, split):
        assert np.array_equal(a, b), (shp, np.abs(a - b).max())


def test_resnet_shape_contract():
    """Any H, W multiple of 16 up to W = 512 (the reference's AdaptiveAvgPool2d takes any size, resnet_custom.py:106);
    other sizes raise instead of computing something else."""
    from toad_b200._lib import ToadError
    params = RO.make_params(1)
    model = build(params)
    with torch.no_grad():
        assert model(torch.zeros(1, 3, 16, 16, device="cuda")).shape == (1, 1024)
        assert model(torch.zeros(2, 3, 48, 80, device="cuda")).shape == (2, 1024)
        with pytest.raises(ToadError):
            model(torch.zeros(1, 3, 100, 100, device="cuda"))


def _oracle_features(params, x):
    return RO.resnet50_baseline_forward(torch.from_numpy(x), params).numpy()


@pytest.mark.parametrize("precision", ["f16x2", "bf16x3"])
def test_resnet_batch_crossing_the_stem_chunk(precision):
    """B = 300 at 64x64: the stem + layer1 loop runs two chunks (256 + 44 images, toad_resnet_fwd) whose layer1 outputs
    land in slices of one buffer; every image is compared with the oracle (the reference's own CPU arithmetic)."""
    params = RO.make_params(1)
    x = RO.make_images(21, 300, 64)
    model = build(params, precision)
    with torch.no_grad():
        y = to_np(model(torch.from_numpy(x).cuda()))
    ref = _oracle_features(params, x)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= TOL[precision] * scale, (np.abs(y - ref).max() / scale, int(np.abs(y - ref).max(1).argmax()))


def test_resnet_config3_batch_512_at_256():
    """Config 3's batch (512 patches of 3x256x256): images drawn from both stem chunks and both ends of the batch are
    checked against the oracle run on just those images (features are batch-independent)."""
    params = RO.make_params(1)
    rng = np.random.default_rng(33)
    x = torch.from_numpy(rng.standard_normal((512, 3, 256, 256), dtype=np.float32))
    model = build(params)
    with torch.no_grad():
        y = to_np(model(x.cuda()))
    assert y.shape == (512, 1024) and np.isfinite(y).all()
    pick = [0, 1, 127, 128, 255, 256, 257, 300, 383, 384, 500, 511]
    ref = _oracle_features(params, x[pick].numpy())
    scale = np.abs(ref).max()
    assert np.abs(y[pick] - ref).max() <= 1e-3 * scale, (np.abs(y[pick] - ref).max(), scale)


def test_resnet_batch_independence_and_determinism():
    """Each image's features do not depend on its batch neighbours; repeated runs are bit-identical."""
    params = RO.make_params(1)
    x = RO.make_images(9, 5, 64)
    model = build(params)
    xd = torch.from_numpy(x).cuda()
    with torch.no_grad():
        y_all = model(xd)
        y_again = model(xd)
        y_one = model(xd[3:4].contiguous())
    assert torch.equal(y_all, y_again)
    np.testing.assert_allclose(to_np(y_all[3:4]), to_np(y_one), rtol=1e-5, atol=1e-6)


def test_resnet_refuses_train_mode_and_cpu():
    from models.resnet_custom import resnet50_baseline
    m = resnet50_baseline().cuda()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 64, 64, device="cuda"))       # constructed in train mode, like the reference
    m.eval()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 64, 64))                       # CPU tensor: no fallback
