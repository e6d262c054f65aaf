"""GPU unit parity of the building-block kernels through the C ABI (ops.* wrappers)."""
import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import load_golden, to_np

pytestmark = pytest.mark.gpu


def _linear_case(m, n, k, relu, variant, seed=0):
    from toad_b200 import ops
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((m, k), dtype=np.float32)
    w = (rng.standard_normal((n, k), dtype=np.float32) * np.float32(1.0 / np.sqrt(k))).astype(np.float32)
    b = (rng.standard_normal(n, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
    ws = ops.Workspace()
    y = ops.linear_bf16x3(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(),
                          relu, ws, variant)
    torch.cuda.synchronize()
    xh, xl = O.split_bf16(x)
    wh, wl = O.split_bf16(w)
    exp_split = O.linear_bf16x3(xh, xl, wh, wl) + b.astype(np.float64)
    exp_true = x.astype(np.float64) @ w.astype(np.float64).T + b.astype(np.float64)
    if relu:
        exp_split = np.maximum(exp_split, 0)
        exp_true = np.maximum(exp_true, 0)
    return to_np(y), exp_split, exp_true


# variant = a_split | (cta_pair << 1) | (tile << 4): tile 1 -> 64, 2 -> 128, 3 -> 256 columns
@pytest.mark.parametrize("m,n,k,variant", [
    (128, 64, 64, 0x10), (128, 64, 64, 0x11),        # one tile, one K block, both A feeds
    (128, 128, 128, 0x20), (128, 256, 128, 0x30),
    (200, 256, 256, 0x31), (1, 64, 64, 0x10), (129, 128, 192, 0x21),
    (1000, 512, 1024, 0x30), (1000, 512, 512, 0x31), (777, 768, 512, 0x30),
    (5000, 256, 64, 0x20), (40000, 512, 1024, 0x00),  # multi-wave persistent schedule
    # CTA pairs (cta_group::2): 256-row tiles, each CTA stages half of the weight tile
    (256, 64, 64, 0x12), (256, 64, 64, 0x13), (256, 128, 128, 0x22), (256, 256, 256, 0x33),
    (1, 256, 64, 0x32), (129, 256, 192, 0x33), (300, 512, 512, 0x32), (1000, 768, 512, 0x33),
    (40000, 512, 1024, 0x02), (50000, 512, 512, 0x03),
    # 512-wide pair tiles (two UMMA halves share one staged A tile; single TMEM accumulator)
    (256, 512, 64, 0x42), (300, 512, 128, 0x43), (1, 512, 64, 0x42), (1000, 1024, 256, 0x43),
    (50000, 512, 1024, 0x42), (40000, 512, 512, 0x43),
])
def test_linear_bf16x3(m, n, k, variant):
    y, exp_split, exp_true = _linear_case(m, n, k, relu=(m % 2 == 0), variant=variant)
    # vs the same split ordering: fp32 accumulation noise only
    np.testing.assert_allclose(y, exp_split, rtol=2e-5, atol=2e-5 * np.sqrt(k / 64.0))
    # vs exact fp64 linear: the 3-pass split keeps fp32-class accuracy
    np.testing.assert_allclose(y, exp_true, rtol=1e-4, atol=1e-4)


def test_attn_net_gated_default_config1():
    """Config 1: Attn_Net_Gated() defaults (L=1024, D=256, n_tasks=1) on a 256x1024 bag."""
    from models.model_toad import Attn_Net_Gated
    g = load_golden("attn_gated_default_n256")
    p = O.make_attn_params(int(g["meta_seed"]), 1024, 256, 1)
    x = O.make_bag(int(g["meta_seed"]) + 1, 256, width=1024)
    net = Attn_Net_Gated()
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=True)
    net = net.cuda().eval()
    xd = torch.from_numpy(x).cuda()
    with torch.no_grad():
        A, xx = net(xd)
    assert xx is xd                                    # passthrough is the same object (model_toad.py:41)
    assert tuple(A.shape) == (256, 1)
    np.testing.assert_allclose(to_np(A), g["f64_A"], rtol=1e-3, atol=1e-5)
    A2, _ = net(xd)                                    # grad-enabled standalone use: same values, with autograd history
    assert A2.requires_grad and torch.allclose(A2.detach(), A, rtol=0, atol=1e-6)   # (tests/test_gpu_attn_gated_train.py)


@pytest.mark.parametrize("n,k", [(5, 1), (5, 5), (1000, 10), (50000, 100), (200000, 1000), (4096, 2048)])
def test_topk_matches_torch(n, k):
    from toad_b200 import ops
    rng = np.random.default_rng(n + k)
    s = rng.standard_normal(n).astype(np.float32)
    if n >= 1000:
        s[rng.integers(0, n, 50)] = s[rng.integers(0, n, 50)]   # inject exact ties
    sd = torch.from_numpy(s).cuda()
    vals, idx = ops.topk(sd, k)
    ev, ei = O.topk_indices(s, k)
    np.testing.assert_array_equal(to_np(vals), ev)
    np.testing.assert_array_equal(to_np(idx), ei)
    tv, _ = torch.topk(sd, k)
    np.testing.assert_array_equal(to_np(vals), to_np(tv))


def test_topk_all_equal_and_negative():
    from toad_b200 import ops
    s = torch.full((1000,), -2.5, device="cuda")
    vals, idx = ops.topk(s, 7)
    assert to_np(idx).tolist() == list(range(7)) and np.all(to_np(vals) == -2.5)
    s2 = torch.tensor([-1.0, -3.0, 0.0, -0.5, 2.0, -7.0], device="cuda")
    vals, idx = ops.topk(s2, 6)
    assert to_np(idx).tolist() == [4, 2, 3, 0, 1, 5]


def test_heatmap_topk_on_giga_slide_scores():
    """Config 5 shape: top-k over N=200000 scores (values exact, indices exact incl. ties)."""
    from toad_b200 import ops
    s = O.make_bag(5, 200000, width=1)[:, 0].copy()
    sd = torch.from_numpy(s).cuda()
    for k in (1, 10, 100, 1000):
        vals, idx = ops.topk(sd, k)
        ev, ei = O.topk_indices(s, k)
        np.testing.assert_array_equal(to_np(idx), ei)
        np.testing.assert_array_equal(to_np(vals), ev)


@pytest.mark.parametrize("n,k,levels", [(300000, 2048, 7), (123457, 1000, 3), (2049, 2048, 2), (1 << 20, 100, 50)])
def test_topk_heavy_ties_across_ctas(n, k, levels):
    """Quantised scores: the k-th value is shared by thousands of patches spread over every CTA's chunk, so the
    index-ordered tie rule (lowest indices win, like torch.topk / a stable sort) is exercised across CTA boundaries."""
    from toad_b200 import ops
    rng = np.random.default_rng(n)
    s = rng.integers(0, levels, n).astype(np.float32) - 1.5
    sd = torch.from_numpy(s).cuda()
    vals, idx = ops.topk(sd, k)
    ev, ei = O.topk_indices(s, k)
    np.testing.assert_array_equal(to_np(vals), ev)
    np.testing.assert_array_equal(to_np(idx), ei)
    # a second call reuses the (self-zeroed) workspace
    vals2, idx2 = ops.topk(sd, k)
    assert torch.equal(vals, vals2) and torch.equal(idx, idx2)
