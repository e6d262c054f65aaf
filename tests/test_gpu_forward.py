"""GPU parity of the TOAD forward: our module (C ABI, sm_100a kernels) vs the reference's outputs.

Golden = the unmodified reference run in fp64/fp32 (tests/golden/make_golden.py).  Tolerances are
the north_star's: logits / attention within 1e-3 relative (fp32), top-k indices exact (up to
reference-score ties).
"""
import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import build_model, case_inputs, load_golden, rel_err, to_np, topk_sets_match

pytestmark = pytest.mark.gpu

SMALL = ["toad_big_n1", "toad_big_n2", "toad_big_n255", "toad_big_n256", "toad_big_n257",
         "toad_small_n300", "toad_big_n1000_relu"]
LARGE = ["toad_big_n10000", "toad_big_n50000"]


def run_case(name, simt):
    import os
    g = load_golden(name)
    params, x, sex = case_inputs(g)
    os.environ["TOAD_B200_SIMT"] = "1" if simt else "0"
    try:
        model = build_model(params, str(g["meta_size_arg"]), int(g["meta_n_classes"]))
        xd = torch.from_numpy(x).cuda()
        sd = torch.tensor([sex], device="cuda")
        with torch.no_grad():
            out = model(xd, sd, return_features=True)
            a_only = model(xd, sd, attention_only=True)
        torch.cuda.synchronize()
    finally:
        os.environ["TOAD_B200_SIMT"] = "0"
    return g, out, a_only


def check_against_golden(g, out, a_only, n):
    # logits / probabilities / pooled features: 1e-3 relative to the fp64 reference (with a small
    # absolute floor for entries that happen to be ~0), and against the reference's own fp32 run.
    for k in ("logits", "site_logits", "Y_prob", "site_prob", "features"):
        ours = to_np(out[k])
        ref64 = g["f64_" + k]
        assert ours.shape == ref64.shape, k
        # the split-bf16 error is absolute at the ~1e-5 level for O(1) activations, so individual
        # pooled features that happen to be ~0 get an absolute floor; logits keep the tight one.
        np.testing.assert_allclose(ours, ref64, rtol=1e-3, atol=2e-5 if k == "features" else 2e-6, err_msg=k)
    A = to_np(out["A"])
    assert A.shape == (2, n)
    # attention scores: absolute 1e-4 (scores are O(0.1-1)); softmax weights 1e-3 relative
    np.testing.assert_allclose(A, g["f64_A"], rtol=0, atol=1e-4)
    P = np.exp(A.astype(np.float64) - A.max(axis=1, keepdims=True))
    P /= P.sum(axis=1, keepdims=True)
    P64 = np.exp(g["f64_A"] - g["f64_A"].max(axis=1, keepdims=True))
    P64 /= P64.sum(axis=1, keepdims=True)
    assert rel_err(P, P64) < 1e-3
    assert np.array_equal(to_np(out["Y_hat"]), g["f64_Y_hat"])
    assert np.array_equal(to_np(out["site_hat"]), g["f64_site_hat"])
    assert out["Y_hat"].dtype == torch.int64 and tuple(out["Y_hat"].shape) == (1, 1)
    assert tuple(out["logits"].shape) == (1, int(g["meta_n_classes"]))
    # attention_only returns task-0 raw scores (model_toad.py:92-94)
    np.testing.assert_array_equal(to_np(a_only), A[0])
    for t in range(2):
        for k in (1, 10, 100):
            if k <= n:
                assert topk_sets_match(A[t], g["f64_A"][t], k), (t, k)


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_forward_tensorcore_path(name):
    g, out, a_only = run_case(name, simt=False)
    check_against_golden(g, out, a_only, int(g["meta_n"]))


@pytest.mark.parametrize("name", SMALL + ["toad_big_n10000"])
def test_forward_fp32_simt_path(name):
    g, out, a_only = run_case(name, simt=True)
    check_against_golden(g, out, a_only, int(g["meta_n"]))


def test_forward_single_cta_tensorcore_variant():
    """The cta_group::1 kernels (debug flag) give the same results as the default CTA-pair kernels."""
    import os
    os.environ["TOAD_B200_CG1"] = "1"
    try:
        g, out, a_only = run_case("toad_big_n10000", simt=False)
    finally:
        os.environ["TOAD_B200_CG1"] = "0"
    check_against_golden(g, out, a_only, int(g["meta_n"]))
    _, out2, _ = run_case("toad_big_n10000", simt=False)
    np.testing.assert_allclose(to_np(out["A"]), to_np(out2["A"]), rtol=0, atol=2e-6)
    os.environ["TOAD_B200_CG2"] = "1"
    try:
        _, out3, _ = run_case("toad_big_n10000", simt=False)
    finally:
        os.environ["TOAD_B200_CG2"] = "0"
    np.testing.assert_allclose(to_np(out3["A"]), to_np(out2["A"]), rtol=0, atol=2e-6)


def test_tensorcore_matches_split_oracle_tightly():
    """Against the oracle's restatement of the SAME split-bf16 ordering the kernel agrees to fp32
    summation noise -- this separates 'kernel bug' from 'expected reordering error'."""
    g = load_golden("toad_big_n257")
    params, x, sex = case_inputs(g)
    exp = O.toad_forward_bf16x3(x, sex, params)
    _, out, _ = run_case("toad_big_n257", simt=False)
    # differences left: fp32 accumulation order (which can flip a bf16 rounding of h1/h) and the SFU
    # tanh/sigmoid of the gate epilogue (~2e-7 each) -- an order of magnitude below the split error.
    np.testing.assert_allclose(to_np(out["A"]), exp["A"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(to_np(out["logits"]), exp["logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(to_np(out["features"]), exp["features"], rtol=1e-4, atol=2e-5)


def test_permutation_invariance_and_equivariance():
    """Bag-order property (size independent): logits invariant, A equivariant under a row permutation."""
    g = load_golden("toad_big_n10000")
    params, x, sex = case_inputs(g)
    model = build_model(params, "big", 18)
    perm = np.random.default_rng(0).permutation(x.shape[0])
    sd = torch.tensor([sex], device="cuda")
    with torch.no_grad():
        o1 = model(torch.from_numpy(x).cuda(), sd)
        o2 = model(torch.from_numpy(x[perm]).cuda(), sd)
    np.testing.assert_allclose(to_np(o1["logits"]), to_np(o2["logits"]), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(to_np(o1["A"])[:, perm], to_np(o2["A"]), rtol=0, atol=1e-6)


def test_single_patch_softmax_weight_is_one():
    """N=1: softmax weight 1, so features[t,:512] == h (SURVEY.md section 4 property)."""
    g = load_golden("toad_big_n1")
    params, x, sex = case_inputs(g)
    f = O.toad_forward(x, sex, params, dtype=np.float64, return_intermediates=True)
    _, out, _ = run_case("toad_big_n1", simt=False)
    feats = to_np(out["features"])
    np.testing.assert_allclose(feats[0, :512], f["h"][0], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(feats[1, :512], f["h"][0], rtol=1e-3, atol=2e-5)
    assert feats[0, 512] == sex and feats[1, 512] == sex


def test_duplicate_patches_tie():
    """All patches identical: uniform attention, pooled vector equals the single patch embedding."""
    g = load_golden("toad_big_n256")
    params, x, sex = case_inputs(g)
    xd = np.repeat(x[:1], 300, axis=0)
    model = build_model(params, "big", 18)
    with torch.no_grad():
        o_many = model(torch.from_numpy(xd).cuda(), torch.tensor([sex], device="cuda"), return_features=True)
        o_one = model(torch.from_numpy(x[:1].copy()).cuda(), torch.tensor([sex], device="cuda"), return_features=True)
    A = to_np(o_many["A"])
    assert np.all(A == A[:, :1])
    np.testing.assert_allclose(to_np(o_many["features"]), to_np(o_one["features"]), rtol=1e-5, atol=1e-6)


def test_repeatable_bitwise():
    """Deterministic reductions: two runs give identical bits (reference sets cudnn deterministic, main:118-119)."""
    g = load_golden("toad_big_n10000")
    params, x, sex = case_inputs(g)
    model = build_model(params, "big", 18)
    xd = torch.from_numpy(x).cuda()
    sd = torch.tensor([sex], device="cuda")
    with torch.no_grad():
        o1 = {k: v.clone() for k, v in model(xd, sd, return_features=True).items()}
        o2 = model(xd, sd, return_features=True)
    for k in o1:
        assert torch.equal(o1[k], o2[k]), k


def test_input_validation():
    g = load_golden("toad_big_n2")
    params, x, sex = case_inputs(g)
    model = build_model(params, "big", 18)
    sd = torch.tensor([sex], device="cuda")
    with pytest.raises(ValueError):
        model(torch.from_numpy(x), sd)                       # CPU tensor
    with pytest.raises(ValueError):
        model(torch.from_numpy(x).cuda().double(), sd)       # wrong dtype
    with pytest.raises(ValueError):
        model(torch.from_numpy(x).cuda()[:, :512], sd)       # wrong width / non-contiguous
    with pytest.raises(ValueError):
        model(torch.empty((0, 1024), device="cuda"), sd)     # empty bag


def test_weight_plane_reuse_tracks_parameter_updates():
    """Eval loops skip the weight split when nothing changed; an in-place parameter update (optimizer
    step, load_state_dict) or a workspace regrow must invalidate the cached planes."""
    g = load_golden("toad_big_n257")
    params, x, sex = case_inputs(g)
    model = build_model(params, "big", 18)
    xd = torch.from_numpy(x).cuda()
    sd = torch.tensor([sex], device="cuda")
    with torch.no_grad():
        o1 = to_np(model(xd, sd)["logits"])
        o2 = to_np(model(xd, sd)["logits"])                  # second call reuses the planes
        assert np.array_equal(o1, o2)
        model.classifier.bias.add_(1.0)                       # head only: trunk planes could be reused, but the key changes
        o3 = to_np(model(xd, sd)["logits"])
        np.testing.assert_allclose(o3, o1 + 1.0, rtol=0, atol=1e-6)
        model.attention_net[0].weight.mul_(0.5)               # trunk weight changed in place
        o4 = to_np(model(xd, sd)["logits"])
        assert np.abs(o4 - o3).max() > 1e-3
        big = torch.randn(5000, 1024, device="cuda")          # forces the workspace to regrow
        model(big, sd)
        o5 = to_np(model(xd, sd)["logits"])
        assert np.array_equal(o4, o5)


def test_giga_slide_attention_only_and_topk():
    """Config 5 shape: one N=200,000 bag -- attention_only scores vs the fp32 oracle and the heat-map top-k
    (values exact w.r.t. our own scores, index sets equal to the oracle's up to its own ulp-level ties)."""
    from toad_b200 import ops
    n = 200000
    params = O.make_params(0, "big", 18)
    x = O.make_bag(5, n)
    model = build_model(params, "big", 18)
    xd = torch.from_numpy(x).cuda()
    with torch.no_grad():
        a0 = model(xd, torch.tensor([0.0], device="cuda"), attention_only=True)
        full = model(xd, torch.tensor([0.0], device="cuda"))
    torch.cuda.synchronize()
    assert tuple(a0.shape) == (n,)
    ref = O.toad_forward(x, 0.0, params, dtype=np.float32)      # numpy fp32 restatement (~2 s on CPU)
    np.testing.assert_allclose(to_np(a0), ref["A"][0], rtol=0, atol=1e-4)
    np.testing.assert_allclose(to_np(full["logits"]), ref["logits"], rtol=1e-3, atol=2e-5)
    for t in range(2):
        scores = full["A"][t].contiguous()
        for k in (1, 10, 100, 1000):
            vals, idx = ops.topk(scores, k)
            tv, ti = torch.topk(scores, k)
            assert torch.equal(vals, tv)
            if k <= 100:
                # vs the fp32 oracle the band must cover the split-bf16 score error (~2e-5 abs), not just ulps
                assert topk_sets_match(to_np(scores), ref["A"][t].astype(np.float64), k, ulps=2048)


def test_sex_covariate_enters_linearly():
    """sex is concatenated to the pooled vector (model_toad.py:99): logits(sex=1) - logits(sex=0) equals the
    last column of each head's weight, and the attention scores do not depend on it."""
    g = load_golden("toad_big_n257")
    params, x, _ = case_inputs(g)
    model = build_model(params, "big", 18)
    xd = torch.from_numpy(x).cuda()
    with torch.no_grad():
        o0 = model(xd, torch.tensor([0.0], device="cuda"), return_features=True)
        o1 = model(xd, torch.tensor([1.0], device="cuda"), return_features=True)
        o1i = model(xd, torch.tensor([1], device="cuda"))          # LongTensor sex, as collate_MIL_mtl_concat builds it
    assert torch.equal(o0["A"], o1["A"])
    np.testing.assert_allclose(to_np(o1["logits"] - o0["logits"])[0], params[O.PARAM_KEYS[10]][:, 512], rtol=0, atol=2e-6)
    np.testing.assert_allclose(to_np(o1["site_logits"] - o0["site_logits"])[0], params[O.PARAM_KEYS[12]][:, 512], rtol=0, atol=2e-6)
    assert to_np(o0["features"])[0, 512] == 0.0 and to_np(o1["features"])[1, 512] == 1.0
    assert torch.equal(o1i["logits"], o1["logits"])


@pytest.mark.parametrize("n", [3, 33, 4768, 4896, 37001])
def test_forward_ragged_partitions_vs_oracle(n):
    """Patch counts that leave trailing CTAs of the pooling tail -- or a whole merge group of them -- without
    rows (148 CTAs x rows_per_block > N): empty partials must carry weight 0 through both merge levels."""
    params = O.make_params(3, "big", 18, 0.02)
    x = O.make_bag(900 + n, n)
    model = build_model(params, "big", 18)
    with torch.no_grad():
        out = model(torch.from_numpy(x).cuda(), torch.tensor([1.0], device="cuda"), return_features=True)
    ref = O.toad_forward(x, 1.0, params, dtype=np.float64)
    for k in ("logits", "site_logits", "Y_prob", "site_prob"):
        np.testing.assert_allclose(to_np(out[k]), ref[k], rtol=1e-3, atol=2e-6, err_msg=k)
    np.testing.assert_allclose(to_np(out["features"]), ref["features"], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(to_np(out["A"]), ref["A"], rtol=0, atol=1e-4)
