"""Property tests (hypothesis) of the oracle helpers and host logic -- CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import toad_oracle as O
from toad_b200.distributed import shard_slides


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(min_value=-(2.0 ** 100), max_value=2.0 ** 100, allow_nan=False, width=32), min_size=1, max_size=64))
def test_split_bf16_pair_is_exact_and_accurate(vals):
    v = np.array(vals, dtype=np.float32)
    hi, lo = O.split_bf16(v)
    assert np.all((hi.view(np.uint32) & 0xFFFF) == 0) and np.all((lo.view(np.uint32) & 0xFFFF) == 0)
    nz = np.abs(v) > 1e-30
    rel = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - v.astype(np.float64))[nz] / np.abs(v.astype(np.float64))[nz]
    assert rel.size == 0 or rel.max() <= 2.0 ** -16


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(1, 300), st.integers(0, 2 ** 31 - 1))
def test_topk_oracle_is_sorted_and_tie_stable(n, k, seed):
    k = min(k, n)
    rng = np.random.default_rng(seed)
    s = rng.integers(-5, 6, size=n).astype(np.float32)          # many exact ties
    vals, idx = O.topk_indices(s, k)
    assert np.all(np.diff(vals) <= 0)
    assert np.array_equal(vals, np.sort(s)[::-1][:k])
    for a, b in zip(range(k - 1), range(1, k)):
        if vals[a] == vals[b]:
            assert idx[a] < idx[b]                                # ties: lower index first
    assert len(set(idx.tolist())) == k


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(1, 100000), min_size=1, max_size=200), st.integers(1, 8))
def test_shard_slides_is_a_balanced_partition(lengths, world):
    shards = [shard_slides(lengths, r, world) for r in range(world)]
    assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(lengths)               # LPT guarantee
    assert shards == [shard_slides(lengths, r, world) for r in range(world)]   # deterministic


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2 ** 62), st.integers(1, 4), st.floats(0.05, 0.9))
def test_dropout_keep_rate(seed, layer, p):
    idx = np.arange(20000, dtype=np.uint64)
    thresh = np.uint32(min(int(p * 4294967296.0), 0xFFFFFFFF))
    keep = (O.dropout_hash(seed, layer, idx) >= thresh).mean()
    assert abs(keep - (1.0 - p)) < 0.02


def test_permutation_invariance_of_the_oracle_forward():
    """Bag order: logits invariant, A equivariant (the property the GPU tests check on the kernels)."""
    params = O.make_params(0, "small", 2, 0.02)
    x = O.make_bag(3, 200)
    perm = np.random.default_rng(0).permutation(200)
    a = O.toad_forward(x, 1.0, params, dtype=np.float64)
    b = O.toad_forward(x[perm], 1.0, params, dtype=np.float64)
    np.testing.assert_allclose(a["logits"], b["logits"], rtol=1e-12)
    np.testing.assert_allclose(a["A"][:, perm], b["A"], rtol=0, atol=1e-14)


def test_topk_accuracy_matches_the_reference_formula():
    """toad_b200.eval.topk_accuracy == accuracy() of utils/eval_utils_mtl_concat.py:49-63 (restated with torch.topk)."""
    import numpy as np
    import torch
    from toad_b200.eval import topk_accuracy
    g = torch.Generator().manual_seed(0)
    probs = torch.softmax(torch.randn(200, 18, generator=g), dim=1)
    labels = torch.randint(0, 18, (200,), generator=g)
    topk = (1, 3, 5)
    _, pred = probs.topk(max(topk), 1, True, True)
    correct = pred.t().eq(labels.view(1, -1).expand_as(pred.t()))
    ref = [float(correct[:k].reshape(-1).float().sum(0) / 200) for k in topk]
    ours = topk_accuracy(probs.numpy(), labels.numpy(), topk)
    np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-6)   # the reference accumulates in fp32


def test_topk_accuracy_properties():
    """Monotone in k, 1.0 at k = n_classes, and k = 1 equals plain argmax accuracy."""
    import numpy as np
    from toad_b200.eval import topk_accuracy
    rng = np.random.default_rng(3)
    probs = rng.random((300, 18))
    labels = rng.integers(0, 18, 300)
    acc = topk_accuracy(probs, labels, tuple(range(1, 19)))
    assert all(a <= b + 1e-15 for a, b in zip(acc, acc[1:])) and acc[-1] == 1.0
    assert acc[0] == float((probs.argmax(1) == labels).mean())


def test_forward_result_buffers_layout():
    """ops.alloc_fwd_out: reference shapes / dtypes, every tensor contiguous and 16-byte aligned (the C ABI writes them
    in place), the differentiable outputs in storage of their own."""
    import torch
    from toad_b200 import ops
    d = ops.make_dims(1024, 512, 384, 18)
    out = ops.alloc_fwd_out(d, 777, torch.device("cpu"))
    shapes = {"a_raw": (2, 777), "features": (2, 513), "logits": (1, 18), "y_prob": (1, 18), "y_hat": (1, 1),
              "site_logits": (1, 2), "site_prob": (1, 2), "site_hat": (1, 1), "softmax_stats": (2, 2)}
    assert set(out) == set(shapes)
    for k, shp in shapes.items():
        t = out[k]
        assert tuple(t.shape) == shp and t.is_contiguous(), k
        assert t.dtype == (torch.int64 if k.endswith("hat") else torch.float32), k
        assert t.data_ptr() % (8 if k.endswith("hat") else 16) == 0, k
    ptr = lambda t: t.untyped_storage().data_ptr()
    others = [ptr(v) for k, v in out.items() if k not in ("logits", "site_logits")]
    assert ptr(out["logits"]) not in others and ptr(out["site_logits"]) not in others
    assert ptr(out["logits"]) != ptr(out["site_logits"])
