"""End-to-end host pipeline on the GPU: .pt / .h5 bags on disk -> PinnedBagLoader -> SlideStreamer (H2D on a copy
stream overlapped with compute) -> results on the host; checked against the ORACLE (the reference's arithmetic on
the same files) and, bit for bit, against the direct per-slide forward."""
import os

import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import build_model, to_np

pytestmark = pytest.mark.gpu


def test_streamer_matches_direct_forward(tmp_path):
    from toad_b200.loader import PinnedBagLoader
    from toad_b200.pipeline import SlideStreamer
    params = O.make_params(0, "big", 18, bias_std=0.02)
    model = build_model(params, "big", 18)
    sizes = [300, 1, 2049, 777, 1500, 64, 4096]
    ids = []
    for i, n in enumerate(sizes):
        torch.save(torch.from_numpy(O.make_bag(50 + i, n)), os.path.join(str(tmp_path), "slide_%d.pt" % i))
        ids.append("slide_%d" % i)
    loader = PinnedBagLoader(str(tmp_path), ids, max_patches=max(sizes), depth=3)
    assert loader.pinned
    streamer = SlideStreamer(model, max(sizes), depth=2)
    sexes = [float(i % 2) for i in range(len(sizes))]
    results = streamer.run(((bag, sexes[i]) for i, (bag, sid) in enumerate(loader)))
    assert len(results) == len(sizes)
    assert streamer.h2d_bytes == sum(sizes) * 1024 * 4
    for i, n in enumerate(sizes):
        x = torch.from_numpy(O.make_bag(50 + i, n)).cuda()
        with torch.no_grad():
            ref = model(x, torch.tensor([sexes[i]], device="cuda"))
        for k in ("logits", "Y_prob", "site_prob"):
            assert torch.equal(results[i][k], ref[k].cpu()), (i, k)       # same kernels, same bits
        assert int(results[i]["Y_hat"]) == int(ref["Y_hat"])
        oref = O.toad_forward(O.make_bag(50 + i, n), sexes[i], params, dtype=np.float64)       # the referee
        np.testing.assert_allclose(results[i]["logits"].numpy(), oref["logits"], rtol=1e-3, atol=2e-5)
        np.testing.assert_allclose(results[i]["Y_prob"].numpy(), oref["Y_prob"], rtol=1e-3, atol=1e-6)
        assert int(results[i]["Y_hat"]) == int(np.asarray(oref["Y_hat"]).reshape(-1)[0])


class _NpzAsH5:
    """h5py.File stand-in over an .npz payload when h5py is absent (see tests/test_loader_cpu.py)."""

    def __init__(self, path, mode="r"):
        self._z = np.load(path, allow_pickle=False)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._z.close()

    def __getitem__(self, k):
        return self._z[k]


def test_h5_bags_feed_the_attention_topk_with_coords(tmp_path, monkeypatch):
    """The reference's h5 bags carry `coords` next to `features` (datasets/dataset_mtl_concat.py:375-383): loader ->
    device -> attention_only scores -> top-k patches -> their coords, against the oracle's scores and numpy indexing."""
    import sys
    import types
    from toad_b200 import ops
    from toad_b200.loader import PinnedBagLoader
    from tests.helpers import topk_sets_match
    try:
        import h5py  # noqa: F401
        have = True
    except ImportError:
        have = False
        monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=_NpzAsH5))
    params = O.make_params(0, "big", 18, bias_std=0.02)
    model = build_model(params, "big", 18)
    rng = np.random.default_rng(5)
    sizes = [3000, 517]
    data = {}
    for i, n in enumerate(sizes):
        feats = O.make_bag(90 + i, n)
        coords = (rng.integers(0, 4000, size=(n, 2)) * 256).astype(np.int64)
        path = os.path.join(str(tmp_path), "s%d.h5" % i)
        if have:
            import h5py
            with h5py.File(path, "w") as f:
                f["features"], f["coords"] = feats, coords
        else:
            with open(path, "wb") as fh:
                np.savez(fh, features=feats, coords=coords)
        data["s%d" % i] = (feats, coords)
    loader = PinnedBagLoader(str(tmp_path), list(data), max_patches=max(sizes), depth=3, use_h5=True)
    for bag, sid, coords in loader:
        feats, cref = data[sid]
        xd, cd = bag.cuda(non_blocking=True), coords.cuda(non_blocking=True)
        with torch.no_grad():
            a = model(xd, torch.tensor([0.0], device="cuda"), attention_only=True)
        oref = O.toad_forward(feats, 0.0, params, dtype=np.float64)
        np.testing.assert_allclose(to_np(a), oref["A"][0], rtol=0, atol=1e-4)
        for k in (1, 10, 100):
            vals, idx, c = ops.topk_patches(a.contiguous(), k, cd)
            assert topk_sets_match(to_np(a), oref["A"][0], k)
            assert np.array_equal(to_np(c), cref[to_np(idx)])
