"""PinnedBagLoader on CPU (unpinned buffers): order, contents, ring reuse contract, error surfacing."""
import os

import pytest
import torch

from toad_b200.loader import PinnedBagLoader


def _write(tmp, name, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n, 1024, generator=g)
    torch.save(t, os.path.join(tmp, name + ".pt"))
    return t


def test_loader_yields_files_in_order(tmp_path):
    bags = {("s%d" % i): _write(str(tmp_path), "s%d" % i, 10 + 7 * i, i) for i in range(7)}
    loader = PinnedBagLoader(str(tmp_path), list(bags), max_patches=64, depth=3, pin=False)
    assert len(loader) == 7
    seen = []
    prev = []
    for view, sid in loader:
        assert tuple(view.shape) == tuple(bags[sid].shape)
        assert torch.equal(view, bags[sid])
        prev.append((view, sid))
        if len(prev) >= 2:                      # the previous item must still be intact (depth 3 ring)
            pv, ps = prev[-2]
            assert torch.equal(pv, bags[ps])
        seen.append(sid)
    assert seen == list(bags)


def test_loader_surfaces_bad_files(tmp_path):
    torch.save(torch.zeros(5, 512), os.path.join(str(tmp_path), "bad.pt"))
    with pytest.raises(ValueError):
        list(PinnedBagLoader(str(tmp_path), ["bad"], max_patches=16, pin=False))
    _write(str(tmp_path), "big", 40, 0)
    with pytest.raises(ValueError):
        list(PinnedBagLoader(str(tmp_path), ["big"], max_patches=16, pin=False))


class _FakeH5File:
    """Stand-in for h5py.File over an .npz payload (h5py is not installed in this image -- the reference has the
    same optional dependency): supports the `with h5py.File(path, 'r') as f: f['features'][:]` protocol the loader and
    the reference (datasets/dataset_mtl_concat.py:377-379) use."""

    def __init__(self, path, mode="r"):
        import numpy as np
        self._z = np.load(path, allow_pickle=False)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self._z.close()

    def __getitem__(self, k):
        return self._z[k]


def test_loader_reads_h5_features_and_coords(tmp_path, monkeypatch):
    import sys
    import types

    import numpy as np
    try:
        import h5py  # noqa: F401
        real = True
    except ImportError:
        real = False
        monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=_FakeH5File))
    rng = np.random.default_rng(0)
    want = {}
    for i in range(4):
        n = 9 + 5 * i
        feats = rng.standard_normal((n, 1024), dtype=np.float32)
        coords = rng.integers(0, 100000, size=(n, 2)).astype(np.int64)        # CLAM writes int64 coords
        path = os.path.join(str(tmp_path), "h%d.h5" % i)
        if real:
            import h5py
            with h5py.File(path, "w") as f:
                f["features"], f["coords"] = feats, coords
        else:
            with open(path, "wb") as fh:
                np.savez(fh, features=feats, coords=coords)
        want["h%d" % i] = (feats, coords)
    loader = PinnedBagLoader(str(tmp_path), list(want), max_patches=32, depth=3, pin=False, use_h5=True)
    seen = []
    for bag, sid, coords in loader:
        assert torch.equal(bag, torch.from_numpy(want[sid][0]))
        assert coords.dtype == torch.int32 and tuple(coords.shape) == (bag.shape[0], 2)
        assert np.array_equal(coords.numpy(), want[sid][1])
        seen.append(sid)
    assert seen == list(want)


def test_loader_h5_without_h5py_says_so(tmp_path, monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "h5py", None)          # import h5py -> ImportError
    open(os.path.join(str(tmp_path), "x.h5"), "wb").close()
    with pytest.raises(ImportError, match="h5py"):
        list(PinnedBagLoader(str(tmp_path), ["x"], max_patches=8, pin=False, use_h5=True))
