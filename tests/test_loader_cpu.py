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
