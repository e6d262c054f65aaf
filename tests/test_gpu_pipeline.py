"""End-to-end host pipeline on the GPU: .pt bags on disk -> PinnedBagLoader -> SlideStreamer (H2D on a copy
stream overlapped with compute) -> results on the host; must equal the direct per-slide forward."""
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
