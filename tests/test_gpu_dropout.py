"""Training-mode Dropout(0.25) (reference models/model_toad.py:27-29,60-64).

Bitwise parity with torch's Philox stream is impossible, so the contract is: (1) the mask is a pure
function of (seed, layer, element) that the oracle regenerates exactly, (2) with that mask the
forward/backward equal the reference's masked arithmetic (oracle, validated against torch autograd),
(3) the keep rate is 1-p with 1/(1-p) rescaling, (4) eval mode is unaffected."""
import os

import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import grads_close, load_golden, to_np

pytestmark = pytest.mark.gpu


def _model(params, size_arg, n_classes):
    from models.model_toad import TOAD_fc_mtl_concat
    m = TOAD_fc_mtl_concat(size_arg=size_arg, dropout=True, n_classes=n_classes)
    sd = dict(zip(m.state_dict().keys(), [torch.from_numpy(v.copy()) for v in params.values()]))  # keys shift to 0,3,6
    m.load_state_dict(sd, strict=True)
    m.relocate()
    return m


@pytest.mark.parametrize("simt", [True, False])
def test_dropout_training_step_matches_masked_oracle(simt):
    n, size_arg, nc, D = 300, "small", 2, 256
    params = O.make_params(3, size_arg, nc, 0.02)
    x = O.make_bag(7, n)
    os.environ["TOAD_B200_SIMT"] = "1" if simt else "0"
    try:
        m = _model(params, size_arg, nc)
        m.train()
        torch.manual_seed(1234)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # what the module will draw
        torch.manual_seed(1234)
        r = m(torch.from_numpy(x).cuda(), torch.tensor([1.0], device="cuda"), return_features=True)
        ce = torch.nn.CrossEntropyLoss()
        loss = 0.75 * ce(r["logits"], torch.tensor([1], device="cuda")) + 0.25 * ce(r["site_logits"], torch.tensor([0], device="cuda"))
        loss.backward()
        torch.cuda.synchronize()
    finally:
        os.environ["TOAD_B200_SIMT"] = "0"
    masks = O.dropout_multipliers(seed, 0.25, n, 512, D)
    f = O.toad_forward(x, 1.0, params, dtype=np.float64, masks=masks)
    np.testing.assert_allclose(to_np(r["logits"]), f["logits"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(to_np(r["A"]), f["A"], rtol=0, atol=2e-4)
    g = O.toad_backward(x, 1.0, params, 1, 0, masks=masks)
    gscale = max(np.abs(v).max() for v in g.values())
    for (k_ours, prm), k in zip(m.named_parameters(), O.PARAM_KEYS):
        ok, err, tol = grads_close(to_np(prm.grad), g[k], 5e-3, 1e-6 * gscale, robust=not simt)
        assert ok, (k_ours, err, tol)


def test_dropout_mask_statistics_and_rescale():
    """Saved activations: ~25% extra zeros, survivors scaled by 4/3 relative to the no-dropout run."""
    from toad_b200 import _lib, ops
    n = 2000
    params = O.make_params(0, "big", 18)
    x = torch.from_numpy(O.make_bag(1, n)).cuda()
    dims = ops.make_dims(1024, 512, 384, 18)
    plist = [torch.from_numpy(v).cuda() for v in params.values()]
    sex = torch.tensor([0.0], device="cuda")
    ws = ops.Workspace()
    base = ops.alloc_saved(dims, n, x.device, _lib.FLAG_SIMT_FP32)
    ops.toad_fwd(dims, plist, x, sex, ws, _lib.FLAG_SIMT_FP32 | _lib.FLAG_SAVE_ACTS, base)
    drop = ops.alloc_saved(dims, n, x.device, _lib.FLAG_SIMT_FP32)
    drop["dropout_seed"], drop["dropout_p"] = 99, 0.25
    ops.toad_fwd(dims, plist, x, sex, ws, _lib.FLAG_SIMT_FP32 | _lib.FLAG_SAVE_ACTS | _lib.FLAG_DROPOUT, drop)
    torch.cuda.synchronize()
    a0, a1 = to_np(base["a"]), to_np(drop["a"])          # h feeding `a` differs (its own dropout) -> compare h1
    h0, h1 = to_np(base["h1"]), to_np(drop["h1"])
    kept = h1 != 0
    pos = h0 > 0
    assert abs(kept[pos].mean() - 0.75) < 0.01
    np.testing.assert_allclose(h1[kept & pos], h0[kept & pos] * (4.0 / 3.0), rtol=1e-6)
    assert np.all(h1[~pos] == 0)
    assert abs((a1 == 0).mean() - 0.25) < 0.01 and (a0 == 0).mean() < 1e-3


def test_dropout_module_in_eval_mode_equals_plain_model():
    g = load_golden("toad_big_n257")
    params = O.make_params(int(g["meta_pseed"]), "big", 18, float(g["meta_bias_std"]))
    x = O.make_bag(int(g["meta_xseed"]), 257)
    m = _model(params, "big", 18)
    m.eval()
    with torch.no_grad():
        r = m(torch.from_numpy(x).cuda(), torch.tensor([float(g["meta_sex"])], device="cuda"))
    np.testing.assert_allclose(to_np(r["logits"]), g["f64_logits"], rtol=1e-3, atol=2e-6)


def test_dropout_masks_differ_between_calls_and_repeat_with_seed():
    params = O.make_params(3, "small", 2, 0.02)
    x = torch.from_numpy(O.make_bag(7, 300)).cuda()
    m = _model(params, "small", 2)
    m.train()
    s = torch.tensor([1.0], device="cuda")
    with torch.no_grad():
        torch.manual_seed(7)
        l1 = to_np(m(x, s)["logits"])
        l2 = to_np(m(x, s)["logits"])
        torch.manual_seed(7)
        l3 = to_np(m(x, s)["logits"])
    assert not np.array_equal(l1, l2) and np.array_equal(l1, l3)
