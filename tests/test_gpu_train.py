"""GPU parity of the fused loss / optimizer kernels and of FusedTrainStep against the reference loop
(utils/core_utils_mtl_concat.py:198-234: model -> 0.75*CE + 0.25*CE -> backward -> Adam step)."""
import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import build_model, case_inputs, load_golden, to_np

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,y,s", [(18, 3, 1), (2, 1, 0), (1000, 999, 1), (1, 0, 0)])
def test_ce_loss_grad_kernel_matches_oracle(C, y, s):
    from toad_b200 import ops
    rng = np.random.default_rng(C)
    z = (rng.standard_normal(C) * 4).astype(np.float32)
    zs = rng.standard_normal(2).astype(np.float32)
    loss3, dl, ds = ops.ce_loss_grad(torch.from_numpy(z).cuda(), torch.from_numpy(zs).cuda(),
                                     torch.tensor([y], device="cuda"), torch.tensor([s], device="cuda"))
    e3, edl, eds = O.ce_loss_grad(z, zs, y, s)
    np.testing.assert_allclose(to_np(loss3), e3, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(to_np(dl), edl, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(to_np(ds), eds, rtol=1e-5, atol=1e-7)


def test_ce_loss_grad_out_of_range_target_is_nan():
    from toad_b200 import ops
    loss3, _, _ = ops.ce_loss_grad(torch.zeros(5, device="cuda"), torch.zeros(2, device="cuda"),
                                   torch.tensor([5], device="cuda"), torch.tensor([0], device="cuda"))
    l = to_np(loss3)
    assert np.isnan(l[0]) and np.isnan(l[1]) and np.isfinite(l[2])


def test_adam_step_kernel_matches_torch_adam_over_steps():
    """Same gradients into torch.optim.Adam (fp32, CUDA, as the reference runs it) and toad_adam_step."""
    from toad_b200 import ops
    params = O.make_params(2, "big", 18, 0.02)
    dims = ops.make_dims(1024, 512, 384, 18)
    ours = [torch.from_numpy(v.copy()).cuda() for v in params.values()]
    theirs = [torch.nn.Parameter(torch.from_numpy(v.copy()).cuda()) for v in params.values()]
    opt = torch.optim.Adam(theirs, lr=1e-3, weight_decay=1e-2)
    off = ops.param_offsets(dims)
    m = torch.zeros(off[14], device="cuda")
    v = torch.zeros(off[14], device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(0)
    for step in range(1, 5):
        g = torch.randn(off[14], device="cuda", generator=gen) * (10.0 ** (step - 3))
        for i, p in enumerate(theirs):
            p.grad = g[off[i]:off[i + 1]].view_as(p).clone()
        opt.step()
        ops.adam_step(dims, ours, g, m, v, step, 1e-3, (0.9, 0.999), 1e-8, 1e-2)
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(ours, theirs)):
            # one update moves a weight by <= ~lr; agree to 1e-3 of that plus fp32 rounding of the value itself
            np.testing.assert_allclose(to_np(a), to_np(b.detach()), rtol=2e-7, atol=1e-6, err_msg="tensor %d step %d" % (i, step))
        st = opt.state[theirs[0]]
        np.testing.assert_allclose(to_np(m[off[0]:off[1]]), to_np(st["exp_avg"]).ravel(), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(to_np(v[off[0]:off[1]]), to_np(st["exp_avg_sq"]).ravel(), rtol=1e-5, atol=1e-12)


def test_adam_grad_scale_is_the_allreduce_average():
    from toad_b200 import ops
    params = O.make_params(3, "small", 2, 0.02)
    dims = ops.make_dims(1024, 512, 256, 2)
    a = [torch.from_numpy(v.copy()).cuda() for v in params.values()]
    b = [torch.from_numpy(v.copy()).cuda() for v in params.values()]
    tot = ops.param_offsets(dims)[14]
    g = torch.randn(tot, device="cuda")
    ma, va, mb, vb = (torch.zeros(tot, device="cuda") for _ in range(4))
    ops.adam_step(dims, a, g * 4.0, ma, va, 1, 1e-3, grad_scale=0.25)
    ops.adam_step(dims, b, g, mb, vb, 1, 1e-3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("name,dropout", [("toad_big_n257", False), ("toad_small_n300", False), ("toad_big_n1000_relu", True)])
def test_fused_train_step_tracks_the_eager_reference_loop(name, dropout):
    """Three optimizer steps: FusedTrainStep vs module + nn.CrossEntropyLoss + torch.optim.Adam (same kernels for
    forward/backward, so the comparison isolates the fused loss and optimizer)."""
    from toad_b200.train import FusedTrainStep
    g = load_golden(name)
    params, x, sex = case_inputs(g)
    size, ncls = str(g["meta_size_arg"]), int(g["meta_n_classes"])
    xd = torch.from_numpy(x).cuda()
    sd = torch.tensor([sex], device="cuda")
    lab = torch.tensor([int(g["meta_label"])], device="cuda")
    site = torch.tensor([int(g["meta_site"])], device="cuda")
    lr, wd = 1e-3, 1e-5

    def mk():
        if not dropout:
            return build_model(params, size, ncls)
        from models.model_toad import TOAD_fc_mtl_concat
        m = TOAD_fc_mtl_concat(size_arg=size, dropout=True, n_classes=ncls)
        m.relocate()
        with torch.no_grad():
            for p, v in zip(m._param_list(), params.values()):
                p.copy_(torch.from_numpy(v))
        return m

    eager = mk()
    eager.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, eager.parameters()), lr=lr, weight_decay=wd)
    ce = torch.nn.CrossEntropyLoss()
    losses_e = []
    torch.manual_seed(11)
    for _ in range(3):
        r = eager(xd, sd)
        cls_l, site_l = ce(r["logits"], lab), ce(r["site_logits"], site)
        loss = cls_l * 0.75 + site_l * 0.25
        losses_e.append([loss.item(), cls_l.item(), site_l.item()])
        loss.backward()
        opt.step()
        opt.zero_grad()

    fused_m = mk()
    fused_m.train()
    fused = FusedTrainStep(fused_m, lr=lr, weight_decay=wd)
    losses_f = []
    torch.manual_seed(11)
    for _ in range(3):
        r = fused(xd, lab, site, sd)
        losses_f.append(to_np(r["loss"]).tolist())
    np.testing.assert_allclose(np.array(losses_f), np.array(losses_e), rtol=2e-4, atol=1e-6)
    for (k, a), b in zip(eager.named_parameters(), fused_m._param_list()):
        # 3 steps of size <= ~lr each.  Adam normalises the update (m / sqrt(v)), so where a gradient entry is ~0 the
        # last-bit differences between the two loss kernels decide its sign: a handful of entries may differ by
        # up to the full 3 * lr, everything else must agree closely.
        d = np.abs(to_np(b).astype(np.float64) - to_np(a))
        assert d.max() <= 3.5 * lr, (k, d.max())
        if k.endswith("attention_c.bias"):
            continue   # its true gradient is 0 (softmax is shift invariant): Adam turns pure rounding noise into +-lr steps
        assert (d <= 0.05 * lr).mean() >= 0.999, (k, (d <= 0.05 * lr).mean())
    # the updated parameters are what a following eval forward sees (weight-plane cache invalidated)
    fused_m.eval()
    eager.eval()
    with torch.no_grad():
        rf, re = fused_m(xd, sd), eager(xd, sd)
    np.testing.assert_allclose(to_np(rf["logits"]), to_np(re["logits"]), rtol=1e-3, atol=1e-4)
    assert fused.step_count == 3 and all(p._version > 0 for p in fused_m._param_list())


def test_fused_train_step_state_dict_roundtrip():
    from toad_b200.train import FusedTrainStep
    g = load_golden("toad_big_n257")
    params, x, sex = case_inputs(g)
    xd, sd = torch.from_numpy(x).cuda(), torch.tensor([sex], device="cuda")
    lab, site = torch.tensor([1], device="cuda"), torch.tensor([0], device="cuda")
    m1 = build_model(params, "big", 18)
    m1.train()
    f1 = FusedTrainStep(m1)
    f1(xd, lab, site, sd)
    f1(xd, lab, site, sd)
    st = f1.state_dict()
    m2 = build_model({k: to_np(p) for k, p in zip(params.keys(), m1._param_list())}, "big", 18)
    m2.train()
    f2 = FusedTrainStep(m2)
    f2.load_state_dict(st)
    f1(xd, lab, site, sd)
    f2(xd, lab, site, sd)
    for a, b in zip(m1._param_list(), m2._param_list()):
        assert torch.equal(a, b)
