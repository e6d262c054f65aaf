"""Standalone Attn_Net_Gated (reference models/model_toad.py:17-41) with autograd and training-mode dropout: the
one place round 1's module surface refused a call the reference accepts.  Gradients against torch fp64 autograd of
the reference's formula; dropout against the masked formula with the masks regenerated from toad_dropout_hash."""
import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import to_np

pytestmark = pytest.mark.gpu


def _module(p, L, D, nt, dropout):
    from models.model_toad import Attn_Net_Gated
    m = Attn_Net_Gated(L=L, D=D, dropout=dropout, n_tasks=nt)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=True)
    return m.cuda()


def _ref_grads(p, x, G, mask_a=None, mask_b=None):
    """fp64 torch autograd of A = (tanh(x Wa^T + ba) (*mask_a) * sigmoid(x Wb^T + bb) (*mask_b)) Wc^T + bc under sum(A * G)."""
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    a = torch.tanh(xt @ t["attention_a.0.weight"].T + t["attention_a.0.bias"])
    b = torch.sigmoid(xt @ t["attention_b.0.weight"].T + t["attention_b.0.bias"])
    if mask_a is not None:
        a, b = a * torch.tensor(mask_a), b * torch.tensor(mask_b)
    A = (a * b) @ t["attention_c.weight"].T + t["attention_c.bias"]
    (A * torch.tensor(G, dtype=torch.float64)).sum().backward()
    return A.detach().numpy(), {k: v.grad.numpy() for k, v in t.items()}, xt.grad.numpy()


@pytest.mark.parametrize("L,D,nt,n", [(1024, 256, 1, 300), (512, 384, 2, 1000), (1024, 256, 2, 2500)])
def test_standalone_block_gradients_match_torch_autograd(L, D, nt, n):
    p = O.make_attn_params(21, L, D, nt)
    x = O.make_bag(22, n, width=L)
    G = np.random.default_rng(5).standard_normal((n, nt)).astype(np.float32)
    m = _module(p, L, D, nt, dropout=False)
    m.train()
    xd = torch.from_numpy(x).cuda().requires_grad_(True)
    A, xx = m(xd)
    assert xx is xd and A.shape == (n, nt)
    (A * torch.from_numpy(G).cuda()).sum().backward()
    torch.cuda.synchronize()
    A_ref, g_ref, dx_ref = _ref_grads(p, x, G)
    np.testing.assert_allclose(to_np(A), A_ref, rtol=0, atol=1e-4)
    for k, prm in m.named_parameters():
        ref = g_ref[k]
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(to_np(prm.grad) - ref).max() <= 2e-3 * scale, (k, np.abs(to_np(prm.grad) - ref).max() / scale)
    assert np.abs(to_np(xd.grad) - dx_ref).max() <= 2e-3 * np.abs(dx_ref).max()


def test_standalone_block_parameters_only_and_no_grad_paths():
    """x without requires_grad: parameter gradients only (dx is not computed); under no_grad: the inference kernel."""
    p = O.make_attn_params(3, 1024, 256, 1)
    x = torch.from_numpy(O.make_bag(4, 257)).cuda()
    m = _module(p, 1024, 256, 1, dropout=False)
    A, _ = m(x)
    A.sum().backward()
    assert all(q.grad is not None for q in m.parameters()) and x.grad is None
    with torch.no_grad():
        A2, _ = m(x)
    assert torch.equal(A.detach(), A2)


def test_standalone_block_training_dropout_matches_masked_formula():
    L, D, nt, n = 512, 384, 2, 700
    p = O.make_attn_params(8, L, D, nt)
    x = O.make_bag(9, n, width=L)
    G = np.random.default_rng(6).standard_normal((n, nt)).astype(np.float32)
    m = _module(p, L, D, nt, dropout=True)
    m.train()
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())        # what the module will draw
    torch.manual_seed(77)
    A, _ = m(torch.from_numpy(x).cuda())
    (A * torch.from_numpy(G).cuda()).sum().backward()
    torch.cuda.synchronize()
    thresh = np.uint32(int(0.25 * 4294967296.0))
    idx = np.arange(n * D, dtype=np.uint64)
    mask_a = ((O.dropout_hash(seed, 3, idx) >= thresh).astype(np.float64) / 0.75).reshape(n, D)
    mask_b = ((O.dropout_hash(seed, 4, idx) >= thresh).astype(np.float64) / 0.75).reshape(n, D)
    assert abs((mask_a == 0).mean() - 0.25) < 0.01
    A_ref, g_ref, _ = _ref_grads(p, x, G, mask_a, mask_b)
    np.testing.assert_allclose(to_np(A), A_ref, rtol=0, atol=2e-4)
    for k, prm in m.named_parameters():
        ref = g_ref[k]
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(to_np(prm.grad) - ref).max() <= 3e-3 * scale, (k, np.abs(to_np(prm.grad) - ref).max() / scale)
    m.eval()                                                   # eval mode: no dropout, the inference kernel
    with torch.no_grad():
        A_eval, _ = m(torch.from_numpy(x).cuda())
    A_plain, _, _ = O.attn_net_gated_forward(x.astype(np.float64), *[p[k].astype(np.float64) for k in (
        "attention_a.0.weight", "attention_a.0.bias", "attention_b.0.weight", "attention_b.0.bias",
        "attention_c.weight", "attention_c.bias")])
    np.testing.assert_allclose(to_np(A_eval), A_plain, rtol=0, atol=1e-4)
