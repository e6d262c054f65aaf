"""Generate golden vectors from the UNMODIFIED reference (mahmoodlab/TOAD).

Run in the authoring container only (needs /root/reference, CPU torch):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [case names ...]

The reference module (`/root/reference/models/model_toad.py`) is imported by
path, loaded with seeded parameters from `oracle.toad_oracle.make_params`
(numpy PCG64 -> identical on the GPU box), run in fp32 (what the reference
computes) and in fp64 (`model.double()`, the arbiter when fp32 summation
orders differ), and its outputs are written to `tests/golden/*.npz`.
Inputs are NOT stored: they are regenerated from the recorded seeds with
`oracle.toad_oracle.make_bag`.

Big gradient tensors are stored as a strided subsample plus row/column sums
(see `grad_digest`) to keep fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import toad_oracle as O  # noqa: E402

from tests.golden.ref_import import load_reference_model_toad  # noqa: E402


def import_reference():
    """The reference's model module, loaded by file path (see ref_import.py: the repo's own `models` package
    shadows the reference's under any sys.path order)."""
    return load_reference_model_toad()


def grad_digest(g: np.ndarray) -> dict:
    g = np.asarray(g, dtype=np.float64)
    if g.ndim == 1 or g.size <= 20000:
        return {"full": g}
    return {"sub": g[::37, ::41].copy(), "rowsum": g.sum(axis=1), "colsum": g.sum(axis=0)}


def run_case(mt, name, n, size_arg, n_classes, pseed, xseed, sex, bias_std=0.0,
             kind="randn", grads=False, label=3, site=1, store_A64=True, light=False, out_dir=None):
    """light=True: gradient-parity fixture for a large bag -- drops the per-patch arrays (A, attention_only) so the
    file stays small; logits / features / loss / gradient digests remain."""
    params = O.make_params(pseed, size_arg, n_classes, bias_std)
    x = O.make_bag(xseed, n, kind=kind)
    torch.manual_seed(0)
    model = mt.TOAD_fc_mtl_concat(size_arg=size_arg, n_classes=n_classes)
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    missing = model.load_state_dict(sd, strict=True)
    model.eval()
    out = {"meta_n": n, "meta_size_arg": size_arg, "meta_n_classes": n_classes, "meta_pseed": pseed,
           "meta_xseed": xseed, "meta_sex": float(sex), "meta_bias_std": bias_std, "meta_kind": kind,
           "meta_label": label, "meta_site": site}
    with torch.no_grad():
        r32 = model(torch.from_numpy(x), torch.tensor([float(sex)]), return_features=True)
        a_only = model(torch.from_numpy(x), torch.tensor([float(sex)]), attention_only=True)
    for k, v in r32.items():
        if light and k == "A":
            continue
        out["f32_" + k] = v.numpy()
    if not light:
        out["f32_attention_only"] = a_only.numpy()
    m64 = mt.TOAD_fc_mtl_concat(size_arg=size_arg, n_classes=n_classes)
    m64.load_state_dict(sd, strict=True)
    m64 = m64.double().eval()
    x64 = torch.from_numpy(x).double()
    s64 = torch.tensor([float(sex)], dtype=torch.float64)
    with torch.no_grad():
        r64 = m64(x64, s64, return_features=True)
    for k, v in r64.items():
        if k == "A" and (light or not store_A64):
            continue
        out["f64_" + k] = v.numpy()
    if grads:
        m64.train()
        r = m64(x64, s64)
        loss_fn = torch.nn.CrossEntropyLoss()
        loss = 0.75 * loss_fn(r["logits"], torch.tensor([label])) + 0.25 * loss_fn(r["site_logits"], torch.tensor([site]))
        loss.backward()
        out["f64_loss"] = np.float64(loss.item())
        for k, prm in m64.named_parameters():
            for dk, dv in grad_digest(prm.grad.numpy()).items():
                out["g64_%s__%s" % (k, dk)] = dv
    path = os.path.join(out_dir or HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    return path


def run_attn_gated(mt, name, n, L, D, n_tasks, seed):
    """Config 1: standalone Attn_Net_Gated defaults (model_toad.py:19)."""
    p = O.make_attn_params(seed, L, D, n_tasks)
    net = mt.Attn_Net_Gated(L=L, D=D, n_tasks=n_tasks)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()}, strict=True)
    x = O.make_bag(seed + 1, n, width=L)
    with torch.no_grad():
        A, xx = net(torch.from_numpy(x))
        A64, _ = net.double()(torch.from_numpy(x).double())
    out = {"meta_n": n, "meta_L": L, "meta_D": D, "meta_n_tasks": n_tasks, "meta_seed": seed,
           "f32_A": A.numpy(), "f64_A": A64.numpy(), "x_passthrough_equal": np.array(bool(torch.equal(xx, torch.from_numpy(x))))}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


# name -> keyword arguments of run_case (the committed fixtures; tests/test_oracle_golden.py regenerates one of the
# small ones from the reference and checks bit-equality with the committed file)
CASES = {}
for _n in (1, 2, 255, 256, 257):
    CASES["toad_big_n%d" % _n] = dict(n=_n, size_arg="big", n_classes=18, pseed=0, xseed=100 + _n, sex=_n % 2,
                                      bias_std=0.02 if _n in (2, 257) else 0.0, grads=(_n in (1, 257)))
CASES["toad_small_n300"] = dict(n=300, size_arg="small", n_classes=2, pseed=3, xseed=7, sex=1, bias_std=0.02, grads=True,
                                label=1, site=0)
CASES["toad_big_n1000_relu"] = dict(n=1000, size_arg="big", n_classes=18, pseed=0, xseed=9, sex=0, kind="relu", grads=True)
CASES["toad_big_n10000"] = dict(n=10000, size_arg="big", n_classes=18, pseed=0, xseed=100, sex=0)
CASES["toad_big_n50000"] = dict(n=50000, size_arg="big", n_classes=18, pseed=0, xseed=150, sex=1)
# backward parity at config-4 bag sizes (split-K slices, the 74-pair schedule and the segmented reductions behave
# differently from the N <= 1000 cases): gradient digests only
CASES["toad_big_n10000_grads"] = dict(n=10000, size_arg="big", n_classes=18, pseed=0, xseed=100, sex=0, bias_std=0.02,
                                      grads=True, label=7, site=0, light=True)
CASES["toad_big_n37123_grads"] = dict(n=37123, size_arg="big", n_classes=18, pseed=0, xseed=371, sex=1, bias_std=0.02,
                                      kind="relu", grads=True, label=11, site=1, light=True)


def main():
    mt = import_reference()
    torch.set_num_threads(os.cpu_count())
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if not only or "attn_gated_default_n256" in only:
        run_attn_gated(mt, "attn_gated_default_n256", 256, 1024, 256, 1, seed=11)
    for name, kw in CASES.items():
        if not only or name in only:
            run_case(mt, name, **kw)


if __name__ == "__main__":
    main()
