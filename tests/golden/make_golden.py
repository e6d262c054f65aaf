"""Generate golden vectors from the UNMODIFIED reference (mahmoodlab/TOAD).

Run in the authoring container only (needs /root/reference, CPU torch):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

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

REF = "/root/reference"


def import_reference():
    """Import the reference's model module without letting it shadow ours."""
    saved = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")
                  or k == "utils" or k.startswith("utils.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import importlib
        mt = importlib.import_module("models.model_toad")
        assert mt.__file__.startswith(REF), mt.__file__
        return mt
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")
                  or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def grad_digest(g: np.ndarray) -> dict:
    g = np.asarray(g, dtype=np.float64)
    if g.ndim == 1 or g.size <= 20000:
        return {"full": g}
    return {"sub": g[::37, ::41].copy(), "rowsum": g.sum(axis=1), "colsum": g.sum(axis=0)}


def run_case(mt, name, n, size_arg, n_classes, pseed, xseed, sex, bias_std=0.0,
             kind="randn", grads=False, label=3, site=1, store_A64=True):
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
        out["f32_" + k] = v.numpy()
    out["f32_attention_only"] = a_only.numpy()
    m64 = mt.TOAD_fc_mtl_concat(size_arg=size_arg, n_classes=n_classes)
    m64.load_state_dict(sd, strict=True)
    m64 = m64.double().eval()
    x64 = torch.from_numpy(x).double()
    s64 = torch.tensor([float(sex)], dtype=torch.float64)
    with torch.no_grad():
        r64 = m64(x64, s64, return_features=True)
    for k, v in r64.items():
        if k == "A" and not store_A64:
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
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


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


def main():
    mt = import_reference()
    torch.set_num_threads(os.cpu_count())
    run_attn_gated(mt, "attn_gated_default_n256", 256, 1024, 256, 1, seed=11)
    for n in (1, 2, 255, 256, 257):
        run_case(mt, "toad_big_n%d" % n, n, "big", 18, pseed=0, xseed=100 + n, sex=n % 2,
                 bias_std=0.02 if n in (2, 257) else 0.0, grads=(n in (1, 257)))
    run_case(mt, "toad_small_n300", 300, "small", 2, pseed=3, xseed=7, sex=1, bias_std=0.02, grads=True,
             label=1, site=0)
    run_case(mt, "toad_big_n1000_relu", 1000, "big", 18, pseed=0, xseed=9, sex=0, kind="relu", grads=True)
    run_case(mt, "toad_big_n10000", 10000, "big", 18, pseed=0, xseed=100, sex=0)
    run_case(mt, "toad_big_n50000", 50000, "big", 18, pseed=0, xseed=150, sex=1)


if __name__ == "__main__":
    main()
