"""Shared helpers for the GPU parity tests (oracle = checker only)."""
import os

import numpy as np

from oracle import toad_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_inputs(g):
    params = O.make_params(int(g["meta_pseed"]), str(g["meta_size_arg"]), int(g["meta_n_classes"]),
                           float(g["meta_bias_std"]))
    x = O.make_bag(int(g["meta_xseed"]), int(g["meta_n"]), kind=str(g["meta_kind"]))
    return params, x, float(g["meta_sex"])


def build_model(params, size_arg, n_classes, device="cuda"):
    """Our module with the given numpy parameters loaded through load_state_dict (reference key names)."""
    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    model = TOAD_fc_mtl_concat(size_arg=size_arg, n_classes=n_classes)
    missing = model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in params.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.relocate()
    model.eval()
    return model


def to_np(t):
    return t.detach().cpu().numpy()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def topk_sets_match(score_ours, score_ref, k, ulps=8):
    """indices equal after removing candidates whose reference scores are within a few ulp
    of the k-th score (any fp32 summation order may swap those)."""
    ref = np.asarray(score_ref, dtype=np.float64)
    o_idx = np.argsort(-np.asarray(score_ours, dtype=np.float64), kind="stable")[:k]
    r_sorted = np.sort(ref)[::-1]
    kth = r_sorted[k - 1]
    band = ulps * np.spacing(np.float32(abs(kth))).astype(np.float64)
    sure = set(np.nonzero(ref > kth + band)[0].tolist())      # must be in
    maybe = set(np.nonzero(ref >= kth - band)[0].tolist())    # may be in
    got = set(o_idx.tolist())
    return sure.issubset(got) and got.issubset(maybe)


def grads_close(ours, ref, rel, floor, robust):
    """Gradient comparison.  robust=False: every entry within rel*max|ref| + floor.
    robust=True (tensor-core forward): the forward's ~1e-5 activation differences can flip a ReLU mask
    whose pre-activation is ~0, which moves ONE row of a weight gradient (and one bias entry) by a whole
    |dz|*x term (a handful of rows per 1000 patches) -- so require 98% of the entries within tolerance and cap the outliers at 10% of the
    gradient scale (a wrong kernel is off everywhere by O(1))."""
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    err = np.abs(ours - ref)
    tol = rel * scale + floor
    if not robust:
        return bool(err.max() <= tol), float(err.max()), float(tol)
    frac_ok = float((err <= tol).mean())
    ok = frac_ok >= 0.98 and err.max() <= 0.1 * scale + floor
    return bool(ok), float(err.max()), float(tol)
