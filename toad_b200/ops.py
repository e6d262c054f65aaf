"""Host-side wrappers: torch tensors in, C-ABI calls on the current CUDA stream.

PyTorch is used here only for device memory and streams; all arithmetic runs in
libtoad_b200.so.  Every function raises (ValueError / ToadError) instead of
falling back to another implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Dims, FwdOut, Params, Saved

PARAM_ORDER = ("w1", "b1", "w2", "b2", "wa", "ba", "wb", "bb", "wc", "bc", "wcls", "bcls", "wsite", "bsite")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_dev_f32(t: torch.Tensor, name: str, shape: Optional[Sequence[int]] = None, align: int = 16) -> None:
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (toad_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if t.data_ptr() % align != 0:
        raise ValueError("%s must be %d-byte aligned (the kernels use 128-bit loads)" % (name, align))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))


class Workspace:
    """Grow-only 256-byte aligned device scratch buffer (caller-owned, per module / per stream)."""

    def __init__(self) -> None:
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device: torch.device) -> Tuple[int, int]:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        ptr = self.buf.data_ptr()
        aligned = (ptr + 255) // 256 * 256
        return aligned, self.buf.numel() - (aligned - ptr)


_TOPK_WS: Dict[tuple, "Workspace"] = {}


def make_dims(in_dim: int, hid_dim: int, attn_dim: int, n_classes: int, n_tasks: int = 2) -> Dims:
    return Dims(in_dim, hid_dim, attn_dim, n_tasks, n_classes)


def param_offsets(dims: Dims):
    off = (C.c_int64 * 15)()
    _lib.check(_lib.load().toad_param_offsets(C.byref(dims), off), "toad_param_offsets")
    return list(off)


def _params_struct(dims: Dims, params: Sequence[torch.Tensor]) -> Params:
    L, Hd, D, T, Cn = dims.in_dim, dims.hid_dim, dims.attn_dim, dims.n_tasks, dims.n_classes
    shapes = [(Hd, L), (Hd,), (Hd, Hd), (Hd,), (D, Hd), (D,), (D, Hd), (D,), (T, D), (T,),
              (Cn, Hd + 1), (Cn,), (2, Hd + 1), (2,)]
    p = Params()
    for name, t, shp in zip(PARAM_ORDER, params, shapes):
        _check_dev_f32(t, name, shp)
        setattr(p, name, t.data_ptr())
    return p


def alloc_fwd_out(dims: Dims, n: int, device: torch.device) -> Dict[str, torch.Tensor]:
    """Fresh result tensors of one forward: five allocations instead of nine (the small non-differentiable fp32
    results and the two int64 predictions are views of shared buffers) -- at N = 10k the host side of a forward costs
    as much as its kernels."""
    T, Hd, Cn = dims.n_tasks, dims.hid_dim, dims.n_classes
    sizes = (T * (Hd + 1), Cn, 2, 2 * T)
    padded = [(v + 3) // 4 * 4 for v in sizes]          # every view starts 16-byte aligned
    small = torch.empty(sum(padded), dtype=torch.float32, device=device)
    feats, y_prob, site_prob, stats = [c[:v] for c, v in zip(torch.split(small, padded), sizes)]
    hats = torch.empty((2, 1, 1), dtype=torch.int64, device=device)
    f32 = dict(dtype=torch.float32, device=device)
    return {
        "a_raw": torch.empty((T, n), **f32),
        "features": feats.view(T, Hd + 1),
        "logits": torch.empty((1, Cn), **f32),          # (the two differentiable outputs keep their own storage)
        "y_prob": y_prob.view(1, Cn),
        "y_hat": hats[0],
        "site_logits": torch.empty((1, 2), **f32),
        "site_prob": site_prob.view(1, 2),
        "site_hat": hats[1],
        "softmax_stats": stats.view(T, 2),
    }


def alloc_saved(dims: Dims, n: int, device: torch.device, flags: int = 0) -> Dict[str, torch.Tensor]:
    """Buffers for the activations the backward needs (toad_saved_t): fp32 a, b always; h1 / h as fp32 on the
    fp32 CUDA-core path, as the (hi, lo) bf16 planes the forward GEMMs consume on the tensor-core path."""
    f32 = dict(dtype=torch.float32, device=device)
    s = {"a": torch.empty((n, dims.attn_dim), **f32), "b": torch.empty((n, dims.attn_dim), **f32)}
    if flags & _lib.FLAG_SIMT_FP32:
        s["h1"] = torch.empty((n, dims.hid_dim), **f32)
        s["h"] = torch.empty((n, dims.hid_dim), **f32)
    else:
        for k in ("h1_hi", "h1_lo", "h_hi", "h_lo"):
            s[k] = torch.empty((n, dims.hid_dim), dtype=torch.bfloat16, device=device)
    return s


def _out_struct(out: Dict[str, torch.Tensor]) -> FwdOut:
    o = FwdOut()
    for k in ("a_raw", "features", "logits", "y_prob", "y_hat", "site_logits", "site_prob", "site_hat", "softmax_stats"):
        setattr(o, k, out[k].data_ptr() if k in out and out[k] is not None else None)
    return o


def _saved_struct(saved: Dict[str, torch.Tensor]) -> Saved:
    s = Saved()
    for k in ("h1", "h", "a", "b", "h1_hi", "h1_lo", "h_hi", "h_lo"):
        setattr(s, k, saved[k].data_ptr() if saved.get(k) is not None else None)
    s.dropout_seed = int(saved.get("dropout_seed", 0))
    s.dropout_p = float(saved.get("dropout_p", 0.0))
    return s


_FWD_WS_BYTES: Dict[tuple, int] = {}


def toad_fwd(dims: Dims, params: Sequence[torch.Tensor], x: torch.Tensor, sex: torch.Tensor, ws: Workspace,
             flags: int = 0, saved: Optional[Dict[str, torch.Tensor]] = None,
             out: Optional[Dict[str, torch.Tensor]] = None, prof: Optional[int] = None,
             pstruct: Optional[Params] = None) -> Dict[str, torch.Tensor]:
    """One TOAD forward (models/model_toad.py:90-116) through the C ABI.

    FLAG_REUSE_WEIGHT_PLANES in `flags` is honoured only if the workspace does not have to grow for this
    call (a regrown buffer has lost the planes)."""
    lib = _lib.load()
    _check_dev_f32(x, "h")
    if x.dim() != 2 or x.shape[1] != dims.in_dim:
        raise ValueError("h must be [N, %d], got %s" % (dims.in_dim, tuple(x.shape)))
    n = x.shape[0]
    if n < 1:
        raise ValueError("empty bag: softmax over zero patches is undefined")
    attn_only = bool(flags & _lib.FLAG_ATTENTION_ONLY)
    if not attn_only:
        _check_dev_f32(sex, "sex", align=4)      # read as scalars: any view of a float tensor will do
        if sex.numel() != 1:
            raise ValueError("sex must hold one value")
    if out is None:
        out = {"a_raw": torch.empty((dims.n_tasks, n), dtype=torch.float32, device=x.device)} if attn_only \
            else alloc_fwd_out(dims, n, x.device)
    # pstruct: a Params block the caller validated earlier for these very tensors (the module caches it per
    # parameter version -- 14 shape / dtype / alignment checks per forward are host time the small-N path can't spare)
    p = pstruct if pstruct is not None else _params_struct(dims, params)
    wkey = (dims.in_dim, dims.hid_dim, dims.attn_dim, dims.n_tasks, dims.n_classes, n, flags)
    need = _FWD_WS_BYTES.get(wkey)
    if need is None:
        nbytes = C.c_size_t()
        _lib.check(lib.toad_fwd_workspace_bytes(C.byref(dims), n, flags, C.byref(nbytes)), "toad_fwd_workspace_bytes")
        if len(_FWD_WS_BYTES) > 4096:
            _FWD_WS_BYTES.clear()
        need = _FWD_WS_BYTES[wkey] = nbytes.value
    if (flags & _lib.FLAG_REUSE_WEIGHT_PLANES) and (ws.buf is None or ws.buf.numel() < need
                                                    or ws.buf.device != x.device):
        flags &= ~_lib.FLAG_REUSE_WEIGHT_PLANES
    wptr, wsize = ws.get(need, x.device)
    o = _out_struct(out)
    s = _saved_struct(saved) if saved is not None else None
    args = [C.byref(dims), C.byref(p), x.data_ptr(), n, None if attn_only else sex.data_ptr(), C.byref(o),
            C.byref(s) if s is not None else None, wptr, wsize, flags, _stream()]
    if prof is None:
        _lib.check(lib.toad_fwd(*args), "toad_fwd")
    else:
        _lib.check(lib.toad_fwd_profiled(*(args + [prof])), "toad_fwd_profiled")
    return out


def toad_fwd_batch(dims: Dims, params: Sequence[torch.Tensor], x: torch.Tensor, lengths: Sequence[int], sex: torch.Tensor,
                   ws: Workspace, flags: int = 0, pstruct: Optional[Params] = None) -> Dict[str, torch.Tensor]:
    """Eval forward of up to 16 slides stored back to back in x [sum(lengths), in_dim] (toad_fwd_batch): one set of
    trunk launches for all of them.  Returns slide-major tensors: a_raw [T, n_total], features [S, T, H+1],
    logits / y_prob [S, C], y_hat / site_hat [S, 1, 1], site_logits / site_prob [S, 2], softmax_stats [S, T, 2]."""
    lib = _lib.load()
    _check_dev_f32(x, "h")
    S = len(lengths)
    if not (1 <= S <= _lib.MAX_BATCH):
        raise ValueError("a batch holds 1..%d slides, got %d" % (_lib.MAX_BATCH, S))
    if any(int(n) < 1 for n in lengths):
        raise ValueError("empty bag: softmax over zero patches is undefined")
    n_total = int(sum(int(n) for n in lengths))
    if x.dim() != 2 or x.shape[1] != dims.in_dim or x.shape[0] != n_total:
        raise ValueError("h must be [sum(lengths) = %d, %d], got %s" % (n_total, dims.in_dim, tuple(x.shape)))
    _check_dev_f32(sex, "sex", (S,), align=4)
    T, Hd, Cn = dims.n_tasks, dims.hid_dim, dims.n_classes
    f32 = dict(dtype=torch.float32, device=x.device)
    out = {"a_raw": torch.empty((T, n_total), **f32), "features": torch.empty((S, T, Hd + 1), **f32),
           "logits": torch.empty((S, Cn), **f32), "y_prob": torch.empty((S, Cn), **f32),
           "y_hat": torch.empty((S, 1, 1), dtype=torch.int64, device=x.device),
           "site_logits": torch.empty((S, 2), **f32), "site_prob": torch.empty((S, 2), **f32),
           "site_hat": torch.empty((S, 1, 1), dtype=torch.int64, device=x.device),
           "softmax_stats": torch.empty((S, T, 2), **f32)}
    p = pstruct if pstruct is not None else _params_struct(dims, params)
    offs = (C.c_int64 * (S + 1))()
    acc = 0
    for i, n in enumerate(lengths):
        offs[i] = acc
        acc += int(n)
    offs[S] = acc
    nbytes = C.c_size_t()
    _lib.check(lib.toad_fwd_batch_workspace_bytes(C.byref(dims), n_total, S, flags, C.byref(nbytes)),
               "toad_fwd_batch_workspace_bytes")
    if (flags & _lib.FLAG_REUSE_WEIGHT_PLANES) and (ws.buf is None or ws.buf.numel() < nbytes.value
                                                    or ws.buf.device != x.device):
        flags &= ~_lib.FLAG_REUSE_WEIGHT_PLANES
    wptr, wsize = ws.get(nbytes.value, x.device)
    o = _out_struct(out)
    _lib.check(lib.toad_fwd_batch(C.byref(dims), C.byref(p), x.data_ptr(), offs, S, sex.data_ptr(), C.byref(o), wptr, wsize,
                                  flags, _stream()), "toad_fwd_batch")
    return out


def toad_bwd(dims: Dims, params: Sequence[torch.Tensor], x: torch.Tensor, out: Dict[str, torch.Tensor],
             saved: Dict[str, torch.Tensor], dlogits: torch.Tensor, dsite_logits: torch.Tensor, ws: Workspace,
             grad_flat: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """Gradients of the 14 parameters as one flat fp32 buffer (toad_param_offsets order)."""
    lib = _lib.load()
    n = x.shape[0]
    total = param_offsets(dims)[14]
    if grad_flat is None:
        grad_flat = torch.empty(total, dtype=torch.float32, device=x.device)
    _check_dev_f32(grad_flat, "grad_flat", (total,))
    dl = dlogits.reshape(-1).contiguous().float()
    ds = dsite_logits.reshape(-1).contiguous().float()
    _check_dev_f32(dl, "dlogits", (dims.n_classes,))
    _check_dev_f32(ds, "dsite_logits", (2,))
    p = _params_struct(dims, params)
    nbytes = C.c_size_t()
    flags &= (_lib.FLAG_SIMT_FP32 | _lib.FLAG_BWD_TRANSPOSED)
    _lib.check(lib.toad_bwd_workspace_bytes(C.byref(dims), n, flags, C.byref(nbytes)), "toad_bwd_workspace_bytes")
    wptr, wsize = ws.get(nbytes.value, x.device)
    o = _out_struct(out)
    s = _saved_struct(saved)
    _lib.check(lib.toad_bwd(C.byref(dims), C.byref(p), x.data_ptr(), n, C.byref(o), C.byref(s), dl.data_ptr(),
                            ds.data_ptr(), grad_flat.data_ptr(), wptr, wsize, flags, _stream()), "toad_bwd")
    return grad_flat


def ce_loss_grad(logits: torch.Tensor, site_logits: torch.Tensor, label: torch.Tensor, site: torch.Tensor,
                 w_cls: float = 0.75, w_site: float = 0.25):
    """(loss3, dlogits, dsite_logits): the training loss of utils/core_utils_mtl_concat.py:213-215 and its gradient
    w.r.t. both logit vectors, one launch.  loss3 = [total, cls, site] stays on the device."""
    lib = _lib.load()
    _check_dev_f32(logits, "logits")
    _check_dev_f32(site_logits, "site_logits")
    n_classes = logits.numel()
    if site_logits.numel() != 2:
        raise ValueError("site_logits must have 2 elements")
    for t, name in ((label, "label"), (site, "site")):
        if not t.is_cuda or t.dtype != torch.int64 or t.numel() != 1:
            raise ValueError("%s must be a CUDA int64 tensor with one element" % name)
    loss3 = torch.empty(3, dtype=torch.float32, device=logits.device)
    dl = torch.empty(n_classes, dtype=torch.float32, device=logits.device)
    ds = torch.empty(2, dtype=torch.float32, device=logits.device)
    _lib.check(lib.toad_ce_loss_grad(logits.data_ptr(), site_logits.data_ptr(), n_classes, label.data_ptr(),
                                     site.data_ptr(), w_cls, w_site, loss3.data_ptr(), dl.data_ptr(), ds.data_ptr(),
                                     _stream()), "toad_ce_loss_grad")
    return loss3, dl, ds


def adam_step(dims: Dims, params: Sequence[torch.Tensor], grad_flat: torch.Tensor, exp_avg: torch.Tensor,
              exp_avg_sq: torch.Tensor, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
              weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
    """torch.optim.Adam's update (utils/utils.py:65) on the 14 parameter tensors in place, one launch."""
    lib = _lib.load()
    total = param_offsets(dims)[14]
    for t, name in ((grad_flat, "grad_flat"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _check_dev_f32(t, name, (total,))
    p = _params_struct(dims, params)
    _lib.check(lib.toad_adam_step(C.byref(dims), C.byref(p), grad_flat.data_ptr(), exp_avg.data_ptr(),
                                  exp_avg_sq.data_ptr(), int(step), lr, betas[0], betas[1], eps, weight_decay,
                                  grad_scale, _stream()), "toad_adam_step")


def _attn_saved_struct(saved: Dict[str, object]) -> "_lib.AttnSaved":
    s = _lib.AttnSaved()
    s.a, s.b = saved["a"].data_ptr(), saved["b"].data_ptr()
    s.dropout_seed = int(saved.get("dropout_seed", 0))
    s.dropout_p = float(saved.get("dropout_p", 0.0))
    return s


def attn_gated_fwd(x: torch.Tensor, wa, ba, wb, bb, wc, bc, ws: Workspace, flags: int = 0,
                   saved: Optional[Dict[str, object]] = None) -> torch.Tensor:
    """Attn_Net_Gated.forward (models/model_toad.py:36-41): A [N, n_tasks].  `saved` ({"a", "b"} fp32 [N, D] buffers,
    optionally "dropout_seed" / "dropout_p" with FLAG_DROPOUT in flags): training forward, filled for attn_gated_bwd."""
    lib = _lib.load()
    _check_dev_f32(x, "x")
    if x.dim() != 2:
        raise ValueError("x must be [N, L]")
    n, L = x.shape
    D, nt = wa.shape[0], wc.shape[0]
    for t, name, shp in ((wa, "wa", (D, L)), (ba, "ba", (D,)), (wb, "wb", (D, L)), (bb, "bb", (D,)),
                         (wc, "wc", (nt, D)), (bc, "bc", (nt,))):
        _check_dev_f32(t, name, shp)
    if n < 1:
        raise ValueError("empty bag")
    A = torch.empty((n, nt), dtype=torch.float32, device=x.device)
    nbytes = C.c_size_t()
    _lib.check(lib.toad_attn_gated_workspace_bytes(L, D, nt, n, flags, C.byref(nbytes)), "toad_attn_gated_workspace_bytes")
    wptr, wsize = ws.get(nbytes.value, x.device)
    sv = None
    if saved is not None:
        _check_dev_f32(saved["a"], "saved a", (n, D))
        _check_dev_f32(saved["b"], "saved b", (n, D))
        sv = C.byref(_attn_saved_struct(saved))
    _lib.check(lib.toad_attn_gated_fwd(L, D, nt, wa.data_ptr(), ba.data_ptr(), wb.data_ptr(), bb.data_ptr(),
                                       wc.data_ptr(), bc.data_ptr(), x.data_ptr(), n, A.data_ptr(), sv, wptr, wsize,
                                       flags, _stream()), "toad_attn_gated_fwd")
    return A


def attn_gated_bwd(x: torch.Tensor, wa, wb, wc, saved: Dict[str, object], dA: torch.Tensor, ws: Workspace,
                   need_dx: bool = False):
    """Gradients of Attn_Net_Gated's six parameters (and of x when need_dx) for the upstream gradient dA [N, n_tasks]:
    (d_wa, d_ba, d_wb, d_bb, d_wc, d_bc, dx or None)."""
    lib = _lib.load()
    n, L = x.shape
    D, nt = wa.shape[0], wc.shape[0]
    dA = dA.contiguous().float()
    _check_dev_f32(dA, "dA", (n, nt), align=4)
    f32 = dict(dtype=torch.float32, device=x.device)
    g = [torch.empty((D, L), **f32), torch.empty((D,), **f32), torch.empty((D, L), **f32), torch.empty((D,), **f32),
         torch.empty((nt, D), **f32), torch.empty((nt,), **f32)]
    dx = torch.empty((n, L), **f32) if need_dx else None
    nbytes = C.c_size_t()
    _lib.check(lib.toad_attn_gated_bwd_workspace_bytes(L, D, nt, n, C.byref(nbytes)), "toad_attn_gated_bwd_workspace_bytes")
    wptr, wsize = ws.get(nbytes.value, x.device)
    sv = _attn_saved_struct(saved)
    _lib.check(lib.toad_attn_gated_bwd(L, D, nt, wa.data_ptr(), wb.data_ptr(), wc.data_ptr(), x.data_ptr(), n, C.byref(sv),
                                       dA.data_ptr(), *[t.data_ptr() for t in g], dx.data_ptr() if need_dx else None,
                                       wptr, wsize, _stream()), "toad_attn_gated_bwd")
    return tuple(g) + (dx,)


def topk(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values, indices) of the k largest scores, values descending, ties -> lower index (torch.topk contract)."""
    lib = _lib.load()
    _check_dev_f32(scores, "scores")
    if scores.dim() != 1:
        raise ValueError("scores must be 1-D")
    n = scores.shape[0]
    if not (1 <= k <= n):
        raise ValueError("k must be in [1, N]")
    vals = torch.empty(k, dtype=torch.float32, device=scores.device)
    idx = torch.empty(k, dtype=torch.int64, device=scores.device)
    nbytes = C.c_size_t()
    _lib.check(lib.toad_topk_workspace_bytes(n, k, C.byref(nbytes)), "toad_topk_workspace_bytes")
    # scratch per (device, stream): two top-k calls in flight on different streams must not share histograms
    key = (scores.device, torch.cuda.current_stream(scores.device).cuda_stream)
    ws = _TOPK_WS.get(key)
    if ws is None:
        ws = _TOPK_WS[key] = Workspace()
    wptr, wsize = ws.get(nbytes.value, scores.device)
    _lib.check(lib.toad_topk(scores.data_ptr(), n, k, vals.data_ptr(), idx.data_ptr(), wptr, wsize, _stream()), "toad_topk")
    return vals, idx


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """table[idx] for a 2-D CUDA table with 4- or 8-byte elements (e.g. patch `coords` [N, 2]) and int64 indices."""
    lib = _lib.load()
    if not (isinstance(table, torch.Tensor) and table.is_cuda and table.dim() == 2 and table.is_contiguous()):
        raise ValueError("table must be a contiguous 2-D CUDA tensor")
    if table.element_size() not in (4, 8):
        raise ValueError("table elements must be 4 or 8 bytes wide, got %s" % table.dtype)
    if not (idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 1 and idx.is_contiguous()):
        raise ValueError("idx must be a contiguous 1-D CUDA int64 tensor")
    out = torch.empty((idx.shape[0], table.shape[1]), dtype=table.dtype, device=table.device)
    if idx.shape[0] == 0:
        return out
    _lib.check(lib.toad_gather_rows(table.data_ptr(), table.shape[0], table.shape[1] * table.element_size(),
                                    idx.data_ptr(), idx.shape[0], out.data_ptr(), _stream()), "toad_gather_rows")
    return out


def topk_patches(scores: torch.Tensor, k: int, coords: Optional[torch.Tensor] = None):
    """The k highest-attention patches of one task row (results['A'][t]): (values, indices[, coords of those patches])
    -- what a heatmap / top-patch consumer of the attention scores asks for."""
    vals, idx = topk(scores, k)
    if coords is None:
        return vals, idx
    if coords.shape[0] != scores.shape[0]:
        raise ValueError("coords must have one row per patch")
    return vals, idx, gather_rows(coords, idx)


def linear_bf16x3(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], relu: bool, ws: Workspace,
                  variant: int = 0) -> torch.Tensor:
    """y = act(x w^T + b) on the tcgen05 split-bf16 GEMM (test hook for the building block)."""
    lib = _lib.load()
    _check_dev_f32(x, "x")
    _check_dev_f32(w, "w")
    m, k = x.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError("shape mismatch")
    if bias is not None:
        _check_dev_f32(bias, "bias", (n,))
    y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    nbytes = C.c_size_t()
    _lib.check(lib.toad_linear_workspace_bytes(m, n, k, C.byref(nbytes)), "toad_linear_workspace_bytes")
    wptr, wsize = ws.get(nbytes.value, x.device)
    _lib.check(lib.toad_linear_bf16x3(x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                      y.data_ptr(), m, n, k, int(relu), variant, wptr, wsize, _stream()),
               "toad_linear_bf16x3")
    return y


class Profile:
    """Per-stage CUDA-event timing of toad_fwd (bench.py roofline leg)."""
    STAGES = ("weight_split", "fc1_gemm", "fc2_gemm", "gate_gemm", "pool_tail")

    def __init__(self, max_calls: int) -> None:
        self.handle = C.c_void_p()
        _lib.check(_lib.load().toad_profile_create(C.byref(self.handle), max_calls), "toad_profile_create")

    def read(self):
        ms = (C.c_double * 5)()
        n = C.c_int32()
        _lib.check(_lib.load().toad_profile_read(self.handle, ms, C.byref(n)), "toad_profile_read")
        return dict(zip(self.STAGES, list(ms))), n.value

    def close(self) -> None:
        if self.handle:
            _lib.load().toad_profile_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
