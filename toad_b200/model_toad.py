"""Drop-in replacements for the reference's ``models/model_toad.py`` classes.

Same constructor arguments, parameter names (``state_dict`` keys), ``relocate()``
and ``forward()`` contract as mahmoodlab/TOAD ``models/model_toad.py:17-116``; the
arithmetic runs in hand-written sm_100a kernels behind the C ABI
(``include/toad_b200.h``).  The ``nn.Linear`` / ``nn.Sequential`` sub-modules exist
only to own the parameters under the reference's names -- they are never called.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
import torch.nn as nn

from . import _lib, ops


def initialize_weights(module: nn.Module) -> None:
    """xavier_normal_ weights, zero biases -- the reference's utils/utils.py:150-154."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            m.bias.data.zero_()


def _default_flags() -> int:
    # TOAD_B200_SIMT=1 selects the fp32 CUDA-core GEMMs (debug aid); default is the tcgen05 CTA-pair path.
    flags = _lib.FLAG_SIMT_FP32 if os.environ.get("TOAD_B200_SIMT", "0") == "1" else 0
    if os.environ.get("TOAD_B200_CG1", "0") == "1":   # debug aid: cta_group::1 tensor-core kernels
        flags |= _lib.FLAG_TC_SINGLE_CTA
    if os.environ.get("TOAD_B200_CG2", "0") == "1":   # debug aid: CTA pairs everywhere
        flags |= _lib.FLAG_TC_PAIR_ALL
    if os.environ.get("TOAD_B200_BWD_T", "0") == "1":  # debug aid: backward wgrads via explicit transposes
        flags |= _lib.FLAG_BWD_TRANSPOSED
    return flags


class Attn_Net_Gated(nn.Module):
    """Attention network with sigmoid gating (reference models/model_toad.py:17-41).

    forward(x) -> (A [N, n_tasks], x) with x the same tensor object.
    """

    def __init__(self, L: int = 1024, D: int = 256, dropout: bool = False, n_tasks: int = 1):
        super().__init__()
        a: List[nn.Module] = [nn.Linear(L, D), nn.Tanh()]
        b: List[nn.Module] = [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_tasks)
        self.dropout = bool(dropout)
        self._ws = ops.Workspace()

    def forward(self, x: torch.Tensor):
        wa, ba = self.attention_a[0].weight, self.attention_a[0].bias
        wb, bb = self.attention_b[0].weight, self.attention_b[0].bias
        wc, bc = self.attention_c.weight, self.attention_c.bias
        drop = self.dropout and self.training
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not need_grad and not drop:
            A = ops.attn_gated_fwd(x, wa.detach(), ba.detach(), wb.detach(), bb.detach(), wc.detach(), bc.detach(),
                                   self._ws, _default_flags())
            return A, x
        if _default_flags() & _lib.FLAG_SIMT_FP32:
            raise NotImplementedError("toad_b200: the standalone Attn_Net_Gated trains on the tensor-core path only "
                                      "(unset TOAD_B200_SIMT)")
        return _AttnGatedFunction.apply(self, x, wa, ba, wb, bb, wc, bc), x


class _AttnGatedFunction(torch.autograd.Function):
    """A = Attn_Net_Gated(x) with a hand-written backward (toad_attn_gated_bwd): training-mode dropout (counter-hash
    masks, like inside TOAD_fc_mtl_concat) and gradients for the six parameters and for x."""

    @staticmethod
    def forward(ctx, module: "Attn_Net_Gated", x, wa, ba, wb, bb, wc, bc):
        n, D = x.shape[0], wa.shape[0]
        f32 = dict(dtype=torch.float32, device=x.device)
        saved = {"a": torch.empty((n, D), **f32), "b": torch.empty((n, D), **f32)}
        flags = _default_flags()
        if module.dropout and module.training:
            flags |= _lib.FLAG_DROPOUT
            saved["dropout_seed"] = int(torch.randint(0, 2 ** 62, (1,)).item())   # torch.manual_seed makes runs repeatable
            saved["dropout_p"] = 0.25
        A = ops.attn_gated_fwd(x.detach(), wa.detach(), ba.detach(), wb.detach(), bb.detach(), wc.detach(), bc.detach(),
                               module._ws, flags, saved=saved)
        ctx.module = module
        ctx.drop = {k: v for k, v in saved.items() if k.startswith("dropout")}
        ctx.save_for_backward(x, wa, wb, wc, saved["a"], saved["b"])
        return A

    @staticmethod
    def backward(ctx, dA):
        x, wa, wb, wc, a, b = ctx.saved_tensors
        saved = dict(ctx.drop, a=a, b=b)
        ws = ctx.module.__dict__.setdefault("_ws_bwd", ops.Workspace())
        g = ops.attn_gated_bwd(x, wa.detach(), wb.detach(), wc.detach(), saved, dA, ws, need_dx=ctx.needs_input_grad[1])
        d_wa, d_ba, d_wb, d_bb, d_wc, d_bc, dx = g
        need = ctx.needs_input_grad
        ctx.module = None
        return (None, dx, d_wa if need[2] else None, d_ba if need[3] else None, d_wb if need[4] else None,
                d_bb if need[5] else None, d_wc if need[6] else None, d_bc if need[7] else None)


class _ToadFunction(torch.autograd.Function):
    """logits / site_logits with a hand-written backward (toad_bwd) for the 14 parameters."""

    @staticmethod
    def forward(ctx, module: "TOAD_fc_mtl_concat", h: torch.Tensor, sex: torch.Tensor, *params: torch.Tensor):
        dims = module._dims
        need_grad = any(ctx.needs_input_grad[3:])
        flags = _default_flags()
        saved = None
        if need_grad or module._dropout_active():
            flags |= _lib.FLAG_SAVE_ACTS
            saved = ops.alloc_saved(dims, h.shape[0], h.device, flags)
            if module._dropout_active():
                # nn.Dropout(0.25) x4 (model_toad.py:27-29,60-64): a fresh mask per forward, drawn from
                # torch's CPU generator so torch.manual_seed() makes training runs repeatable.
                flags |= _lib.FLAG_DROPOUT
                saved["dropout_seed"] = int(torch.randint(0, 2 ** 62, (1,)).item())
                saved["dropout_p"] = 0.25
        out = ops.toad_fwd(dims, params, h, sex, module._ws, flags, saved)
        ctx.module = module
        # Nothing this Function RETURNS may be kept on ctx as a plain attribute: output -> grad_fn -> ctx -> output is
        # a reference cycle Python's GC cannot see through (two differentiable outputs share the grad_fn), and it
        # would pin the whole N x 1024 bag of every training step.  The backward only reads a_raw / features /
        # softmax_stats (toad_bwd), which go through save_for_backward (autograd's cycle-free slot for outputs)
        # together with the bag and the parameters; `saved` holds workspace-like buffers that are not outputs.
        ctx.saved = saved
        ctx.n_classes = dims.n_classes
        ctx.save_for_backward(h, out["a_raw"], out["features"], out["softmax_stats"], *params)
        res = (out["logits"], out["site_logits"], out["y_prob"], out["y_hat"], out["site_prob"], out["site_hat"],
               out["a_raw"], out["features"])
        ctx.mark_non_differentiable(*res[2:])
        return res

    @staticmethod
    def backward(ctx, dlogits, dsite_logits, *unused):
        module = ctx.module
        dims = module._dims
        if ctx.saved is None:
            raise RuntimeError("toad_b200: backward called but the forward ran without saved activations")
        h, a_raw, features, stats = ctx.saved_tensors[:4]
        params = ctx.saved_tensors[4:]
        if dlogits is None:
            dlogits = torch.zeros((1, ctx.n_classes), dtype=torch.float32, device=h.device)
        if dsite_logits is None:
            dsite_logits = torch.zeros((1, 2), dtype=torch.float32, device=h.device)
        fwd_out = {"a_raw": a_raw, "features": features, "softmax_stats": stats}
        flat = ops.toad_bwd(dims, [p.detach() for p in params], h, fwd_out, ctx.saved, dlogits, dsite_logits,
                            module._ws_bwd, flags=_default_flags())
        off = ops.param_offsets(dims)
        grads = tuple(flat[off[i]:off[i + 1]].view_as(p) if ctx.needs_input_grad[3 + i] else None
                      for i, p in enumerate(params))
        ctx.saved = None
        ctx.module = None
        return (None, None, None) + grads


_DP_PREFIX = "attention_net.module."     # key form of checkpoints written by the reference on a multi-GPU host


class TOAD_fc_mtl_concat(nn.Module):
    """TOAD multi-task attention-MIL classifier (reference models/model_toad.py:53-116).

    args: gate (only True is valid -- the reference's gate=False branch raises NameError,
    model_toad.py:68), size_arg "big"|"small", dropout, n_classes.
    """

    def __init__(self, gate: bool = True, size_arg: str = "big", dropout: bool = False, n_classes: int = 2):
        super().__init__()
        self.size_dict = {"small": [1024, 512, 256], "big": [1024, 512, 384]}
        size = self.size_dict[size_arg]
        fc: List[nn.Module] = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        fc.extend([nn.Linear(size[1], size[1]), nn.ReLU()])
        if dropout:
            fc.append(nn.Dropout(0.25))
        if not gate:
            raise NameError("name 'Attn_Net' is not defined")  # what the reference does (model_toad.py:68)
        fc.append(Attn_Net_Gated(L=size[1], D=size[2], dropout=dropout, n_tasks=2))
        self.attention_net = nn.Sequential(*fc)
        self.classifier = nn.Linear(size[1] + 1, n_classes)
        self.site_classifier = nn.Linear(size[1] + 1, 2)
        initialize_weights(self)
        self.dropout = bool(dropout)
        self._dims = ops.make_dims(size[0], size[1], size[2], n_classes)
        self._ws_by_stream: Dict[int, ops.Workspace] = {}   # forward scratch, one per CUDA stream (concurrent slides)
        self._ws_bwd = ops.Workspace()
        self._prof = None  # optional ops.Profile handle (bench.py roofline leg)
        self._plane_state: Dict[int, object] = {}  # per stream: (parameter versions, workspace identity) of the cached planes
        self._plane_key_pending = None
        self._register_load_state_dict_pre_hook(self._strip_dataparallel_prefix)
        self.register_load_state_dict_post_hook(self._warn_missing_keys)

    # -- checkpoint compatibility with the reference on multi-GPU hosts.  There `relocate()` wraps attention_net in
    # nn.DataParallel (model_toad.py:79-82), so its checkpoints name the trunk `attention_net.module.<i>...`; the
    # callers load with strict=False (eval_utils_mtl_concat.py:28-30), which would silently leave 10 of the 14 tensors at
    # their random initial values.  Accept both key forms; never load a partial TOAD checkpoint silently.
    @staticmethod
    def _strip_dataparallel_prefix(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        dp = prefix + _DP_PREFIX
        for k in [k for k in state_dict if k.startswith(dp)]:
            state_dict[prefix + "attention_net." + k[len(dp):]] = state_dict.pop(k)

    @staticmethod
    def _warn_missing_keys(module, incompatible_keys):
        if incompatible_keys.missing_keys:
            import warnings
            warnings.warn("toad_b200: load_state_dict left %d of this model's tensors untouched (strict=False hides "
                          "this): %s" % (len(incompatible_keys.missing_keys), incompatible_keys.missing_keys[:4]),
                          RuntimeWarning, stacklevel=3)

    def state_dict_dataparallel(self) -> Dict[str, torch.Tensor]:
        """state_dict() with the trunk keys in the `attention_net.module.*` form a multi-GPU reference process expects
        (its attention_net is an nn.DataParallel, model_toad.py:79-82)."""
        return {(_DP_PREFIX + k[len("attention_net."):]) if k.startswith("attention_net.") else k: v
                for k, v in self.state_dict().items()}

    def invalidate_weight_cache(self) -> None:
        """Forget the cached bf16 weight planes / validated parameter block.  They are keyed on each parameter's
        (data_ptr, autograd version); writes that bypass the version counter -- `p.data.copy_(...)`, `p.data.mul_()`,
        raw-pointer writes from another library -- must be followed by this call (optimizers, `load_state_dict`,
        in-place torch ops on the parameter itself and FusedTrainStep are tracked automatically)."""
        self._plane_state.clear()
        self.__dict__.pop("_pcache", None)

    # -- parameters in C-ABI (= state_dict) order
    def _param_list(self) -> List[torch.Tensor]:
        cached = self.__dict__.get("_plist")
        if cached is not None:   # the Parameter objects are fixed at construction (.to() / load_state_dict keep them)
            return cached
        fc1 = self.attention_net[0]
        fc2 = self.attention_net[3 if self.dropout else 2]
        gate = self.attention_net[-1]
        plist = [fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                 gate.attention_a[0].weight, gate.attention_a[0].bias,
                 gate.attention_b[0].weight, gate.attention_b[0].bias,
                 gate.attention_c.weight, gate.attention_c.bias,
                 self.classifier.weight, self.classifier.bias,
                 self.site_classifier.weight, self.site_classifier.bias]
        self.__dict__["_plist"] = plist
        return plist

    def relocate(self) -> None:
        """Move to the GPU (reference model_toad.py:77-88).

        The reference wraps attention_net in nn.DataParallel and splits one bag across GPUs;
        here one process drives one GPU and whole slides are sharded across ranks
        (toad_b200.distributed), so this only moves parameters to the current CUDA device.
        """
        if not torch.cuda.is_available():
            raise RuntimeError("toad_b200 needs a CUDA device (B200); there is no CPU path")
        device = torch.device("cuda", torch.cuda.current_device())
        self.attention_net = self.attention_net.to(device)
        self.classifier = self.classifier.to(device)
        self.site_classifier = self.site_classifier.to(device)

    @property
    def _ws(self) -> "ops.Workspace":
        """Forward workspace of the CURRENT stream: forwards issued on different streams (two slides in
        flight so that one's partial last wave overlaps the other's kernels) must not share scratch."""
        sid = torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0
        ws = self._ws_by_stream.get(sid)
        if ws is None:
            ws = self._ws_by_stream[sid] = ops.Workspace()
        return ws

    def _weight_plane_flag(self, params, device, pkey=None) -> int:
        """Inference loops reuse the bf16 weight planes that the previous forward left in the workspace
        as long as no parameter changed (tensor identity + autograd version counter) and the workspace
        buffer is still the same allocation (it is grow-only, and the planes sit at n-independent offsets
        -- but a regrown buffer has lost them)."""
        self._plane_key_pending = None
        if _default_flags() & _lib.FLAG_SIMT_FP32:
            return 0
        key = (pkey if pkey is not None else tuple((p.data_ptr(), p._version) for p in params)) + (str(device),)
        ws = self._ws
        buf = ws.buf
        state = (key, None if buf is None else buf.data_ptr(), None if buf is None else buf.numel())
        reuse = buf is not None and self._plane_state.get(id(ws)) == state
        self._plane_key_pending = key
        return _lib.FLAG_REUSE_WEIGHT_PLANES if reuse else 0

    def _note_planes_written(self) -> None:
        ws = self._ws
        buf = ws.buf
        self._plane_state[id(ws)] = None if self._plane_key_pending is None or buf is None else \
            (self._plane_key_pending, buf.data_ptr(), buf.numel())

    @staticmethod
    def _flags() -> int:
        return _default_flags()

    def _dropout_active(self) -> bool:
        return self.dropout and self.training

    @torch.no_grad()
    def forward_batch(self, h: torch.Tensor, lengths, sex: torch.Tensor, return_features: bool = False) -> List[Dict[str, torch.Tensor]]:
        """Eval forward of several slides at once (beyond the reference's surface, which always runs batch_size 1):
        `h` is the bags concatenated along dim 0 -- what collate_MIL_mtl_concat (utils/utils.py:30-35) builds --,
        `lengths` their patch counts, `sex` one value per slide.  Returns one results_dict per slide, identical to
        `forward(h_i, sex_i)` up to fp32 summation order in the pooled quantities (raw scores are bit-identical).
        Small bags share one set of trunk launches and fill the GPU together; chunks of 16 slides per call."""
        if self._dropout_active():
            raise RuntimeError("forward_batch is an inference path: call model.eval() first")
        params = self._param_list()
        lengths = [int(n) for n in lengths]
        sex_f = sex.reshape(-1).to(device=h.device, dtype=torch.float32).contiguous()
        if sex_f.numel() != len(lengths):
            raise ValueError("sex must hold one value per slide")
        pkey = tuple((p.data_ptr(), p._version) for p in params)
        pc = self.__dict__.get("_pcache")
        if pc is None or pc[0] != pkey:
            det = [p.detach() for p in params]
            pc = (pkey, det, ops._params_struct(self._dims, det))
            self.__dict__["_pcache"] = pc
        results: List[Dict[str, torch.Tensor]] = []
        row = 0
        for s0 in range(0, len(lengths), _lib.MAX_BATCH):
            chunk = lengths[s0:s0 + _lib.MAX_BATCH]
            n_chunk = sum(chunk)
            out = ops.toad_fwd_batch(self._dims, pc[1], h[row:row + n_chunk], chunk, sex_f[s0:s0 + len(chunk)], self._ws,
                                     _default_flags() | self._weight_plane_flag(params, h.device, pkey), pstruct=pc[2])
            self._note_planes_written()
            off = 0
            for i, n in enumerate(chunk):
                r: Dict[str, torch.Tensor] = {}
                if return_features:
                    r["features"] = out["features"][i]
                r.update({"logits": out["logits"][i:i + 1], "Y_prob": out["y_prob"][i:i + 1], "Y_hat": out["y_hat"][i],
                          "site_logits": out["site_logits"][i:i + 1], "site_prob": out["site_prob"][i:i + 1],
                          "site_hat": out["site_hat"][i], "A": out["a_raw"][:, off:off + n]})
                results.append(r)
                off += n
            row += n_chunk
        return results

    def forward(self, h: torch.Tensor, sex: torch.Tensor, return_features: bool = False,
                attention_only: bool = False):
        params = self._param_list()
        if attention_only:
            out = ops.toad_fwd(self._dims, [p.detach() for p in params], h, sex, self._ws,
                               _default_flags() | _lib.FLAG_ATTENTION_ONLY)
            return out["a_raw"][0]
        if not isinstance(sex, torch.Tensor):
            raise ValueError("sex must be a tensor")
        sex_f = sex.reshape(-1).to(device=h.device, dtype=torch.float32) if isinstance(h, torch.Tensor) and h.is_cuda else sex
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if need_grad or self._dropout_active():
            (logits, site_logits, y_prob, y_hat, site_prob, site_hat, a_raw, features) = _ToadFunction.apply(
                self, h, sex_f, *params)
        else:
            # host fast path: the detached views and the validated C parameter block are cached per parameter
            # version (identity + autograd version counter), so a steady-state forward re-checks nothing
            pkey = tuple((p.data_ptr(), p._version) for p in params)
            pc = self.__dict__.get("_pcache")
            if pc is None or pc[0] != pkey:
                det = [p.detach() for p in params]
                pc = (pkey, det, ops._params_struct(self._dims, det))
                self.__dict__["_pcache"] = pc
            out = ops.toad_fwd(self._dims, pc[1], h, sex_f, self._ws,
                               _default_flags() | self._weight_plane_flag(params, h.device, pkey), prof=self._prof,
                               pstruct=pc[2])
            self._note_planes_written()
            logits, site_logits, y_prob, y_hat = out["logits"], out["site_logits"], out["y_prob"], out["y_hat"]
            site_prob, site_hat, a_raw, features = out["site_prob"], out["site_hat"], out["a_raw"], out["features"]
        results_dict: Dict[str, torch.Tensor] = {}
        if return_features:
            results_dict.update({"features": features})
        results_dict.update({"logits": logits, "Y_prob": y_prob, "Y_hat": y_hat, "site_logits": site_logits,
                             "site_prob": site_prob, "site_hat": site_hat, "A": a_raw})
        return results_dict
