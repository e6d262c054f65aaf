"""One slide's training step of the reference loop with loss and optimizer fused.

The reference (utils/core_utils_mtl_concat.py:198-234) does, per slide::

    results = model(data, sex)
    loss = 0.75 * CE(results['logits'], label) + 0.25 * CE(results['site_logits'], site)
    loss.backward(); optimizer.step(); optimizer.zero_grad()          # optimizer = Adam (utils/utils.py:65)

With the drop-in module that already runs on this library's forward / backward kernels, but the loss is ~25
eager launches and Adam ~110 (8 multi-tensor passes over 14 tensors): about 0.25 ms of GPU time in a 1.7 ms step at
N = 50k, and most of the step for small bags.  ``FusedTrainStep`` keeps the same arithmetic and issues

    toad_fwd (SAVE_ACTS) -> toad_ce_loss_grad -> toad_bwd -> [one NCCL all-reduce of the flat gradient] -> toad_adam_step

with no autograd graph, no per-parameter .grad tensors and no host synchronisation.  The parameters stay the module's
own ``nn.Parameter`` tensors (updated in place), so ``state_dict()`` / evaluation work unchanged.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from .model_toad import TOAD_fc_mtl_concat, _default_flags


class FusedTrainStep:
    """``step(data, label, site, sex)`` = forward + weighted CE + backward + (all-reduce) + Adam, in place.

    lr / weight_decay / betas / eps have torch.optim.Adam's meaning (weight decay is the L2 term the reference
    passes as ``args.reg``).  With an initialised process group of world size G the flat gradient is summed over
    ranks and scaled by 1/G inside the Adam kernel: one optimizer step per G slides, identical on every rank.
    """

    def __init__(self, model: TOAD_fc_mtl_concat, lr: float = 1e-4, weight_decay: float = 1e-5,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, w_cls: float = 0.75,
                 w_site: float = 0.25, max_patches: int = 0):
        if not isinstance(model, TOAD_fc_mtl_concat):
            raise ValueError("FusedTrainStep drives toad_b200's TOAD_fc_mtl_concat")
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), tuple(betas), float(eps)
        self.w_cls, self.w_site = float(w_cls), float(w_site)
        self.step_count = 0
        params = model._param_list()
        if not all(p.is_cuda for p in params):
            raise ValueError("call model.relocate() first: toad_b200 has no CPU path")
        total = ops.param_offsets(model._dims)[14]
        dev = params[0].device
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        # saved activations (~7 KB per patch): ONE grow-only set of buffers, sliced per slide.  Allocating them per step
        # with a different N every time misses the caching allocator and turns into synchronous cudaMallocs in the
        # middle of a data-parallel step; `max_patches` sizes them up front.
        self._saved_cap = 0
        self._saved_full: Optional[Dict[str, torch.Tensor]] = None
        self.allreduce = True         # False: skip the gradient all-reduce even inside a process group (benchmarks)
        self.timing_events = None     # optional [e_start, e_bwd_done, e_allreduce_done, e_adam_done] recorded by step()
        if max_patches > 0:
            self._saved_for(int(max_patches), dev, _default_flags() | _lib.FLAG_SAVE_ACTS)

    def _saved_for(self, n: int, device, flags: int) -> Dict[str, torch.Tensor]:
        if self._saved_full is None or n > self._saved_cap:
            self._saved_full = None                       # release before growing
            self._saved_full = ops.alloc_saved(self.model._dims, n, device, flags)
            self._saved_cap = n
        return {k: v[:n] for k, v in self._saved_full.items()}

    def _finish(self, params, scale: float) -> None:
        m = self.model
        ev = self.timing_events
        if ev is not None:
            ev[1].record()
        if self.allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
        if ev is not None:
            ev[2].record()
        self.step_count += 1
        ops.adam_step(m._dims, params, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas,
                      self.eps, self.weight_decay, scale)
        if ev is not None:
            ev[3].record()
        # the kernel wrote the parameters behind autograd's back: bump their version counters (what an in-place
        # torch op would have done) so the cached bf16 weight planes of the inference path are re-split
        for p in m._param_list():
            torch.autograd.graph.increment_version(p)
        m._plane_state.clear()

    def step_idle(self, n_slides_in_round: int) -> None:
        """This rank has no slide in the current round (tail of an epoch, distributed.aligned_rounds pads with -1): it
        contributes a zero gradient to the all-reduce and applies the same averaged update as every other rank."""
        params = [p.detach() for p in self.model._param_list()]
        if self.timing_events is not None:
            self.timing_events[0].record()
        self.grad.zero_()
        self._finish(params, 1.0 / max(1, int(n_slides_in_round)))

    # -- checkpointing of the optimizer moments (flat, toad_param_offsets order)
    def state_dict(self) -> Dict[str, object]:
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, state: Dict[str, object]) -> None:
        self.step_count = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])

    def step(self, data: torch.Tensor, label: torch.Tensor, site: torch.Tensor, sex: torch.Tensor,
             return_features: bool = False, n_slides_in_round: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Returns the reference's results_dict plus 'loss' = device tensor [total, cls, site] (no host sync).
        n_slides_in_round: slides contributing to this step's all-reduce (default: the world size)."""
        m = self.model
        if not m.training:
            raise RuntimeError("FusedTrainStep.step needs model.train()")
        dims = m._dims
        params = [p.detach() for p in m._param_list()]
        flags = _default_flags() | _lib.FLAG_SAVE_ACTS
        if self.timing_events is not None:
            self.timing_events[0].record()
        saved = self._saved_for(data.shape[0], data.device, flags)
        if m._dropout_active():
            flags |= _lib.FLAG_DROPOUT
            saved["dropout_seed"] = int(torch.randint(0, 2 ** 62, (1,)).item())
            saved["dropout_p"] = 0.25
        sex_f = sex.reshape(-1).to(device=data.device, dtype=torch.float32)
        out = ops.toad_fwd(dims, params, data, sex_f, m._ws, flags, saved)
        lab = label.reshape(-1).to(device=data.device, dtype=torch.int64)
        sit = site.reshape(-1).to(device=data.device, dtype=torch.int64)
        loss3, dl, ds = ops.ce_loss_grad(out["logits"].reshape(-1), out["site_logits"].reshape(-1), lab, sit,
                                         self.w_cls, self.w_site)
        ops.toad_bwd(dims, params, data, out, saved, dl, ds, m._ws_bwd, grad_flat=self.grad,
                     flags=_default_flags())
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._finish(params, 1.0 / (n_slides_in_round if n_slides_in_round else world))
        res: Dict[str, torch.Tensor] = {}
        if return_features:
            res["features"] = out["features"]
        res.update({"logits": out["logits"], "Y_prob": out["y_prob"], "Y_hat": out["y_hat"],
                    "site_logits": out["site_logits"], "site_prob": out["site_prob"], "site_hat": out["site_hat"],
                    "A": out["a_raw"], "loss": loss3})
        return res

    __call__ = step
