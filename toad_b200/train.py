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
                 w_site: float = 0.25):
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

    # -- checkpointing of the optimizer moments (flat, toad_param_offsets order)
    def state_dict(self) -> Dict[str, object]:
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, state: Dict[str, object]) -> None:
        self.step_count = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])

    def step(self, data: torch.Tensor, label: torch.Tensor, site: torch.Tensor, sex: torch.Tensor,
             return_features: bool = False) -> Dict[str, torch.Tensor]:
        """Returns the reference's results_dict plus 'loss' = device tensor [total, cls, site] (no host sync)."""
        m = self.model
        if not m.training:
            raise RuntimeError("FusedTrainStep.step needs model.train()")
        dims = m._dims
        params = [p.detach() for p in m._param_list()]
        flags = _default_flags() | _lib.FLAG_SAVE_ACTS
        saved = ops.alloc_saved(dims, data.shape[0], data.device, flags)
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
        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            scale = 1.0 / dist.get_world_size()
        self.step_count += 1
        ops.adam_step(dims, params, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas,
                      self.eps, self.weight_decay, scale)
        # the kernel wrote the parameters behind autograd's back: bump their version counters (what an in-place
        # torch op would have done) so the cached bf16 weight planes of the inference path are re-split
        for p in m._param_list():
            torch.autograd.graph.increment_version(p)
        m._plane_state.clear()
        res: Dict[str, torch.Tensor] = {}
        if return_features:
            res["features"] = out["features"]
        res.update({"logits": out["logits"], "Y_prob": out["y_prob"], "Y_hat": out["y_hat"],
                    "site_logits": out["site_logits"], "site_prob": out["site_prob"], "site_hat": out["site_hat"],
                    "A": out["a_raw"], "loss": loss3})
        return res

    __call__ = step
