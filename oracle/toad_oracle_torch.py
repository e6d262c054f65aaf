"""CPU port of the reference forward in functional torch ops -- TEST / BASELINE INFRASTRUCTURE ONLY.

Same arithmetic and the same library kernels (ATen addmm -> MKL sgemm, tanh, sigmoid, softmax,
mm) the reference's nn.Module dispatches to on CPU (models/model_toad.py:36-41, 90-116), written
as plain functions over a parameter dict so it travels to the GPU box (the reference tree does
not).  Used by bench.py for the cpu_baseline leg and the `--impl reference` arm, and pinned
against the reference-generated golden vectors by tests/test_oracle_golden.py.
Nothing in toad_b200/ or models/ imports this.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .toad_oracle import PARAM_KEYS


def to_torch_params(params: dict) -> dict:
    return {k: torch.from_numpy(v.copy()) if not isinstance(v, torch.Tensor) else v for k, v in params.items()}


@torch.no_grad()
def toad_forward(h: torch.Tensor, sex: torch.Tensor, p: dict, return_features: bool = False,
                 attention_only: bool = False):
    """TOAD_fc_mtl_concat.forward restated (model_toad.py:90-116), eval mode / dropout=False."""
    h = F.relu(F.linear(h, p[PARAM_KEYS[0]], p[PARAM_KEYS[1]]))            # :59
    h = F.relu(F.linear(h, p[PARAM_KEYS[2]], p[PARAM_KEYS[3]]))            # :62
    a = torch.tanh(F.linear(h, p[PARAM_KEYS[4]], p[PARAM_KEYS[5]]))        # :37
    b = torch.sigmoid(F.linear(h, p[PARAM_KEYS[6]], p[PARAM_KEYS[7]]))     # :38
    A = F.linear(a.mul(b), p[PARAM_KEYS[8]], p[PARAM_KEYS[9]])             # :39-40
    A = torch.transpose(A, 1, 0)                                           # :92
    if attention_only:
        return A[0]
    A_raw = A
    A = F.softmax(A, dim=1)                                                # :97
    M = torch.mm(A, h)                                                     # :98
    M = torch.cat([M, sex.repeat(M.size(0), 1)], dim=1)                    # :99
    logits = F.linear(M[0].unsqueeze(0), p[PARAM_KEYS[10]], p[PARAM_KEYS[11]])
    site_logits = F.linear(M[1].unsqueeze(0), p[PARAM_KEYS[12]], p[PARAM_KEYS[13]])
    out = {}
    if return_features:
        out["features"] = M
    out.update({"logits": logits, "Y_prob": F.softmax(logits, dim=1), "Y_hat": torch.topk(logits, 1, dim=1)[1],
                "site_logits": site_logits, "site_prob": F.softmax(site_logits, dim=1),
                "site_hat": torch.topk(site_logits, 1, dim=1)[1], "A": A_raw})
    return out
