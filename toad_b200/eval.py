"""Slide-level evaluation and the consumers of the attention scores.

* `SlideEvaluator` -- the loop of the reference's `summary()` (utils/eval_utils_mtl_concat.py:65-177) without its
  per-slide host round trips: every forward writes its probabilities / predictions straight into one row of
  device-resident tables, and the tables come back to the host ONCE at the end (the reference does
  `.cpu().numpy()` / `.item()` on five tensors per slide, eval_utils_mtl_concat.py:99-107).
* `PatchPipeline` -- patches -> resnet50_baseline -> TOAD with the [N, 1024] feature matrix kept on the GPU
  (the reference round-trips it through `{slide_id}.pt` files, docs/README.md:24-39).
* `ops.topk_patches` gives the top-k patches (and their coords) of an attention row for heatmaps.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .pipeline import SlideStreamer


def topk_accuracy(probs: np.ndarray, labels: np.ndarray, topk: Sequence[int] = (1,)) -> list:
    """Fraction of slides whose label is among the k most probable classes -- `accuracy()` of
    utils/eval_utils_mtl_concat.py:49-63 on host arrays (ties resolved like a stable descending sort)."""
    probs = np.asarray(probs)
    labels = np.asarray(labels).astype(np.int64)
    order = np.argsort(-probs, axis=1, kind="stable")
    return [float((order[:, :k] == labels[:, None]).any(axis=1).mean()) for k in topk]


class SlideEvaluator:
    """Runs `model` (toad_b200 TOAD_fc_mtl_concat, eval mode) over slides and aggregates like `summary()`."""

    def __init__(self, model, max_slides: int, max_patches: int, width: int = 1024, device: Optional[torch.device] = None):
        self.model = model
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        d = model._dims
        f32 = dict(dtype=torch.float32, device=self.device)
        self.max_slides = max_slides
        # one row per slide; the C ABI writes them in place (toad_fwd_out_t pointers into these tables)
        self.probs = torch.zeros((max_slides, d.n_classes), **f32)
        self.site_probs = torch.zeros((max_slides, 2), **f32)
        self.logits = torch.zeros((max_slides, d.n_classes), **f32)
        self.site_logits = torch.zeros((max_slides, 2), **f32)
        self.hats = torch.zeros((max_slides, 2), dtype=torch.int64, device=self.device)   # (Y_hat, site_hat)
        self._scratch = {"features": torch.empty((d.n_tasks, d.hid_dim + 1), **f32),
                         "softmax_stats": torch.empty((d.n_tasks, 2), **f32)}
        self.streamer = SlideStreamer(model, max_patches, width, depth=2, device=self.device)

    def _forward_into(self, i: int, bag: torch.Tensor, sex: torch.Tensor) -> None:
        m = self.model
        n = bag.shape[0]
        out = dict(self._scratch)
        out.update({"a_raw": torch.empty((m._dims.n_tasks, n), dtype=torch.float32, device=self.device),
                    "logits": self.logits[i:i + 1], "y_prob": self.probs[i:i + 1], "y_hat": self.hats[i:i + 1, 0:1],
                    "site_logits": self.site_logits[i:i + 1], "site_prob": self.site_probs[i:i + 1],
                    "site_hat": self.hats[i:i + 1, 1:2]})
        params = m._param_list()
        ops.toad_fwd(m._dims, [p.detach() for p in params], bag, sex, m._ws,
                     m._flags() | m._weight_plane_flag(params, bag.device), out=out)
        m._note_planes_written()

    @torch.no_grad()
    def run(self, slides: Iterable[Tuple[torch.Tensor, float]], labels: Optional[Sequence[int]] = None,
            sites: Optional[Sequence[int]] = None, topk: Sequence[int] = (1, 3, 5)) -> Dict[str, object]:
        """slides: (pinned host bag [N, 1024] fp32, sex) pairs.  Returns the arrays `summary()` builds
        (all_cls_probs, all_site_probs, predictions) and, when labels / sites are given, its error rates and
        top-k accuracies.  One device->host transfer per table."""
        if self.model.training:
            raise RuntimeError("SlideEvaluator needs model.eval() (summary() calls model.eval(), eval_utils:69)")
        count = [0]

        def fwd(i, bag, sex_t):
            if i >= self.max_slides:
                raise ValueError("more slides than max_slides=%d" % self.max_slides)
            self._forward_into(i, bag, sex_t)
            count[0] = i + 1

        self.streamer.run(slides, forward=fwd)
        s = count[0]
        host = {"all_cls_probs": self.probs[:s].cpu().numpy(), "all_site_probs": self.site_probs[:s].cpu().numpy(),
                "cls_logits": self.logits[:s].cpu().numpy(), "site_logits": self.site_logits[:s].cpu().numpy()}
        hats = self.hats[:s].cpu().numpy()
        host["Y_hat"], host["site_hat"] = hats[:, 0].copy(), hats[:, 1].copy()
        host["n_slides"] = s
        if labels is not None:
            lab = np.asarray(labels[:s], dtype=np.int64)
            host["cls_test_error"] = float((host["Y_hat"] != lab).mean())      # calculate_error, utils/utils.py:135-138
            ks = [k for k in topk if k <= host["all_cls_probs"].shape[1]]
            host["topk_acc"] = dict(zip(ks, topk_accuracy(host["all_cls_probs"], lab, ks)))
        if sites is not None:
            st = np.asarray(sites[:s], dtype=np.int64)
            host["site_test_error"] = float((host["site_hat"] != st).mean())
        return host


class PatchPipeline:
    """patches [N, 3, H, W] -> features [N, 1024] (resnet50_baseline, batches of `batch`) -> TOAD forward, the
    feature matrix never leaving the GPU: the trunk's average-pool kernel writes each batch straight into its rows
    of the slide's feature buffer, which the attention-MIL forward then reads once."""

    def __init__(self, extractor, classifier, max_patches: int, batch: int = 256, device: Optional[torch.device] = None):
        self.extractor, self.classifier, self.batch = extractor, classifier, batch
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.features = torch.empty((max_patches, 1024), dtype=torch.float32, device=self.device)

    @torch.no_grad()
    def run(self, patch_batches: Iterable[torch.Tensor], sex: torch.Tensor, **forward_kwargs):
        """patch_batches: CUDA (or pinned host) fp32 tensors [b, 3, H, W], b <= batch each; sex: float tensor [1]."""
        n = 0
        for pb in patch_batches:
            if not pb.is_cuda:
                pb = pb.to(self.device, non_blocking=True)
            b = pb.shape[0]
            if n + b > self.features.shape[0]:
                raise ValueError("more patches than max_patches=%d" % self.features.shape[0])
            self.extractor(pb, out=self.features[n:n + b])
            n += b
        if n == 0:
            raise ValueError("empty slide")
        return self.classifier(self.features[:n], sex, **forward_kwargs), self.features[:n]
