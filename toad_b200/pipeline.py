"""Host->device streaming of slides for inference: pinned host bags are copied on a side stream
into a small ring of device buffers while the previous slide computes (the reference does a
blocking pageable `data.to(device)` per bag, utils/core_utils_mtl_concat.py:201).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

import torch


class SlideStreamer:
    """Runs `model(h, sex)` over an iterable of (pinned host bag [N,1024] fp32, sex float) pairs.

    Device-side ring of `depth` buffers; H2D on a copy stream overlapped with compute; the small
    result tensors of every slide are copied back to pinned host memory (D2H) on the compute stream.
    """

    def __init__(self, model, max_patches: int, width: int = 1024, depth: int = 2, device: Optional[torch.device] = None):
        self.model = model
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.depth = depth
        self.bufs = [torch.empty((max_patches, width), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.freed = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @torch.no_grad()
    def run(self, slides: Iterable[Tuple[torch.Tensor, float]],
            sink: Optional[Callable[[int, dict], None]] = None,
            forward: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None) -> List[dict]:
        """forward(i, device bag view, sex tensor): replaces `model(bag, sex)` + the per-slide result copies (the
        caller keeps the results on the device, e.g. toad_b200.eval.SlideEvaluator's tables)."""
        compute = torch.cuda.current_stream(self.device)
        results: List[dict] = []
        pending = []
        it = iter(slides)

        def stage(slot: int, item):
            bag, sex = item
            n = bag.shape[0]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.freed[slot])
                self.bufs[slot][:n].copy_(bag, non_blocking=True)
                self.copied[slot].record(self.copy_stream)
            self.h2d_bytes += bag.numel() * 4
            return (slot, n, sex)

        for s in range(self.depth):
            self.freed[s].record(compute)
        nxt = next(it, None)
        slot = 0
        staged = stage(slot, nxt) if nxt is not None else None
        i = 0
        while staged is not None:
            cur = staged
            if i >= 1:
                # the source may recycle the host buffer of item i-1 as soon as item i+1 is requested
                # (ring-buffer loaders): its H2D copy must have completed, not merely been enqueued
                self.copied[(cur[0] + self.depth - 1) % self.depth].synchronize()
            nxt = next(it, None)
            staged = stage((cur[0] + 1) % self.depth, nxt) if nxt is not None else None
            cslot, n, sex = cur
            compute.wait_event(self.copied[cslot])
            sex_t = torch.tensor([float(sex)], device=self.device)
            if forward is not None:
                forward(i, self.bufs[cslot][:n], sex_t)
                self.freed[cslot].record(compute)
                i += 1
                continue
            out = self.model(self.bufs[cslot][:n], sex_t)
            self.freed[cslot].record(compute)
            host = {k: out[k].to("cpu", non_blocking=True) for k in ("logits", "Y_prob", "Y_hat", "site_prob", "site_hat")}
            self.d2h_bytes += sum(v.numel() * v.element_size() for v in host.values())
            if sink is not None:
                sink(i, host)
            else:
                results.append(host)
            i += 1
        compute.synchronize()
        return results


class ResidentRunner:
    """Forward over slides that are already resident in device memory, `n_streams` slides in flight.

    The GEMM kernels are persistent (one CTA per SM, static tile lists), so a single slide leaves SMs idle in
    its partially filled last waves (196 pair tiles on 74 CTA pairs at N = 50k) and during the serial merge of
    the pooling tail; with a second slide queued on another CUDA stream those SMs pick up the other slide's
    CTAs.  The module keeps one forward workspace per stream (`TOAD_fc_mtl_concat._ws`).
    """

    def __init__(self, model, n_streams: int = 2, device: Optional[torch.device] = None):
        self.model = model
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, n_streams))]

    @torch.no_grad()
    def run(self, bags: Iterable[torch.Tensor], sexes: Iterable[torch.Tensor]) -> List[dict]:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)            # inputs produced on the caller's stream are visible
        out: List[dict] = []
        for i, (bag, sex) in enumerate(zip(bags, sexes)):
            with torch.cuda.stream(self.streams[i % len(self.streams)]):
                out.append(self.model(bag, sex))
        for s in self.streams:
            cur.wait_stream(s)            # results are ordered before anything the caller enqueues next
        # the result tensors were allocated on the side streams but are consumed on the caller's: tell the caching
        # allocator, or a later side-stream forward could be handed their blocks while `cur` still reads them
        for r in out:
            for t in (r.values() if isinstance(r, dict) else (r,)):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(cur)
        return out
