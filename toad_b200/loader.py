"""Bag loader for the slide pipeline: the reference's on-disk format is one `{slide_id}.pt` per slide
holding an [N, 1024] fp32 tensor (datasets/dataset_mtl_concat.py:368-371, docs/README.md:24), read by
DataLoader workers and moved with a blocking pageable `.to(device)` (utils/core_utils_mtl_concat.py:201).

`PinnedBagLoader` reads the same files on a background thread into a small ring of PINNED host
buffers, so `toad_b200.pipeline.SlideStreamer` can overlap disk read -> H2D copy -> compute.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

import torch


class PinnedBagLoader:
    """Iterates (bag_view [N, width] fp32 in pinned memory, slide_id) over `.pt` files.

    A yielded view stays valid until `depth - 1` further items have been taken (ring of `depth` buffers) --
    exactly what SlideStreamer needs (it has issued the H2D copy of item i before it asks for item i+2).
    """

    def __init__(self, data_dir: str, slide_ids: Sequence[str], max_patches: int, width: int = 1024, depth: int = 3,
                 pin: Optional[bool] = None):
        self.paths: List[Tuple[str, str]] = [(s, os.path.join(data_dir, "%s.pt" % s)) for s in slide_ids]
        self.max_patches, self.width, self.depth = max_patches, width, depth
        pin = torch.cuda.is_available() if pin is None else pin
        self.bufs = [torch.empty((max_patches, width), dtype=torch.float32, pin_memory=pin) for _ in range(depth)]
        self.pinned = pin

    def __len__(self) -> int:
        return len(self.paths)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, str]]:
        free: "queue.Queue[int]" = queue.Queue()
        ready: "queue.Queue[object]" = queue.Queue()
        for i in range(self.depth):
            free.put(i)

        def worker():
            try:
                for slide_id, path in self.paths:
                    t = torch.load(path, map_location="cpu")
                    if t.dim() != 2 or t.shape[1] != self.width or t.dtype != torch.float32:
                        raise ValueError("%s: expected an [N, %d] float32 tensor, got %s %s" % (path, self.width, tuple(t.shape), t.dtype))
                    if t.shape[0] > self.max_patches:
                        raise ValueError("%s: %d patches exceed max_patches=%d" % (path, t.shape[0], self.max_patches))
                    slot = free.get()
                    self.bufs[slot][:t.shape[0]].copy_(t)
                    ready.put((slot, t.shape[0], slide_id))
                ready.put(None)
            except Exception as e:  # surfaced in the consumer thread
                ready.put(e)

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        held: List[int] = []
        while True:
            item = ready.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            slot, n, slide_id = item
            held.append(slot)
            if len(held) >= self.depth:          # the oldest view is no longer in use by contract
                free.put(held.pop(0))
            yield self.bufs[slot][:n], slide_id
        th.join()
