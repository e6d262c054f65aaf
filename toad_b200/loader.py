"""Bag loader for the slide pipeline: the reference's on-disk format is one `{slide_id}.pt` per slide
holding an [N, 1024] fp32 tensor (datasets/dataset_mtl_concat.py:368-371, docs/README.md:24), read by
DataLoader workers and moved with a blocking pageable `.to(device)` (utils/core_utils_mtl_concat.py:201) -- or,
with `load_from_h5(True)`, one `{slide_id}.h5` with datasets `features` [N, 1024] and `coords` [N, 2]
(datasets/dataset_mtl_concat.py:375-383), the patch coordinates a heatmap / top-k consumer needs.

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
                 pin: Optional[bool] = None, use_h5: bool = False):
        """use_h5: read `{slide_id}.h5` (`features` + `coords`, the reference's `load_from_h5(True)` mode) instead of
        `{slide_id}.pt`; the iterator then yields (bag_view, slide_id, coords_view [N, 2] int32 in pinned memory).
        Needs the optional `h5py` package, exactly like the reference."""
        self.use_h5 = bool(use_h5)
        ext = "h5" if self.use_h5 else "pt"
        self.paths: List[Tuple[str, str]] = [(s, os.path.join(data_dir, "%s.%s" % (s, ext))) for s in slide_ids]
        self.max_patches, self.width, self.depth = max_patches, width, depth
        pin = torch.cuda.is_available() if pin is None else pin
        self.bufs = [torch.empty((max_patches, width), dtype=torch.float32, pin_memory=pin) for _ in range(depth)]
        self.coord_bufs = [torch.empty((max_patches, 2), dtype=torch.int32, pin_memory=pin) for _ in range(depth)] \
            if self.use_h5 else None
        self.pinned = pin

    def _read(self, path: str):
        """-> (features [N, width] fp32 CPU tensor, coords [N, 2] int32 tensor or None)"""
        if not self.use_h5:
            return torch.load(path, map_location="cpu"), None
        try:
            import h5py
        except ImportError as e:     # same optional dependency as the reference's h5 mode
            raise ImportError("reading .h5 bags needs the `h5py` package (datasets/dataset_mtl_concat.py:377)") from e
        with h5py.File(path, "r") as f:
            feats = torch.from_numpy(f["features"][:])
            coords = torch.from_numpy(f["coords"][:].astype("int32", copy=False))
        if coords.dim() != 2 or coords.shape[1] != 2 or coords.shape[0] != feats.shape[0]:
            raise ValueError("%s: coords must be [N, 2] with one row per patch, got %s" % (path, tuple(coords.shape)))
        return feats, coords

    def __len__(self) -> int:
        return len(self.paths)

    def __iter__(self) -> Iterator[tuple]:
        free: "queue.Queue[int]" = queue.Queue()
        ready: "queue.Queue[object]" = queue.Queue()
        for i in range(self.depth):
            free.put(i)

        def worker():
            try:
                for slide_id, path in self.paths:
                    t, coords = self._read(path)
                    if t.dim() != 2 or t.shape[1] != self.width or t.dtype != torch.float32:
                        raise ValueError("%s: expected an [N, %d] float32 tensor, got %s %s" % (path, self.width, tuple(t.shape), t.dtype))
                    if t.shape[0] > self.max_patches:
                        raise ValueError("%s: %d patches exceed max_patches=%d" % (path, t.shape[0], self.max_patches))
                    slot = free.get()
                    self.bufs[slot][:t.shape[0]].copy_(t)
                    if coords is not None:
                        self.coord_bufs[slot][:t.shape[0]].copy_(coords)
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
            if self.use_h5:
                yield self.bufs[slot][:n], slide_id, self.coord_bufs[slot][:n]
            else:
                yield self.bufs[slot][:n], slide_id
        th.join()
