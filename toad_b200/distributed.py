"""One-slide-per-GPU data parallelism.

One process per GPU (torchrun); slides are independent units, so inference needs no collective
at all (replicas), and training needs exactly one: an all-reduce of the flat fp32 gradient
(1,192,490 floats = 4.77 MB for the "big" model) per optimizer step over NCCL/NVLink.  This
replaces the reference's single-process nn.DataParallel that splits ONE bag's patches across
GPUs and gathers N x 512 activations back to cuda:0 (models/model_toad.py:77-88).

Semantics note: with G ranks one optimizer step averages the gradients of G slides; the
reference steps after every slide (utils/utils.py:51-55 batch_size=1).  Per-slide
forward/backward parity is the tested contract.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn


def init_from_env(backend: Optional[str] = None) -> Dict[str, int]:
    """Initialise torch.distributed from torchrun's environment; returns rank/world/local_rank."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if use_cuda else "gloo")
        kwargs = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return {"rank": rank, "world": world, "local_rank": local}


def shard_slides(lengths: Sequence[int], rank: int, world: int, balance: bool = True) -> List[int]:
    """Slide indices owned by `rank`.

    balance=False: round-robin (slide i -> rank i mod world).  balance=True: longest-processing-
    time greedy packing on the patch counts, since bag sizes span 5k-80k patches (16x cost
    spread) and the slowest rank sets the step time.  Deterministic, identical on every rank.
    """
    n = len(lengths)
    if not balance:
        return [i for i in range(n) if i % world == rank]
    order = sorted(range(n), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    owner = [0] * n
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += int(lengths[i])
    return [i for i in range(n) if owner[i] == rank]


def aligned_rounds(lengths: Sequence[int], world: int, seed: Optional[int] = None) -> List[List[int]]:
    """Synchronous training rounds of `world` slides of SIMILAR size (length bucketing).

    A data-parallel step ends with an all-reduce, so it lasts as long as its largest slide: pairing whatever slide is
    s-th on every rank makes each step wait for the max of `world` draws from the size distribution (N in [5k, 80k]:
    ~1.7x the mean at world = 8).  Here the slides are sorted by patch count and cut into consecutive groups of
    `world`; a group is one step, one slide per rank.  `seed` shuffles the ORDER of the groups (the sampler's
    randomness is kept at group granularity) and rotates the rank assignment inside a group so that no rank always
    draws the group's largest slide.  rounds[s][r] = slide index for rank r in step s, -1 = no slide (tail group).
    Deterministic, identical on every rank."""
    n = len(lengths)
    order = sorted(range(n), key=lambda i: (-int(lengths[i]), i))
    rounds = [order[i:i + world] for i in range(0, n, world)]
    if seed is not None:
        import random
        rng = random.Random(seed)
        rng.shuffle(rounds)
        rounds = [r[k:] + r[:k] for r in rounds for k in [rng.randrange(len(r))]]
    return [r + [-1] * (world - len(r)) for r in rounds]


class FlatGradBucket:
    """All parameter gradients as views of ONE flat fp32 buffer -> one all-reduce per step."""

    def __init__(self, module: nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("all parameters must share one device and dtype")
            p.grad = self.flat[off:off + p.numel()].view_as(p)   # autograd accumulates in place
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, average: bool = True) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())


def train_step(model: nn.Module, optimizer: torch.optim.Optimizer, bucket: FlatGradBucket, data: torch.Tensor,
               label: torch.Tensor, site: torch.Tensor, sex: torch.Tensor, loss_fn=None) -> Dict[str, float]:
    """One training step of the reference loop (utils/core_utils_mtl_concat.py:206-234) on this
    rank's slide, with the gradient averaged over ranks before the (identical) optimizer step."""
    loss_fn = loss_fn or nn.CrossEntropyLoss()
    results = model(data, sex)
    cls_loss = loss_fn(results["logits"], label)
    site_loss = loss_fn(results["site_logits"], site)
    loss = cls_loss * 0.75 + site_loss * 0.25
    loss.backward()
    bucket.allreduce(average=True)
    optimizer.step()
    bucket.zero()   # instead of optimizer.zero_grad(): keep .grad aliased to the flat buffer
    return {"cls_loss": float(cls_loss.item()), "site_loss": float(site_loss.item())}
