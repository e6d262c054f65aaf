"""Config 4 (BASELINE.json): training on synthetic slides, one slide per GPU per step, one NCCL
all-reduce of the flat gradient per optimizer step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_synthetic.py --slides 64

Slides: N_i ~ randint(5000, 80001) (seed 7), features generated on the GPU per slide (never stored),
labels randint(18), site randint(2), sex randint(2); Adam lr 1e-4 wd 1e-5 (main_mtl_concat.py:85-88).
Prints one JSON line from rank 0 (slides/s, step time, loss trajectory).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slides", type=int, default=64)
    ap.add_argument("--min-n", type=int, default=5000)
    ap.add_argument("--max-n", type=int, default=80000)
    ap.add_argument("--eager", action="store_true",
                    help="module + nn.CrossEntropyLoss + torch.optim.Adam (the reference loop) instead of FusedTrainStep")
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from toad_b200.distributed import FlatGradBucket, init_from_env, shard_slides, train_step
    from models.model_toad import TOAD_fc_mtl_concat

    info = init_from_env()
    rank, world = info["rank"], info["world"]
    dev = torch.device("cuda", info["local_rank"])
    rng = np.random.default_rng(7)
    lengths = rng.integers(a.min_n, a.max_n + 1, size=a.slides).tolist()
    labels = rng.integers(0, 18, size=a.slides).tolist()
    sites = rng.integers(0, 2, size=a.slides).tolist()
    sexes = rng.integers(0, 2, size=a.slides).tolist()
    mine = shard_slides(lengths, rank, world, balance=True)
    steps = min(len(shard_slides(lengths, r, world)) for r in range(world))   # every rank steps together

    torch.manual_seed(0)                       # identical initial weights on every rank
    model = TOAD_fc_mtl_concat(n_classes=18)
    model.relocate()
    model.train()
    if a.eager:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
        bucket = FlatGradBucket(model)
    else:
        from toad_b200.train import FusedTrainStep
        fused = FusedTrainStep(model, lr=1e-4, weight_decay=1e-5)
    gen = torch.Generator(device=dev)
    losses = []
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    patches = 0
    for s in range(steps):
        i = mine[s]
        gen.manual_seed(1000 + i)
        x = torch.randn(lengths[i], 1024, generator=gen, device=dev)
        lab, sit = torch.tensor([labels[i]], device=dev), torch.tensor([sites[i]], device=dev)
        sx = torch.tensor([float(sexes[i])], device=dev)
        if a.eager:
            out = train_step(model, opt, bucket, x, lab, sit, sx)
            losses.append(out["cls_loss"])
        else:
            losses.append(fused(x, lab, sit, sx)["loss"])      # stays on the device: no host sync per step
        patches += lengths[i]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if not a.eager:
        losses = [float(l[1].item()) for l in losses]
    # all ranks hold identical parameters after identical averaged steps
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, flat))
        t = torch.tensor([float(same), float(patches)], device=dev)
        dist.all_reduce(t)
        same_all, patches_all = t[0].item() == world, t[1].item()
    else:
        same_all, patches_all = True, float(patches)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "steps": steps, "slides": steps * world, "seconds": dt,
                          "slides_per_s": steps * world / dt, "patches_per_s": patches_all / dt,
                          "loop": "eager" if a.eager else "fused", "params_identical_across_ranks": same_all, "first_losses": losses[:3], "last_losses": losses[-3:]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
