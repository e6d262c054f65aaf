"""resnet50_baseline throughput (config 3 shape: 3x256x256 patches) for both arithmetic modes, next to the same network
in PyTorch eager on the same GPU (cuDNN fp32 and cuDNN with TF32 allowed = torch's default for convolutions) -- the
reference's own GPU path (models/resnet_custom.py run by torch).  Prints one JSON line.

    python tools/resnet_bench.py --batch 512 --reps 4 [--no-eager] [--precisions f16x2,bf16x3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--precisions", default="f16x2,bf16x3")
    a = ap.parse_args()
    import numpy as np
    import torch
    from models.resnet_custom import resnet50_baseline
    from oracle import resnet_oracle as RO
    torch.manual_seed(0)
    params = RO.make_params(1)
    sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in params.items()}
    x = torch.randn(a.batch, 3, a.size, a.size, device="cuda")
    out = {"batch": a.batch, "size": a.size, "chunk": os.environ.get("TOAD_RESNET_CHUNK", "default")}
    feats = {}
    with torch.no_grad():
        for prec in a.precisions.split(","):
            m = resnet50_baseline(pretrained=False)
            m.load_state_dict(sd, strict=True)
            m = m.cuda().eval()
            m.precision = prec
            t = timed(lambda: m(x), a.reps)
            out["patches_per_s_" + prec] = round(a.batch / t, 1)
            feats[prec] = m(x[:8].contiguous())
        if not a.no_eager:
            # the reference network in eager torch on this GPU (oracle = the reference's own library calls)
            dp = {k: (v.cuda() if v.dim() else v) for k, v in sd.items()}
            nb = min(a.batch, 128)
            xe = x[:nb]
            for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True
                t = timed(lambda: RO.resnet50_baseline_forward(xe, dp), max(2, a.reps // 2))
                out["patches_per_s_eager_" + name] = round(nb / t, 1)
                feats[name] = RO.resnet50_baseline_forward(x[:8].contiguous(), dp)
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            # accuracy of every arm against the fp64 reference arithmetic on 8 patches (CPU, fp64)
            ref = RO.resnet50_baseline_forward(x[:8].cpu().double(), params)
            scale = float(ref.abs().max())
            out["err_over_scale"] = {k: float((v.cpu().double() - ref).abs().max()) / scale for k, v in feats.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
