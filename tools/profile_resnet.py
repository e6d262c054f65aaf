"""Short resnet50_baseline workload for ncu captures (never a bench number)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    import torch
    from models.resnet_custom import resnet50_baseline
    torch.manual_seed(0)
    m = resnet50_baseline().cuda().eval()
    x = torch.randn(a.batch, 3, 256, 256, device="cuda")
    with torch.no_grad():
        for _ in range(a.iters):
            m(x)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
