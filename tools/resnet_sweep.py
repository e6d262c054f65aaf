"""resnet50_baseline throughput vs batch size (L2 residency of inter-layer activations vs wave filling)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from models.resnet_custom import resnet50_baseline
    torch.manual_seed(1)
    m = resnet50_baseline(pretrained=False).cuda().eval()
    total = 256
    x = torch.randn(total, 3, 256, 256, device="cuda")
    res = {}
    with torch.no_grad():
        for b in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256]:
            for _ in range(2):
                for i in range(0, total, b):
                    m(x[i:i + b])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 3
            for _ in range(reps):
                for i in range(0, total, b):
                    m(x[i:i + b])
            e1.record()
            torch.cuda.synchronize()
            res[b] = round(total * reps / (e0.elapsed_time(e1) / 1e3), 1)
    print(json.dumps({"patches_per_s_by_batch": res}))


if __name__ == "__main__":
    main()
