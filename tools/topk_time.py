"""Timing of toad_topk (config 5: N = 200k giga-slide, k in {1,10,100,1000}) next to torch.topk."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from toad_b200 import ops
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    s = torch.randn(n, device="cuda")
    res = {"n": n}
    for k in (1, 10, 100, 1000):
        for name, fn in (("ours", lambda: ops.topk(s, k)), ("torch", lambda: torch.topk(s, k))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res["%s_k%d_us" % (name, k)] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
