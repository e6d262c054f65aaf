"""Serial forward() throughput (one model(data, sex) call per slide on one stream) with the trunk on one stream vs split
over two streams (toad_fwd_2s), and a bit-equality check of the two.  Tuning / evidence tool, not a bench number.

    python tools/fwd_modes.py [N ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from toad_b200 import model_toad as MT
    torch.manual_seed(0)
    model = MT.TOAD_fc_mtl_concat(n_classes=18)
    model.relocate()
    model.eval()
    sex = torch.tensor([1.0], device="cuda")
    res = {}
    for n in [int(a) for a in sys.argv[1:]] or [10000, 25000, 50000, 100000]:
        bags = [torch.randn(n, 1024, device="cuda") for _ in range(4)]
        reps = max(40, int(4e6 // n))
        outs = {}
        for mode, thr in (("one_stream", 1 << 60), ("two_streams", 0)):
            MT.TWO_STREAM_MIN_PATCHES = thr
            with torch.no_grad():
                for i in range(6):
                    model(bags[i % 4], sex)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(reps):
                    model(bags[i % 4], sex)
                e1.record()
                torch.cuda.synchronize()
                res["n%d_%s_slides_per_s" % (n, mode)] = round(reps / (e0.elapsed_time(e1) / 1e3), 1)
                outs[mode] = model(bags[0], sex, return_features=True)
        a, b = outs["one_stream"], outs["two_streams"]
        res["n%d_bit_identical" % n] = bool(all(torch.equal(a[k], b[k]) for k in a))
        del bags
    print(json.dumps(res))


if __name__ == "__main__":
    main()
