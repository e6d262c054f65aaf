"""Serial forward() throughput: one model(data, sex) call per slide on one stream.  Run it twice to A/B a library
switch, e.g. TOAD_B200_PDL=0 python tools/fwd_modes.py vs TOAD_B200_PDL=1 (programmatic dependent launch).
Tuning / evidence tool, not a bench number.

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
    res = {"pdl": os.environ.get("TOAD_B200_PDL", "1")}
    for n in [int(a) for a in sys.argv[1:]] or [10000, 50000]:
        bags = [torch.randn(n, 1024, device="cuda") for _ in range(4)]
        reps = max(40, int(8e6 // n))
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
            res["n%d_slides_per_s" % n] = round(reps / (e0.elapsed_time(e1) / 1e3), 1)
            out = model(bags[0], sex, return_features=True)
            res["n%d_logit0" % n] = float(out["logits"][0, 0])
        del bags
    print(json.dumps(res))


if __name__ == "__main__":
    main()
