"""Short forward-only workload for ncu captures (never a bench number)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--train", action="store_true")
    a = ap.parse_args()
    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    torch.manual_seed(0)
    model = TOAD_fc_mtl_concat(n_classes=18)
    model.relocate()
    x = torch.randn(a.n, 1024, device="cuda")
    sex = torch.tensor([1.0], device="cuda")
    if a.train:
        model.train()
        loss_fn = torch.nn.CrossEntropyLoss()
        for _ in range(a.iters):
            r = model(x, sex)
            loss = 0.75 * loss_fn(r["logits"], torch.tensor([3], device="cuda")) + \
                0.25 * loss_fn(r["site_logits"], torch.tensor([1], device="cuda"))
            loss.backward()
    else:
        model.eval()
        with torch.no_grad():
            for _ in range(a.iters):
                model(x, sex)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
