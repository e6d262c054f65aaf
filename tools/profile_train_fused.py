"""Short FusedTrainStep workload for ncu launch lists (never a bench number)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    from toad_b200.train import FusedTrainStep
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    torch.manual_seed(0)
    model = TOAD_fc_mtl_concat(n_classes=18)
    model.relocate()
    model.train()
    step = FusedTrainStep(model, lr=1e-4, weight_decay=1e-5)
    x = torch.randn(n, 1024, device="cuda")
    sex = torch.ones(1, device="cuda")
    lab, site = torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda")
    for _ in range(4):
        step(x, lab, site, sex)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
