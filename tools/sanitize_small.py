"""Tiny end-to-end workload for compute-sanitizer (forward, backward, top-k, resnet): correctness tooling only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from models.model_toad import TOAD_fc_mtl_concat
    from models.resnet_custom import resnet50_baseline
    from toad_b200 import ops
    from toad_b200.train import FusedTrainStep
    torch.manual_seed(0)
    model = TOAD_fc_mtl_concat(n_classes=18)
    model.relocate()
    x = torch.randn(257, 1024, device="cuda")
    sex = torch.ones(1, device="cuda")
    model.eval()
    with torch.no_grad():
        r = model(x, sex)
        a = model(x, sex, attention_only=True)
    ops.topk_patches(r["A"][0].contiguous(), 50, torch.arange(514, device="cuda", dtype=torch.int32).reshape(257, 2))
    model.train()
    step = FusedTrainStep(model)
    step(x, torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda"), sex)
    ext = resnet50_baseline().cuda().eval()
    with torch.no_grad():
        ext(torch.randn(2, 3, 64, 64, device="cuda"))
    torch.cuda.synchronize()
    print("sanitize workload done")


if __name__ == "__main__":
    main()
