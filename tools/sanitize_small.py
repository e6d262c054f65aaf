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
        for prec in ("f16x2", "bf16x3"):       # both arithmetic modes; shapes whose conv tiles are clipped / hold several images
            ext.precision = prec
            for shape in ((2, 3, 64, 64), (3, 3, 96, 160), (1, 3, 224, 224), (1, 3, 32, 512), (5, 3, 16, 16),
                          (3, 3, 128, 128), (2, 3, 64, 256)):   # the last two: halo 3x3 convolutions, stem with fused max-pool
                ext(torch.randn(*shape, device="cuda"))
    # standalone gated-attention block: training forward with dropout + backward (parameters and x)
    from models.model_toad import Attn_Net_Gated
    blk = Attn_Net_Gated(L=512, D=384, dropout=True, n_tasks=2).cuda().train()
    xg = torch.randn(300, 512, device="cuda", requires_grad=True)
    A, _ = blk(xg)
    A.sum().backward()
    torch.cuda.synchronize()
    print("sanitize workload done")


if __name__ == "__main__":
    main()
