"""GPU tests of the evaluation loop, the attention top-k consumer and the patch -> slide pipeline
(SURVEY.md section 8f rows 2 and 4): results must equal the plain per-slide module calls."""
import numpy as np
import pytest
import torch

from oracle import toad_oracle as O
from tests.helpers import build_model, to_np

pytestmark = pytest.mark.gpu


def test_slide_evaluator_equals_per_slide_forward():
    from toad_b200.eval import SlideEvaluator
    params = O.make_params(5, "big", 18, 0.02)
    model = build_model(params, "big", 18)
    ns = [700, 1, 2048, 333, 1500]
    bags = [torch.from_numpy(O.make_bag(50 + i, n)).pin_memory() for i, n in enumerate(ns)]
    sexes = [float(i % 2) for i in range(len(ns))]
    ev = SlideEvaluator(model, max_slides=8, max_patches=2048)
    labels = [3, 0, 17, 5, 9]
    sites = [0, 1, 1, 0, 1]
    res = ev.run(zip(bags, sexes), labels=labels, sites=sites, topk=(1, 3, 5))
    assert res["n_slides"] == len(ns)
    for i, (bag, sex) in enumerate(zip(bags, sexes)):
        with torch.no_grad():
            r = model(bag.cuda(), torch.tensor([sex], device="cuda"))
        np.testing.assert_array_equal(res["all_cls_probs"][i], to_np(r["Y_prob"])[0])
        np.testing.assert_array_equal(res["all_site_probs"][i], to_np(r["site_prob"])[0])
        np.testing.assert_array_equal(res["cls_logits"][i], to_np(r["logits"])[0])
        assert res["Y_hat"][i] == int(r["Y_hat"]) and res["site_hat"][i] == int(r["site_hat"])
    assert res["cls_test_error"] == float(np.mean(res["Y_hat"] != np.array(labels)))
    assert res["site_test_error"] == float(np.mean(res["site_hat"] != np.array(sites)))
    assert set(res["topk_acc"]) == {1, 3, 5} and res["topk_acc"][1] == pytest.approx(1.0 - res["cls_test_error"])
    # a second run on the same evaluator (tables reused) gives the same answer
    res2 = ev.run(zip(bags, sexes))
    np.testing.assert_array_equal(res2["all_cls_probs"], res["all_cls_probs"])


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64, torch.float32])
def test_topk_patches_with_coords(dtype):
    from toad_b200 import ops
    g = torch.Generator().manual_seed(3)
    n = 20000
    scores = torch.randn(n, generator=g).cuda()
    coords = torch.randint(0, 100000, (n, 2), generator=g).to(dtype).cuda()
    for k in (1, 10, 1000):
        vals, idx, c = ops.topk_patches(scores, k, coords)
        tv, ti = torch.topk(scores, k)
        assert torch.equal(vals, tv) and torch.equal(idx, ti)
        assert c.dtype == dtype and torch.equal(c, coords[ti])
    with pytest.raises(ValueError):
        ops.topk_patches(scores, 5, coords[:100])


def test_patch_pipeline_equals_sequential_modules():
    from models.resnet_custom import resnet50_baseline
    from toad_b200.eval import PatchPipeline
    torch.manual_seed(2)
    ext = resnet50_baseline().cuda().eval()
    for m in ext.modules():   # non-trivial BatchNorm statistics so that folding is exercised
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    model = build_model(O.make_params(7, "big", 18, 0.02), "big", 18)
    x = torch.randn(40, 3, 64, 64, device="cuda")
    sex = torch.tensor([1.0], device="cuda")
    pipe = PatchPipeline(ext, model, max_patches=64, batch=16)
    out, feats = pipe.run([x[0:16], x[16:32], x[32:40]], sex, return_features=True)
    with torch.no_grad():
        f_ref = torch.cat([ext(x[0:16]), ext(x[16:32]), ext(x[32:40])], 0)
        r_ref = model(f_ref, sex, return_features=True)
    assert torch.equal(feats, f_ref)
    for k in ("logits", "site_logits", "Y_prob", "A", "features"):
        assert torch.equal(out[k], r_ref[k]), k
    # the referee: the reference's arithmetic end to end (oracle trunk -> oracle TOAD head), fp64
    from oracle import resnet_oracle as RO
    sd = {k: v.detach().cpu().numpy() for k, v in ext.state_dict().items()}
    f64 = RO.resnet50_baseline_forward(x.cpu().double(), sd).numpy()
    fscale = np.abs(f64).max()
    assert np.abs(to_np(feats) - f64).max() <= 1e-3 * fscale
    oref = O.toad_forward(f64, 1.0, O.make_params(7, "big", 18, 0.02), dtype=np.float64)
    lscale = np.abs(oref["logits"]).max()
    assert np.abs(to_np(out["logits"]) - oref["logits"]).max() <= 2e-3 * lscale
    np.testing.assert_allclose(to_np(out["Y_prob"]), oref["Y_prob"], rtol=1e-2, atol=1e-4)


@pytest.mark.parametrize("simt", [False, True])
def test_forward_batch_equals_per_slide_forward(simt, monkeypatch):
    """toad_fwd_batch: slides back to back in one matrix, one set of trunk launches.  Raw scores are bit-identical to
    the per-slide forward (row-wise independent GEMMs, fixed-order partial sums); pooled results agree to fp32
    summation order; the oracle is the referee for the heads."""
    monkeypatch.setenv("TOAD_B200_SIMT", "1" if simt else "0")
    params = O.make_params(11, "big", 18, 0.02)
    model = build_model(params, "big", 18)
    ns = [300, 1, 2500, 257, 4097, 33] + [64] * 12          # 18 slides: two C-ABI chunks (16 + 2)
    bags = [O.make_bag(700 + i, n) for i, n in enumerate(ns)]
    sexes = torch.tensor([float(i % 2) for i in range(len(ns))], device="cuda")
    h = torch.from_numpy(np.concatenate(bags, 0)).cuda()
    res = model.forward_batch(h, ns, sexes, return_features=True)
    assert len(res) == len(ns)
    for i, (bag, n) in enumerate(zip(bags, ns)):
        with torch.no_grad():
            r = model(torch.from_numpy(bag).cuda(), sexes[i:i + 1], return_features=True)
        assert res[i]["A"].shape == (2, n) and res[i]["logits"].shape == (1, 18) and res[i]["Y_hat"].shape == (1, 1)
        assert torch.equal(res[i]["A"], r["A"]), i
        for k in ("logits", "site_logits", "Y_prob", "site_prob", "features"):
            np.testing.assert_allclose(to_np(res[i][k]), to_np(r[k]), rtol=2e-5, atol=2e-6, err_msg="%s[%d]" % (k, i))
        if i < 6:
            ref = O.toad_forward(bag, float(i % 2), params, dtype=np.float64)
            np.testing.assert_allclose(to_np(res[i]["logits"]), ref["logits"], rtol=1e-3, atol=2e-5)   # logits that happen to be ~0 get an absolute floor
            assert int(res[i]["Y_hat"]) == int(np.asarray(ref["Y_hat"]).reshape(-1)[0])
            assert int(res[i]["site_hat"]) == int(np.asarray(ref["site_hat"]).reshape(-1)[0])
    with pytest.raises(ValueError):
        model.forward_batch(h, [h.shape[0] - 1, 0, 1], sexes[:3])
