"""First-contact GPU diagnostics: each step in its own process (a trapped kernel poisons the
context), bounded by a timeout, results appended to gpurun_out/diag.jsonl, arrays dumped for
offline analysis when something mismatches.

    python tools/gpu_diag.py            # run all steps
    python tools/gpu_diag.py --step X   # one step, in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

STEPS = ["lin_64_f32", "lin_64_split", "lin_128", "lin_256", "lin_multi", "pair_64_f32", "pair_64_split", "pair_256",
         "pair_multi", "wide_small", "wide_multi", "attn_gated", "simt_fwd", "tc_fwd_small", "tc_fwd_10k", "tc_fwd_10k_cg1", "topk", "bwd_simt",
         "bwd_tc", "timing", "lin_timing", "resnet_s64", "resnet_s256", "resnet_timing", "train_timing", "train_breakdown", "train_spike", "train_fused"]


def log(rec):
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def lin(step, m, n, k, variant, dump=True):
    import numpy as np
    import torch
    from oracle import toad_oracle as O
    from toad_b200 import ops
    rng = np.random.default_rng(1)
    x = rng.standard_normal((m, k), dtype=np.float32)
    w = (rng.standard_normal((n, k), dtype=np.float32) / np.float32(np.sqrt(k))).astype(np.float32)
    ws = ops.Workspace()
    y = ops.linear_bf16x3(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), None, False, ws, variant)
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    xh, xl = O.split_bf16(x)
    wh, wl = O.split_bf16(w)
    exp = O.linear_bf16x3(xh, xl, wh, wl)
    err = float(np.abs(y - exp).max())
    rec = {"step": step, "m": m, "n": n, "k": k, "variant": variant, "max_abs_err": err,
           "ok": bool(err < 1e-4), "y_absmax": float(np.abs(y).max()), "nan": bool(np.isnan(y).any())}
    if not rec["ok"] and dump:
        np.savez_compressed(os.path.join(OUT, step + ".npz"), y=y, exp=exp.astype(np.float32), x=x, w=w,
                            hihi=(xh.astype(np.float64) @ wh.T.astype(np.float64)).astype(np.float32))
    return rec


def fwd_case(step, name, simt):
    import numpy as np
    import torch
    from tests.helpers import build_model, case_inputs, load_golden, to_np
    os.environ["TOAD_B200_SIMT"] = "1" if simt else "0"
    g = load_golden(name)
    params, x, sex = case_inputs(g)
    model = build_model(params, str(g["meta_size_arg"]), int(g["meta_n_classes"]))
    with torch.no_grad():
        out = model(torch.from_numpy(x).cuda(), torch.tensor([sex], device="cuda"), return_features=True)
    torch.cuda.synchronize()
    rec = {"step": step, "case": name, "simt": simt}
    for k in ("logits", "site_logits", "features", "A"):
        o = to_np(out[k]).astype(np.float64)
        r = g["f64_" + k]
        rec["abs_" + k] = float(np.abs(o - r).max())
        rec["rel_" + k] = float((np.abs(o - r) / np.maximum(np.abs(r), 1e-6)).max())
    rec["ok"] = bool(rec["abs_A"] < 1e-4 and rec["rel_logits"] < 1e-3)
    if not rec["ok"]:
        np.savez_compressed(os.path.join(OUT, step + ".npz"), **{k: to_np(v) for k, v in out.items()})
    return rec


def bwd_case(step, simt):
    from tests.test_gpu_backward import _grads
    import numpy as np
    from tests.helpers import to_np
    g, model, loss = _grads("toad_big_n257", simt)
    rec = {"step": step, "loss": loss, "loss_ref": float(g["f64_loss"])}
    worst = 0.0
    for k, prm in model.named_parameters():
        gk = to_np(prm.grad).astype(np.float64)
        if ("g64_%s__full" % k) in g:
            ref = g["g64_%s__full" % k]
            e = np.abs(gk - ref).max() / (np.abs(ref).max() + 1e-12)
        else:
            ref = g["g64_%s__sub" % k]
            e = np.abs(gk[::37, ::41] - ref).max() / (np.abs(ref).max() + 1e-12)
        rec["e_" + k] = float(e)
        worst = max(worst, float(e))
    rec["ok"] = bool(worst < 2e-3)
    return rec


def timing(step):
    import numpy as np
    import torch
    from oracle import toad_oracle as O
    from tests.helpers import build_model
    from toad_b200 import ops, _lib
    rec = {"step": step}
    params = O.make_params(0, "big", 18)
    model = build_model(params, "big", 18)
    for n in (10000, 50000):
        x = torch.randn(n, 1024, device="cuda")
        sd = torch.tensor([1.0], device="cuda")
        for mode in ("tc", "tc_cg1", "tc_cg2", "simt"):
            simt = mode == "simt"
            prof = ops.Profile(16)
            plist = [p.detach() for p in model._param_list()]
            flags = _lib.FLAG_SIMT_FP32 if simt else {"tc_cg1": _lib.FLAG_TC_SINGLE_CTA, "tc_cg2": _lib.FLAG_TC_PAIR_ALL}.get(mode, 0)
            for _ in range(3):
                ops.toad_fwd(model._dims, plist, x, sd, model._ws, flags)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                ops.toad_fwd(model._dims, plist, x, sd, model._ws, flags, prof=prof.handle)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            stages, calls = prof.read()
            rec["n%d_%s_ms" % (n, mode)] = dt * 1e3
            rec["n%d_%s_stages_ms" % (n, mode)] = {k: round(v / max(calls, 1), 4) for k, v in stages.items()}
    rec["ok"] = True
    return rec


def run_step(step):
    if step == "lin_64_f32":
        return lin(step, 128, 64, 64, 0x10)
    if step == "lin_64_split":
        return lin(step, 128, 64, 64, 0x11)
    if step == "lin_128":
        return lin(step, 128, 128, 128, 0x20)
    if step == "lin_256":
        return lin(step, 256, 256, 256, 0x31)
    if step == "lin_multi":
        return lin(step, 40000, 512, 1024, 0x00, dump=False)
    if step == "pair_64_f32":
        return lin(step, 256, 64, 64, 0x12)
    if step == "pair_64_split":
        return lin(step, 256, 64, 64, 0x13)
    if step == "pair_256":
        return lin(step, 300, 256, 256, 0x33)
    if step == "pair_multi":
        return lin(step, 40000, 512, 1024, 0x02, dump=False)
    if step == "wide_small":
        return lin(step, 300, 512, 128, 0x42)
    if step == "wide_multi":
        return lin(step, 50000, 512, 1024, 0x42, dump=False)
    if step == "tc_fwd_10k_cg1":
        os.environ["TOAD_B200_CG1"] = "1"
        return fwd_case(step, "toad_big_n10000", False)
    if step == "attn_gated":
        import numpy as np
        import torch
        from oracle import toad_oracle as O
        from models.model_toad import Attn_Net_Gated
        from tests.helpers import load_golden, to_np
        g = load_golden("attn_gated_default_n256")
        p = O.make_attn_params(int(g["meta_seed"]), 1024, 256, 1)
        x = O.make_bag(int(g["meta_seed"]) + 1, 256, width=1024)
        net = Attn_Net_Gated()
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        net = net.cuda().eval()
        with torch.no_grad():
            A, _ = net(torch.from_numpy(x).cuda())
        e = float(np.abs(to_np(A) - g["f64_A"]).max())
        return {"step": step, "max_abs_err": e, "ok": bool(e < 1e-4)}
    if step == "simt_fwd":
        return fwd_case(step, "toad_big_n257", True)
    if step == "tc_fwd_small":
        return fwd_case(step, "toad_big_n257", False)
    if step == "tc_fwd_10k":
        return fwd_case(step, "toad_big_n10000", False)
    if step == "topk":
        import numpy as np
        import torch
        from oracle import toad_oracle as O
        from toad_b200 import ops
        s = O.make_bag(5, 200000, width=1)[:, 0].copy()
        vals, idx = ops.topk(torch.from_numpy(s).cuda(), 1000)
        ev, ei = O.topk_indices(s, 1000)
        return {"step": step, "ok": bool(np.array_equal(idx.cpu().numpy(), ei) and np.array_equal(vals.cpu().numpy(), ev))}
    if step == "bwd_simt":
        return bwd_case(step, True)
    if step == "bwd_tc":
        return bwd_case(step, False)
    if step == "timing":
        return timing(step)
    if step == "train_timing":
        import torch
        from models.model_toad import TOAD_fc_mtl_concat
        rec = {"step": step, "ok": True}
        for mode in ("tc", "simt"):
            os.environ["TOAD_B200_SIMT"] = "1" if mode == "simt" else "0"
            torch.manual_seed(0)
            model = TOAD_fc_mtl_concat(n_classes=18)
            model.relocate()
            model.train()
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
            x = torch.randn(50000, 1024, device="cuda")
            sex = torch.ones(1, device="cuda")
            lab, site = torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda")
            ce = torch.nn.CrossEntropyLoss()

            def one(with_opt=True):
                r = model(x, sex)
                loss = 0.75 * ce(r["logits"], lab) + 0.25 * ce(r["site_logits"], site)
                loss.backward()
                if with_opt:
                    opt.step()
                    opt.zero_grad()
            for _ in range(3):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                one()
            e1.record()
            torch.cuda.synchronize()
            rec[mode + "_train_step_ms"] = round(e0.elapsed_time(e1) / 10, 3)
        os.environ["TOAD_B200_SIMT"] = "0"
        return rec
    if step == "train_breakdown":
        import torch
        from models.model_toad import TOAD_fc_mtl_concat
        os.environ["TOAD_B200_SIMT"] = "0"
        torch.manual_seed(0)
        model = TOAD_fc_mtl_concat(n_classes=18)
        model.relocate()
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
        x = torch.randn(50000, 1024, device="cuda")
        sex = torch.ones(1, device="cuda")
        lab, site = torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda")
        ce = torch.nn.CrossEntropyLoss()
        rec = {"step": step, "ok": True, "iters": []}
        for it in range(12):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = model(x, sex)
            t1 = time.perf_counter()
            loss = 0.75 * ce(r["logits"], lab) + 0.25 * ce(r["site_logits"], site)
            t2 = time.perf_counter()
            loss.backward()
            t3 = time.perf_counter()
            opt.step()
            opt.zero_grad()
            t4 = time.perf_counter()
            torch.cuda.synchronize()
            t5 = time.perf_counter()
            rec["iters"].append([round(1e3 * v, 3) for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)])
        return rec
    if step == "train_fused":
        # FusedTrainStep vs the eager reference loop on the same module, N = 50k and a small bag
        import torch
        from models.model_toad import TOAD_fc_mtl_concat
        from toad_b200.train import FusedTrainStep
        rec = {"step": step, "ok": True}
        for n in (50000, 2000):
            torch.manual_seed(0)
            model = TOAD_fc_mtl_concat(n_classes=18)
            model.relocate()
            model.train()
            x = torch.randn(n, 1024, device="cuda")
            sex = torch.ones(1, device="cuda")
            lab, site = torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda")
            ce = torch.nn.CrossEntropyLoss()
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
            fused = FusedTrainStep(model, lr=1e-4, weight_decay=1e-5)

            def eager_step():
                r = model(x, sex)
                loss = 0.75 * ce(r["logits"], lab) + 0.25 * ce(r["site_logits"], site)
                loss.backward()
                opt.step()
                opt.zero_grad()

            def fused_step():
                fused(x, lab, site, sex)

            for name, fn in (("eager", eager_step), ("fused", fused_step)):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                for _ in range(30):
                    fn()
                e1.record()
                t_issue = time.perf_counter() - t0
                torch.cuda.synchronize()
                rec["%s_n%d_ms" % (name, n)] = round(e0.elapsed_time(e1) / 30, 4)
                rec["%s_n%d_host_issue_ms" % (name, n)] = round(1e3 * t_issue / 30, 4)
        return rec
    if step == "train_spike":
        # where do the periodic multi-ms host stalls in the training loop come from?  Log Python GC events
        # (generation, duration) against per-iteration times, then repeat with the collector frozen.
        import gc
        import torch
        from models.model_toad import TOAD_fc_mtl_concat
        torch.manual_seed(0)
        model = TOAD_fc_mtl_concat(n_classes=18)
        model.relocate()
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
        x = torch.randn(50000, 1024, device="cuda")
        sex = torch.ones(1, device="cuda")
        lab, site = torch.tensor([3], device="cuda"), torch.tensor([1], device="cuda")
        ce = torch.nn.CrossEntropyLoss()
        events = []
        t_gc = [0.0]

        def cb(phase, info):
            if phase == "start":
                t_gc[0] = time.perf_counter()
            else:
                events.append((info["generation"], round(1e3 * (time.perf_counter() - t_gc[0]), 3), info["collected"]))
        gc.callbacks.append(cb)
        rec = {"step": step, "ok": True}
        for mode in ("gc_on", "gc_frozen"):
            if mode == "gc_frozen":
                gc.collect()
                gc.freeze()
                gc.disable()
            its = []
            for it in range(25):
                torch.cuda.synchronize()
                n_ev = len(events)
                t0 = time.perf_counter()
                r = model(x, sex)
                loss = 0.75 * ce(r["logits"], lab) + 0.25 * ce(r["site_logits"], site)
                loss.backward()
                t1 = time.perf_counter()
                opt.step()
                t2 = time.perf_counter()
                opt.zero_grad()
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                its.append([round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2), round(1e3 * (t3 - t0), 2), events[n_ev:]])
            rec[mode] = its
        rec["mem_alloc_retries"] = torch.cuda.memory_stats().get("num_alloc_retries", -1)
        rec["num_device_alloc"] = torch.cuda.memory_stats().get("num_device_alloc", -1)
        return rec
    if step == "lin_timing":
        import torch
        from toad_b200 import ops
        rec = {"step": step, "ok": True}
        ws = ops.Workspace()
        for (m, n, k) in ((50000, 512, 1024), (50000, 512, 512), (50000, 768, 512)):
            x = torch.randn(m, k, device="cuda")
            w = torch.randn(n, k, device="cuda") / k ** 0.5
            b = torch.zeros(n, device="cuda")
            for variant in (0x00, 0x01, 0x02, 0x03):
                for _ in range(3):
                    ops.linear_bf16x3(x, w, b, True, ws, variant)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    ops.linear_bf16x3(x, w, b, True, ws, variant)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                rec["m%d_n%d_k%d_v%d_ms" % (m, n, k, variant)] = round(ms, 4)
        return rec
    if step in ("resnet_s64", "resnet_s256"):
        import numpy as np
        import torch
        from oracle import resnet_oracle as RO
        from tests.test_gpu_resnet import build
        name = "resnet_b2_s64" if step == "resnet_s64" else "resnet_b2_s256"
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        params = RO.make_params(int(z["meta_pseed"]))
        x = RO.make_images(int(z["meta_xseed"]), int(z["meta_batch"]), int(z["meta_size"]))
        model = build(params)
        with torch.no_grad():
            y = model(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = z["f64_out"]
        err = float(np.abs(y - ref).max())
        rec = {"step": step, "max_abs_err": err, "scale": float(np.abs(ref).max()), "nan": bool(np.isnan(y).any())}
        rec["ok"] = bool(err <= 1e-3 * rec["scale"])
        if not rec["ok"]:
            np.savez_compressed(os.path.join(OUT, step + ".npz"), y=y, ref=ref)
        return rec
    if step == "resnet_timing":
        import torch
        from oracle import resnet_oracle as RO
        from tests.test_gpu_resnet import build
        model = build(RO.make_params(1))
        rec = {"step": step, "ok": True}
        for B in (64, 256):
            x = torch.randn(B, 3, 256, 256, device="cuda")
            with torch.no_grad():
                model(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    model(x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            rec["b%d_ms" % B] = round(ms, 3)
            rec["b%d_patches_per_s" % B] = round(B / ms * 1e3, 1)
        return rec
    raise SystemExit("unknown step " + step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=120)
    a = ap.parse_args()
    if a.step:
        try:
            log(run_step(a.step))
        except Exception as e:
            log({"step": a.step, "ok": False, "error": repr(e), "tb": traceback.format_exc()[-1500:]})
            sys.exit(1)
        return
    steps = [s for s in STEPS if not a.only or s in a.only.split(",")]
    for s in steps:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--step", s], timeout=a.timeout,
                               capture_output=True, text=True)
            if r.returncode != 0:
                log({"step": s, "ok": False, "rc": r.returncode, "stderr": r.stderr[-1500:], "secs": time.time() - t0})
        except subprocess.TimeoutExpired:
            log({"step": s, "ok": False, "error": "timeout", "secs": time.time() - t0})


if __name__ == "__main__":
    main()
