#!/usr/bin/env python
"""bench.py -- slides/s of the TOAD attention-MIL forward at N=50k x 1024 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (1 process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One step = one pass of the hot path over `slides_per_step` synthetic slides (N x 1024 fp32
each): by default one `TOAD_fc_mtl_concat.forward_batch` call over the step's 16 slides (bags back to back, one set
of trunk launches, per-slide pooling); `--batch 1` runs one forward per slide through
`toad_b200.pipeline.ResidentRunner` (three slides in flight on three CUDA streams).  `value` = whole-job slides/s with
bags resident in HBM (4 distinct 205 MB bags per GPU, larger than L2, rotated); `value_single_stream`
= the same steps strictly serial; `e2e` = the same metric through the public API with pinned host
bags copied H2D inside the timed region and results read back.  Prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PATCHES = 50000
WIDTH = 1024
FLOP_PER_PATCH = 2362880            # reference forward, "big", 2 tasks
FC1_FLOP_PER_PATCH = 2 * 1024 * 512  # dominant kernel: fc1 GEMM
BYTES_PER_PATCH = 4096 + 8
TAIL_BYTES_PER_PATCH = 2056          # h row as (hi, lo) bf16 planes + 2 scores


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML every 5 ms
    (nvidia_ml_py), falling back to polling nvidia-smi."""

    HW_SLOWDOWN, SW_POWER_CAP, SW_THERMAL, HW_THERMAL = 0x8, 0x4, 0x20, 0x40

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.sm = []      # (perf_counter timestamp, MHz)
        self.t0 = None    # timed region [t0, t1]; samples outside are dropped in summary()
        self.t1 = None
        self.max_mhz = None
        self.reason_bits = 0
        self.source = None
        self.stop_flag = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            while not self.stop_flag.is_set():
                self.sm.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))))
                if self.t0 is not None:
                    try:
                        self.reason_bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        self.reason_bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.stop_flag.wait(0.002)
            return
        except Exception:
            pass
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [self.HW_SLOWDOWN, self.HW_THERMAL, self.SW_THERMAL, self.SW_POWER_CAP]
        while not self.stop_flag.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in r.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.sm.append((time.perf_counter(), float(parts[0])))
                    self.max_mhz = float(parts[1])
                    for i, b in enumerate(bits):
                        if parts[2 + i].lower().startswith("active"):
                            self.reason_bits |= b
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        inside = [m for (t, m) in self.sm if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30)]
        if not inside:
            inside = [m for (_, m) in self.sm[-3:]]
        self.sm = inside
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        names = [(self.HW_SLOWDOWN, "hw_slowdown"), (self.HW_THERMAL, "hw_thermal_slowdown"),
                 (self.SW_THERMAL, "sw_thermal_slowdown"), (self.SW_POWER_CAP, "sw_power_cap")]
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in names if self.reason_bits & b], "samples": len(self.sm), "source": self.source}


def cpu_reference_rate(n_patches: int, steps: int, warmup: int, budget_s: float):
    """Reference CPU path (oracle's functional-torch port) on the host cores: slides/s.

    The thread count is calibrated first (all cores, half, quarter, ... -- one forward each) and the
    fastest setting is used, so the baseline is the best the reference's library path does here."""
    import torch
    from oracle import toad_oracle as O
    from oracle import toad_oracle_torch as OT
    cores = os.cpu_count() or 1
    params = OT.to_torch_params(O.make_params(0, "big", 18))
    g = torch.Generator().manual_seed(1)
    bags = [torch.randn(n_patches, WIDTH, generator=g) for _ in range(2)]
    sex = torch.tensor([1.0])
    best_t, best_threads = None, cores
    cand = sorted({cores, max(1, cores // 2), max(1, cores // 4), max(1, cores // 8)}, reverse=True)
    for th in cand:
        torch.set_num_threads(th)
        OT.toad_forward(bags[0], sex, params)                    # warm this setting
        t0 = time.perf_counter()
        OT.toad_forward(bags[1], sex, params)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_threads = dt, th
    torch.set_num_threads(best_threads)
    for i in range(warmup):
        OT.toad_forward(bags[i % 2], sex, params)
    times = []
    t_start = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        OT.toad_forward(bags[i % 2], sex, params)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    total = sum(times)
    return len(times) / total, len(times), best_threads, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, done, cores, total = cpu_reference_rate(N_PATCHES, args.steps, args.warmup, budget_s=150.0)
    sample = "%d forwards of one %dx%d fp32 bag on %d host threads (best of a thread-count sweep on %d cores; torch CPU ops, oracle port of models/model_toad.py)" % (
        done, N_PATCHES, WIDTH, cores, os.cpu_count() or 1)
    line = {
        "impl": "reference", "metric": "slides_per_sec_n50k", "value": rate, "unit": "slides/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * total / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "TOAD_fc_mtl_concat forward, N=%d x %d fp32, big, n_classes=18" % (N_PATCHES, WIDTH),
                   "slides_per_step": 1},
        "cpu_baseline": {"value": rate, "unit": "slides/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from toad_b200 import ops
    from toad_b200.pipeline import SlideStreamer
    from models.model_toad import TOAD_fc_mtl_concat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION/INFO in some environments) goes
        # to stdout too, so keep NCCL at WARN unless the caller insists
        if os.environ.get("TOAD_BENCH_KEEP_NCCL_DEBUG", "0") != "1":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    S = args.slides_per_step
    n = args.n_patches

    torch.manual_seed(0)
    model = TOAD_fc_mtl_concat(n_classes=18)      # reference init: xavier-normal weights, zero biases
    model.relocate()
    model.eval()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    n_bags = 4
    bags = [torch.randn(n, WIDTH, generator=g, device=dev) for _ in range(n_bags)]   # 4 x 205 MB > L2
    sex = torch.tensor([1.0], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from toad_b200.pipeline import ResidentRunner
    runner = ResidentRunner(model, n_streams=args.streams, device=dev)

    B = max(1, min(args.batch, 16, S))
    while S % B:
        B -= 1
    if B > 1:   # small bags: B slides back to back per forward_batch call (one set of trunk launches for all of them)
        cats = [torch.cat([bags[(j + b) % n_bags] for b in range(B)], 0) for j in range(2)]
        sexes_b = torch.ones(B, device=dev)

    def step(i):   # one step = one batch of S resident slides through the public runner (--streams slides in flight)
        if B > 1:
            for c in range(S // B):
                model.forward_batch(cats[(i + c) % 2], [n] * B, sexes_b)
            return
        runner.run([bags[(i * S + s) % n_bags] for s in range(S)], [sex] * S)

    def step_serial(i):
        with torch.no_grad():
            for s in range(S):
                model(bags[(i * S + s) % n_bags], sex)

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                       # NVML init happens during the warm-up, sampling every 2 ms
    for i in range(args.warmup):
        step(i)
    prof = ops.Profile(args.steps * S)
    model._prof = prof.handle
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    if sampler:
        sampler.t1 = time.perf_counter()
        sampler.stop_flag.set()
    elapsed_ms = ev0.elapsed_time(ev1)
    model._prof = None
    stages, calls = prof.read()
    prof.close()
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_slides = args.steps * S * world
    value = total_slides / (elapsed_ms / 1e3)

    # ---- extra: the same K steps strictly serial on one stream: clean per-kernel durations for the roofline
    # explanation (in the timed region above two slides share the SMs, which stretches every kernel's wall time)
    prof2 = ops.Profile(args.steps * S)
    model._prof = prof2.handle
    for i in range(min(args.warmup, 2)):
        step_serial(i)
    prof2.read()
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_serial(i)
    ev1.record()
    barrier()
    model._prof = None
    stages_serial, calls_serial = prof2.read()
    prof2.close()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value_single_stream = total_slides / (float(t.item()) / 1e3)

    # ---- e2e: pinned host bags -> H2D -> forward -> D2H results, through the public API
    host_bags = [torch.randn(n, WIDTH).pin_memory() for _ in range(2)]
    streamer = SlideStreamer(model, n, WIDTH, depth=2, device=dev)
    e2e_slides = max(4, min(args.steps * S, 32))
    streamer.run([(host_bags[i % 2], 1.0) for i in range(3)])      # warm-up
    streamer.h2d_bytes = streamer.d2h_bytes = 0
    barrier()
    ev0.record()
    streamer.run([(host_bags[i % 2], float(i % 2)) for i in range(e2e_slides)])
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = e2e_slides * world / (e2e_ms / 1e3)
    e2e_steps = e2e_slides / S

    if rank == 0:
        pk = peaks()
        fc1_ms = stages["fc1_gemm"] / max(calls, 1)
        fc1_tflops = FC1_FLOP_PER_PATCH * n / (fc1_ms * 1e-3) / 1e12 if fc1_ms > 0 else 0.0
        whole_ms = sum(stages_serial.values()) / max(calls_serial, 1)
        fc1_serial_ms = stages_serial["fc1_gemm"] / max(calls_serial, 1)
        fc1_serial_tflops = FC1_FLOP_PER_PATCH * n / (fc1_serial_ms * 1e-3) / 1e12 if fc1_serial_ms > 0 else 0.0
        line = {
            "metric": "slides_per_sec_n50k", "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3-split (fp32 accumulate; fp32-class accuracy)", "data": "synthetic",
            "config": {"workload": "TOAD_fc_mtl_concat forward (eval), N=%d x %d fp32, big, n_classes=18" % (n, WIDTH),
                       "slides_per_step": S, "slides_in_flight": args.streams if B == 1 else 1, "slides_per_call": B,
                       "parallelism": "one slide per GPU, replicas (no collective in eval)",
                       "l2_policy": "%d distinct %.0f MB bags per GPU rotated (inputs larger than the 126 MB L2)" % (
                           n_bags, n * WIDTH * 4 / 1e6)},
            "clocks": sampler.summary() if sampler else None,
            "e2e": {"value": e2e_value, "unit": "slides/s", "h2d_bytes_per_step": int(streamer.h2d_bytes / max(e2e_steps, 1e-9)),
                    "d2h_bytes_per_step": int(streamer.d2h_bytes / max(e2e_steps, 1e-9)), "slides": e2e_slides,
                    "note": "pinned host bags, double-buffered H2D overlapped with compute (toad_b200.pipeline.SlideStreamer)"},
            "value_single_stream": value_single_stream,   # the same K steps strictly serial on one stream
            # per forward / forward_batch call: 3 tcgen05 GEMMs + 1 pooling launch (+ 3 weight-split launches once)
            "gpu_launches": 4 * (S // B if B > 1 else S) * args.steps + 3,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16x3_kernel<512,A_F32,EPI_LINEAR,2> (fc1, 44% of FLOPs)",
                         "achieved": fc1_serial_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": fc1_serial_tflops / pk["bf16_tflops"], "peak_source": pk["src"] + " bf16 burst",
                         "executed_tflops": 3 * fc1_serial_tflops, "executed_frac": 3 * fc1_serial_tflops / pk["bf16_tflops"],
                         "note": "achieved = algorithmic fp32 FLOPs of fc1 / its average CUDA-event duration when the kernel has "
                                 "the GPU to itself (the single-stream pass behind value_single_stream); the kernel executes 3 "
                                 "bf16 tensor passes per algorithmic FLOP (split precision), so frac tops out at 1/3",
                         "traffic": 256.4e6 if n == N_PATCHES else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch, ncu --set full "
                                         "(profiles/r1n_fwd_full.txt: 207.0 MB + 49.4 MB; algorithmic 205 MB x + 102 MB h1 planes, "
                                         "part of which is still in L2 when the kernel ends)",
                         "stage_ms": {k: v / max(calls_serial, 1) for k, v in stages_serial.items()},
                         "in_flight": {"achieved": fc1_tflops, "frac": fc1_tflops / pk["bf16_tflops"],
                                       "stage_ms": {k: v / max(calls, 1) for k, v in stages.items()},
                                       "note": "the same launches inside the headline region, where the slides in flight share the "
                                               "SMs: each launch is stretched, the sum of both streams' work finishes sooner"},
                         "forward_hbm_gbs": BYTES_PER_PATCH * n / (whole_ms * 1e-3) / 1e9 if whole_ms > 0 else None,
                         "hbm_peak_gbs": pk["hbm_gbs"]},
        }
        # the HBM-bound kernel of the path (softmax over N, P.h pooling, heads): 2,056 B/patch algorithmic
        tail_ms = stages_serial["pool_tail"] / max(calls_serial, 1)
        tail_gbs = TAIL_BYTES_PER_PATCH * n / (tail_ms * 1e-3) / 1e9 if tail_ms > 0 else 0.0
        line["roofline_tail"] = {"bound": "hbm", "kernel": "pool_heads_kernel", "achieved": tail_gbs, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": tail_gbs / pk["hbm_gbs"],
                                 "note": "CUDA-event stage time (includes the launch gap after the gate GEMM); ~10 us of it is "
                                         "the serial two-level merge + heads after the streaming phase (DESIGN.md section 5)"}
        if B > 1:   # forward_batch has no per-stage events: the per-kernel times come from the serial pass only
            line["roofline"].pop("in_flight", None)
        if world == 1 and not args.no_eager_baseline:
            line["eager_gpu_baseline"] = eager_gpu_leg(dev, n)
        if world == 1 and not args.no_resnet:
            line["resnet50_baseline"] = resnet_leg(dev)
        if world == 1 and not args.no_cpu_baseline:
            rate, done, cores, total = cpu_reference_rate(n, 40, 1, budget_s=15.0)
            line["cpu_baseline"] = {"value": rate, "unit": "slides/s", "cores": cores, "kind": "port",
                                    "sample": "%d forwards of one %dx%d bag in %.1f s, torch CPU ops on %d threads (best of a "
                                              "thread sweep, %d cores; oracle port of models/model_toad.py)" % (
                                                  done, n, WIDTH, total, cores, os.cpu_count() or 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def eager_gpu_leg(dev, n):
    """Secondary bar: the reference's own ATen op sequence (oracle torch port) run
    eagerly on the SAME GPU -- cuBLAS fp32 SGEMMs + elementwise kernels, what the unmodified reference
    module does on a B200.  A baseline, not part of the product path."""
    import torch
    from oracle import toad_oracle as O
    from oracle import toad_oracle_torch as OT
    params = {k: v.to(dev) for k, v in OT.to_torch_params(O.make_params(0, "big", 18)).items()}
    x = torch.randn(n, WIDTH, device=dev)
    sex = torch.ones(1, device=dev)
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(3):
            OT.toad_forward(x, sex, params)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            OT.toad_forward(x, sex, params)
        e1.record()
        torch.cuda.synchronize()
        out["slides_per_s_tf32" if tf32 else "slides_per_s_fp32"] = 10 / (e0.elapsed_time(e1) / 1e3)
    torch.backends.cuda.matmul.allow_tf32 = False
    out["note"] = "torch eager on the same B200: fp32 = the reference's default (allow_tf32 off); tf32 = single-pass TF32, which fails the 1e-3 parity bar (SURVEY F4)"
    return out


def resnet_leg(dev):
    """Secondary number (config 3): resnet50_baseline feature extraction, synthetic 3x256x256 patches."""
    import torch
    from models.resnet_custom import resnet50_baseline
    torch.manual_seed(1)
    model = resnet50_baseline(pretrained=False).to(dev).eval()
    B = 256
    x = torch.randn(B, 3, 256, 256, device=dev)
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 4
        for _ in range(reps):
            model(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pps = B / ms * 1e3
    tf = 8.562671616e9 * pps / 1e12
    pk = peaks()
    return {"patches_per_s": pps, "batch": B, "ms_per_batch": ms, "algorithmic_tflops": tf,
            "executed_tflops": 3 * tf, "executed_frac_of_bf16_peak": 3 * tf / pk["bf16_tflops"],
            "note": "BN-folded implicit-GEMM convs on the split-bf16 tcgen05 kernel (3 tensor passes per FLOP), "
                    "8.563 GFLOP/patch (SURVEY.md R3); kaiming-init weights, eval mode"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-patches", type=int, default=N_PATCHES)
    ap.add_argument("--slides-per-step", type=int, default=16)
    ap.add_argument("--streams", type=int, default=3, help="slides in flight per GPU (toad_b200.pipeline.ResidentRunner)")
    ap.add_argument("--batch", type=int, default=16,
                    help="slides per forward_batch call (<= 16, must divide --slides-per-step); 1 = one forward per slide")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resnet", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager GPU leg (ncu launch lists)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
