#!/usr/bin/env python
"""bench.py -- slides/s of the TOAD attention-MIL forward at N=50k x 1024 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (1 process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One step = one pass of the hot path over `slides_per_step` synthetic slides (N x 1024 fp32 each), one
`model(data, sex)` call per slide on one stream -- the call the reference's train_loop / summary make.  `value` =
whole-job slides/s with bags resident in HBM (4 distinct 205 MB bags per GPU, larger than L2, rotated);
`value_in_flight` / `value_forward_batch` = the same work with several slides in flight / per call (APIs beyond the
reference's); `e2e` = the same metric through the public API with pinned host bags copied H2D inside the timed
region and results read back (+ the H2D-only ceiling); `train` = config 4, the fused training step with the NCCL
gradient all-reduce on N GPUs; `resnet50_baseline` = config 3 next to eager cuDNN.  Prints ONE JSON line.
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


def _max_over_ranks(ms, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(v, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_ours(args):
    import torch
    import torch.distributed as dist
    from toad_b200 import ops
    from toad_b200.pipeline import ResidentRunner, SlideStreamer
    from models.model_toad import TOAD_fc_mtl_concat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_log = None
    if world > 1:
        # stdout carries exactly one JSON line.  NCCL's own log (communicator / rank / transport lines: the evidence that
        # N ranks really formed one communicator) is NOT muted: every rank logs at INFO into its own file, and replays the
        # file on stderr when it is done (NCCL's default sink is stdout, where it would break the one-line contract).
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        nccl_log = "/tmp/toad_bench_nccl_rank%d_%d.log" % (rank, os.getpid())
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
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

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(step_fn, steps, warmup):
        """W warm-up steps, then exactly `steps` steps between barrier + synchronize, device-timed, max over ranks."""
        for i in range(warmup):
            step_fn(i)
        barrier()
        ev0.record()
        for i in range(steps):
            step_fn(i)
        ev1.record()
        barrier()
        return _max_over_ranks(ev0.elapsed_time(ev1), dev, world)

    # ---- value: the reference's own surface -- one `model(data, sex)` call per slide (core_utils_mtl_concat.py:206,
    # eval_utils_mtl_concat.py:91), strictly serial on the current stream, bags resident in HBM
    def step_forward(i):
        with torch.no_grad():
            for s in range(S):
                model(bags[(i * S + s) % n_bags], sex)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                       # NVML init happens during the warm-up, sampling every 2 ms
    for i in range(args.warmup):
        step_forward(i)
    barrier()
    if sampler:
        sampler.t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step_forward(i)
    ev1.record()
    barrier()
    if sampler:
        sampler.t1 = time.perf_counter()
        sampler.stop_flag.set()
    elapsed_ms = _max_over_ranks(ev0.elapsed_time(ev1), dev, world)
    total_slides = args.steps * S * world
    value = total_slides / (elapsed_ms / 1e3)

    # ---- per-kernel CUDA-event stage times of the same forwards (a separate, shorter pass: the stage events sit
    # between the kernels, so they are kept out of the headline region)
    prof_calls = min(args.steps * S, 256)
    prof = ops.Profile(prof_calls)
    model._prof = prof.handle
    with torch.no_grad():
        for c in range(prof_calls):
            model(bags[c % n_bags], sex)
    torch.cuda.synchronize()
    model._prof = None
    stages_serial, calls_serial = prof.read()
    prof.close()

    # ---- extras beyond the reference's surface: slides in flight on several streams, and several bags per call
    extra_steps = max(1, args.steps // 4)
    runner = ResidentRunner(model, n_streams=args.streams, device=dev)
    ms = timed(lambda i: runner.run([bags[(i * S + s) % n_bags] for s in range(S)], [sex] * S), extra_steps, 1)
    value_in_flight = extra_steps * S * world / (ms / 1e3)
    B = max(1, min(args.batch, 16, S))
    while S % B:
        B -= 1
    cats = [torch.cat([bags[(j + b) % n_bags] for b in range(B)], 0) for j in range(2)]
    sexes_b = torch.ones(B, device=dev)

    def step_batch(i):
        for c in range(S // B):
            model.forward_batch(cats[(i + c) % 2], [n] * B, sexes_b)
    ms = timed(step_batch, extra_steps, 1)
    value_forward_batch = extra_steps * S * world / (ms / 1e3)
    del cats

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
    e2e_ms = _max_over_ranks(ev0.elapsed_time(ev1), dev, world)
    e2e_value = e2e_slides * world / (e2e_ms / 1e3)
    e2e_steps = e2e_slides / S
    h2d_per_step, d2h_per_step = int(streamer.h2d_bytes / max(e2e_steps, 1e-9)), int(streamer.d2h_bytes / max(e2e_steps, 1e-9))
    # the ceiling of that number: the same pinned -> device copies alone, every rank at once
    copy_reps = 8
    dst = bags[0]
    for _ in range(2):
        dst.copy_(host_bags[0], non_blocking=True)
    barrier()
    ev0.record()
    for i in range(copy_reps):
        dst.copy_(host_bags[i % 2], non_blocking=True)
    ev1.record()
    barrier()
    h2d_ms = _max_over_ranks(ev0.elapsed_time(ev1), dev, world)
    h2d_gbs = copy_reps * world * n * WIDTH * 4 / (h2d_ms / 1e3) / 1e9
    del host_bags, streamer

    train = None if args.no_train else train_leg(args, model, dev, world, rank, barrier)

    if rank == 0:
        pk = peaks()
        whole_ms = sum(stages_serial.values()) / max(calls_serial, 1)
        fc1_serial_ms = stages_serial["fc1_gemm"] / max(calls_serial, 1)
        fc1_serial_tflops = FC1_FLOP_PER_PATCH * n / (fc1_serial_ms * 1e-3) / 1e12 if fc1_serial_ms > 0 else 0.0
        traffic = None if (args.no_traffic or world > 1) else measure_fc1_traffic(n)
        line = {
            "metric": "slides_per_sec_n50k", "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3-split (fp32 accumulate; fp32-class accuracy)", "data": "synthetic",
            "config": {"workload": "TOAD_fc_mtl_concat forward (eval), N=%d x %d fp32, big, n_classes=18" % (n, WIDTH),
                       "slides_per_step": S, "slides_per_call": 1, "slides_in_flight": 1,
                       "api": "model(data, sex) per slide on one stream -- the call the reference's train_loop / summary make",
                       "parallelism": "one slide per GPU, replicas (no collective in eval)",
                       "l2_policy": "%d distinct %.0f MB bags per GPU rotated (inputs larger than the 126 MB L2)" % (
                           n_bags, n * WIDTH * 4 / 1e6)},
            "clocks": sampler.summary() if sampler else None,
            "e2e": {"value": e2e_value, "unit": "slides/s", "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step, "slides": e2e_slides,
                    "h2d_only": {"gb_per_s": h2d_gbs, "slides_per_s_ceiling": h2d_gbs * 1e9 / (n * WIDTH * 4),
                                 "note": "the same pinned-host -> device copies with no compute, all %d rank(s) at once: "
                                         "the host-link ceiling of the e2e number" % world},
                    "note": "pinned host bags, double-buffered H2D overlapped with compute (toad_b200.pipeline.SlideStreamer)"},
            # extras beyond the reference's API (not the headline): several slides in flight / per call
            "value_in_flight": value_in_flight, "value_forward_batch": value_forward_batch,
            "extras_note": "value_in_flight = the same per-slide forwards with %d slides in flight on %d CUDA streams "
                           "(ResidentRunner); value_forward_batch = %d bags per forward_batch call" % (args.streams, args.streams, B),
            # per forward: 3 tcgen05 GEMMs + 1 pooling launch (+ 3 weight-split launches once)
            "gpu_launches": 4 * S * args.steps + 3,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16x3_kernel<512,A_F32,EPI_LINEAR,2> (fc1, 44% of FLOPs)",
                         "achieved": fc1_serial_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": fc1_serial_tflops / pk["bf16_tflops"], "peak_source": pk["src"] + " bf16 burst",
                         "executed_tflops": 3 * fc1_serial_tflops, "executed_frac": 3 * fc1_serial_tflops / pk["bf16_tflops"],
                         "note": "achieved = algorithmic fp32 FLOPs of fc1 / its average CUDA-event duration over %d serial "
                                 "forwards; the kernel executes 3 bf16 tensor passes per algorithmic FLOP (split precision), "
                                 "so frac tops out at 1/3" % calls_serial,
                         "traffic": traffic["bytes"] if traffic else None,
                         "traffic_note": (traffic["note"] if traffic else
                                          "not measured in this run (ncu unavailable or --no-traffic); see profiles/ for the "
                                          "committed capture") + "; algorithmic: %.1f MB of x + %.1f MB of h1 planes" % (
                                              n * 4096 / 1e6, n * 2048 / 1e6),
                         "stage_ms": {k: v / max(calls_serial, 1) for k, v in stages_serial.items()},
                         "forward_hbm_gbs": BYTES_PER_PATCH * n / (whole_ms * 1e-3) / 1e9 if whole_ms > 0 else None,
                         "hbm_peak_gbs": pk["hbm_gbs"]},
        }
        # the HBM-bound kernel of the path (softmax over N, P.h pooling, heads): 2,056 B/patch algorithmic
        tail_ms = stages_serial["pool_tail"] / max(calls_serial, 1)
        tail_gbs = TAIL_BYTES_PER_PATCH * n / (tail_ms * 1e-3) / 1e9 if tail_ms > 0 else 0.0
        line["roofline_tail"] = {"bound": "hbm", "kernel": "pool_heads_kernel", "achieved": tail_gbs, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": tail_gbs / pk["hbm_gbs"],
                                 "note": "CUDA-event stage time (includes the launch gap after the gate GEMM)"}
        if train is not None:
            line["train"] = train
        if world == 1 and not args.no_eager_baseline:
            line["eager_gpu_baseline"] = eager_gpu_leg(dev, n)
        if world == 1 and not args.no_resnet:
            line["resnet50_baseline"] = resnet_leg(dev, args)
        if world == 1 and not args.no_cpu_baseline:
            rate, done, cores, total = cpu_reference_rate(n, 40, 1, budget_s=15.0)
            line["cpu_baseline"] = {"value": rate, "unit": "slides/s", "cores": cores, "kind": "port",
                                    "sample": "%d forwards of one %dx%d bag in %.1f s, torch CPU ops on %d threads (best of a "
                                              "thread sweep, %d cores; oracle port of models/model_toad.py)" % (
                                                  done, n, WIDTH, total, cores, os.cpu_count() or 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
        try:
            with open(nccl_log) as f:
                sys.stderr.write(f.read())
            sys.stderr.flush()
            os.remove(nccl_log)
        except OSError:
            pass


def train_leg(args, model, dev, world, rank, barrier):
    """Config 4: the reference's training step (core_utils_mtl_concat.py:198-234: forward, 0.75 CE +
    0.25 CE, backward, Adam) on synthetic slides of N ~ U[5k, 80k] patches, one slide per GPU per step, the flat fp32
    gradient (4.77 MB) all-reduced over NCCL before the (identical) Adam step on every rank.  64 slides per GPU (weak
    scaling: 512 slides at 8 GPUs = the config), length-bucketed rounds (toad_b200.distributed.aligned_rounds)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from toad_b200.distributed import aligned_rounds
    from toad_b200.train import FusedTrainStep
    per_gpu = args.train_slides
    rng = np.random.default_rng(7)
    total = per_gpu * world
    lengths = rng.integers(5000, 80001, size=total).tolist()
    labels = rng.integers(0, 18, size=total).tolist()
    sites = rng.integers(0, 2, size=total).tolist()
    sexes = rng.integers(0, 2, size=total).tolist()
    rounds = aligned_rounds(lengths, world, seed=0)
    max_n = max(lengths)
    gen = torch.Generator(device=dev).manual_seed(500 + rank)
    pool = [torch.randn(max_n, WIDTH, generator=gen, device=dev) for _ in range(3)]   # 3 x 328 MB > L2; slides = row prefixes
    lab_t = [torch.tensor([c], device=dev) for c in range(18)]
    bin_t = [torch.tensor([c], device=dev) for c in range(2)]
    sex_t = [torch.tensor([float(c)], device=dev) for c in range(2)]
    torch.manual_seed(0)
    for p in model.parameters():            # identical start on every rank (the eval legs never changed them, but be explicit)
        if world > 1:
            dist.broadcast(p.data, 0)
    model.train()
    fused = FusedTrainStep(model, lr=1e-4, weight_decay=1e-5, max_patches=max_n)

    def run_round(s, r):
        i = r[rank]
        real = sum(1 for j in r if j >= 0)
        if i < 0:
            fused.step_idle(real)
            return 0
        x = pool[s % 3][:lengths[i]]
        fused.step(x, lab_t[labels[i]], bin_t[sites[i]], sex_t[sexes[i]], n_slides_in_round=real)
        return lengths[i]

    # warm-up: >= 5 steps, the first on the largest slide (sizes every workspace; first NCCL collective)
    big = max(range(total), key=lambda j: lengths[j])
    fused.step(pool[0][:lengths[big]], lab_t[0], bin_t[0], sex_t[0])
    for s in range(5):
        run_round(s, rounds[s % len(rounds)])
    barrier()
    nst = len(rounds)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nst)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    patches = 0
    e0.record()
    for s, r in enumerate(rounds):
        fused.timing_events = evs[s]
        patches += run_round(s, r)
    e1.record()
    barrier()
    fused.timing_events = None
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world)
    comp = sum(e[0].elapsed_time(e[1]) for e in evs) / nst
    allr = sum(e[1].elapsed_time(e[2]) for e in evs) / nst
    adam = sum(e[2].elapsed_time(e[3]) for e in evs) / nst
    comp_max = _max_over_ranks(comp, dev, world)
    allr_min = -_max_over_ranks(-allr, dev, world)      # the rank that waits least ~ the collective itself
    patches_all = _sum_over_ranks(float(patches), dev, world)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    same = 1.0
    if world > 1:
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = float(torch.equal(ref, flat))
    same_all = _sum_over_ranks(same, dev, world) == world
    # the same slides of THIS rank with no collective: what N independent GPUs would do (efficiency denominator)
    fused.allreduce = False
    for s in range(3):
        run_round(s, rounds[s])
    torch.cuda.synchronize()
    e0.record()
    mine = 0
    for s, r in enumerate(rounds):
        if r[rank] >= 0:
            mine += 1
            run_round(s, r)
    e1.record()
    torch.cuda.synchronize()
    solo_rate = mine / (e0.elapsed_time(e1) / 1e3)
    fused.allreduce = True
    ideal = _sum_over_ranks(solo_rate, dev, world)
    model.eval()
    value = total / (ms / 1e3)
    return {"metric": "train_slides_per_sec_config4", "value": value, "unit": "slides/s", "patches_per_s": patches_all / (ms / 1e3),
            "steps": nst, "slides": total, "ms_per_step": ms / nst,
            "split_ms": {"fwd_loss_bwd_max_rank": comp_max, "allreduce_min_rank": allr_min, "allreduce_mean_this_rank": allr,
                         "adam": adam},
            "independent_gpus_slides_per_s": ideal, "efficiency_vs_independent_gpus": value / ideal if ideal > 0 else None,
            "params_identical_across_ranks": bool(same_all),
            "config": {"workload": "config 4: FusedTrainStep (fwd + 0.75/0.25 CE + bwd + Adam lr 1e-4 wd 1e-5), %d slides per GPU, "
                                   "N ~ U[5000, 80000] seed 7, mean %.0f patches" % (per_gpu, sum(lengths) / total),
                       "collective": "1 NCCL all-reduce of the flat fp32 gradient (4,769,960 B) per step" if world > 1 else "none (1 GPU)",
                       "schedule": "length-bucketed synchronous rounds (aligned_rounds), warm-up 6 steps, device-timed, max over ranks"}}


def measure_fc1_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fc1 kernel, per launch, measured NOW on this box: one short
    forward-only workload (tools/profile_fwd.py) re-run under `ncu --metrics` (2 counters, 1 replay pass)."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:gemm_bf16x3", "--csv", sys.executable, os.path.join(ROOT, "tools", "profile_fwd.py"), "--n", str(n),
           "--iters", "3"]
    try:
        # own process group: a hung profiler run is killed together with the workload it spawned
        import signal
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, start_new_session=True,
                                env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
        try:
            stdout, _ = proc.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, signal.SIGKILL)
            proc.communicate()
            return None
        rows = list(csv.reader(l for l in stdout.splitlines() if l.startswith('"')))
        hdr = rows[0]
        ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        per = {}
        for row in rows[1:]:
            if "<512, 0, 0, 2" in row[ki].replace("(int)", ""):
                per.setdefault(row[ii], {})[row[mi]] = float(row[vi].replace(",", ""))
        vals = [d for d in per.values() if len(d) == 2]
        if not vals:
            return None
        last = vals[-1]                       # the last launch: warm instruction cache, cold data (the bag exceeds L2)
        rd, wr = last["dram__bytes_read.sum"], last["dram__bytes_write.sum"]
        return {"bytes": rd + wr, "note": "measured in this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on the fc1 "
                                          "launch of tools/profile_fwd.py --n %d (read %.1f MB + written %.1f MB)" % (n, rd / 1e6, wr / 1e6)}
    except Exception:
        return None


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


def resnet_leg(dev, args):
    """Config 3: resnet50_baseline feature extraction, synthetic 3x256x256 patches, batch 512, both arithmetic modes,
    next to the same network in torch eager on this GPU (cuDNN fp32 and cuDNN + TF32, torch's default for
    convolutions = what the unmodified reference runs on a GPU)."""
    import numpy as np
    import torch
    from models.resnet_custom import resnet50_baseline
    from oracle import resnet_oracle as RO
    params = RO.make_params(1)
    sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in params.items()}
    B = args.resnet_batch
    x = torch.randn(B, 3, 256, 256, device=dev)
    pk = peaks()
    out = {"batch": B, "flop_per_patch": 8562671616, "weights": "kaiming-normal convs, randomised BatchNorm statistics (oracle.make_params(1)), eval mode"}
    feats = {}

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.no_grad():
        for prec, passes in (("f16x2", 2), ("bf16x3", 3)):
            model = resnet50_baseline(pretrained=False)
            model.load_state_dict(sd, strict=True)
            model = model.to(dev).eval()
            model.precision = prec
            ms = timed(lambda: model(x), 6)
            pps = B / ms * 1e3
            tf = 8.562671616e9 * pps / 1e12
            out[prec] = {"patches_per_s": pps, "ms_per_batch": ms, "algorithmic_tflops": tf, "tensor_passes": passes,
                         "executed_tflops": passes * tf,
                         "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                      "frac": tf / pk["bf16_tflops_sustained"],
                                      "executed_frac": passes * tf / pk["bf16_tflops_sustained"],
                                      "peak_source": pk["src"] + " bf16 sustained (a whole-network step)"}}
            feats[prec] = model(x[:8].contiguous())
            del model
        if not args.no_eager_baseline:
            dp = {k: (v.to(dev) if v.dim() else v) for k, v in sd.items()}
            nb = min(B, 128)
            for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True
                ms = timed(lambda: RO.resnet50_baseline_forward(x[:nb], dp), 3)
                out["eager_" + name] = {"patches_per_s": nb / ms * 1e3, "batch": nb}
                feats[name] = RO.resnet50_baseline_forward(x[:8].contiguous(), dp)
            torch.backends.cudnn.allow_tf32 = True
            out["eager_note"] = ("the reference network (oracle = its own torch calls) in eager mode on this GPU; cudnn_tf32 is "
                                 "torch's default for convolutions, i.e. what the unmodified reference computes on a GPU")
        ref = RO.resnet50_baseline_forward(x[:8].cpu().double(), params)         # fp64 reference arithmetic, 8 patches
        scale = float(ref.abs().max())
        out["max_err_over_feature_scale_vs_fp64"] = {k: float((v.cpu().double() - ref).abs().max()) / scale for k, v in feats.items()}
    out["patches_per_s"] = out["f16x2"]["patches_per_s"]
    out["note"] = ("default mode f16x2: fp16 activation planes between layers, fp16 (hi, lo) weights, 2 tensor passes, fp32 "
                   "accumulation; bf16x3: (hi, lo) bf16 planes, 3 passes.  DRAM bytes per patch: profiles/ ncu launch lists")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-patches", type=int, default=N_PATCHES)
    ap.add_argument("--slides-per-step", type=int, default=256,
                    help="slides per step (20 steps x 256 slides = a timed region well above 1 s)")
    ap.add_argument("--train-slides", type=int, default=64, help="config-4 training leg: slides per GPU")
    ap.add_argument("--resnet-batch", type=int, default=512)
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu sub-run that measures fc1's DRAM bytes")
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
