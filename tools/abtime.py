"""A/B timing of toad_fwd stages for a given libtoad_b200.so build (ctypes only, ABI-tolerant).

    python tools/abtime.py path/to/lib.so [flags]
"""
import ctypes as C
import json
import subprocess
import sys

import torch


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "hid_dim", "attn_dim", "n_tasks", "n_classes")]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "w2", "b2", "wa", "ba", "wb", "bb", "wc", "bc", "wcls", "bcls", "wsite", "bsite")]


class FwdOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("a_raw", "features", "logits", "y_prob", "y_hat", "site_logits", "site_prob", "site_hat", "softmax_stats")]


def smi():
    try:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active",
                            "--format=csv,noheader"], capture_output=True, text=True, timeout=10)
        return r.stdout.strip()
    except Exception as e:
        return repr(e)


def main():
    lib = C.CDLL(sys.argv[1])
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
    dev = torch.device("cuda")
    torch.manual_seed(0)
    d = Dims(1024, 512, 384, 2, 18)
    shapes = [(512, 1024), (512,), (512, 512), (512,), (384, 512), (384,), (384, 512), (384,), (2, 384), (2,), (18, 513), (18,), (2, 513), (2,)]
    ps = [torch.randn(s, device=dev) * 0.03 for s in shapes]
    P = Params(*[p.data_ptr() for p in ps])
    x = torch.randn(n, 1024, device=dev)
    sex = torch.ones(1, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    outs = [torch.empty((2, n), **f32), torch.empty((2, 513), **f32), torch.empty(18, **f32), torch.empty(18, **f32),
            torch.empty(1, dtype=torch.int64, device=dev), torch.empty(2, **f32), torch.empty(2, **f32),
            torch.empty(1, dtype=torch.int64, device=dev), torch.empty(4, **f32)]
    O = FwdOut(*[o.data_ptr() for o in outs])
    nb = C.c_size_t()
    lib.toad_fwd_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int64, C.c_uint32, C.POINTER(C.c_size_t)]
    assert lib.toad_fwd_workspace_bytes(C.byref(d), n, flags, C.byref(nb)) == 0
    ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    lib.toad_fwd_profiled.argtypes = [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(FwdOut),
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]
    prof = C.c_void_p()
    lib.toad_profile_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.toad_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    assert lib.toad_profile_create(C.byref(prof), 64) == 0
    st = torch.cuda.current_stream().cuda_stream
    res = {"lib": sys.argv[1], "flags": flags, "smi_before": smi()}
    for rep in range(3):
        for _ in range(5):
            rc = lib.toad_fwd_profiled(C.byref(d), C.byref(P), x.data_ptr(), n, sex.data_ptr(), C.byref(O), None, wp, nb.value, flags, st, None)
            assert rc == 0, rc
        torch.cuda.synchronize()
        for _ in range(20):
            lib.toad_fwd_profiled(C.byref(d), C.byref(P), x.data_ptr(), n, sex.data_ptr(), C.byref(O), None, wp, nb.value, flags, st, prof)
        ms = (C.c_double * 5)()
        cnt = C.c_int32()
        lib.toad_profile_read(prof, ms, C.byref(cnt))
        res["rep%d" % rep] = [round(v / cnt.value, 4) for v in ms]
    res["smi_after"] = smi()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
