"""Build libtoad_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libtoad_b200.so")
SOURCES = ["toad_abi.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "toad_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_id() -> str:
    """Content hash of everything the library is compiled from (sources, headers, flags): the binary carries it
    (toad_build_id), so staleness does not depend on file times -- which a snapshot copy to the GPU box rewrites."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def built_id(path: str = LIB):
    """toad_build_id() of the binary on disk, read without loading it: the id is stored as 'TOADID:<16 hex>'."""
    try:
        with open(path, "rb") as fh:
            data = fh.read()
    except OSError:
        return None
    i = data.find(b"TOADID:")
    return data[i + 7:i + 23].decode("ascii", "replace") if i >= 0 else None


def is_stale() -> bool:
    return built_id() != source_id()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:           # several ranks may arrive here at once (torchrun)
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():         # another process built it while we waited
                return LIB
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc_path()] + NVCC_FLAGS + ['-DTOAD_BUILD_ID="TOADID:%s"' % source_id(), "-o", tmp] + \
                [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB)                     # atomic: a concurrent dlopen sees the old or the new file
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
