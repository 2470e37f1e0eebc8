"""Build libmarlc.so (hand-written sm_100a CUDA, C ABI) in-tree with nvcc."""
from __future__ import annotations

import fcntl
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarlc.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "0",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > os.path.getmtime(LIB) for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libmarlc.so.  nvcc cross-compiles
    without a GPU.  Raises if nvcc is missing or compilation fails."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmarlc.so")
    # Under torchrun every rank of a fresh checkout gets here at once: one builds, the others wait on the
    # lock and then find the library up to date.  The compiler writes to a per-process name and the
    # result is moved into place atomically, so no rank can dlopen a half-written file.
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("MARLC_NVCC_EXTRA", "").split(), "-o", tmp, *sources()]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, LIB)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
