"""In-tree build recipes (nvcc / gcc).  Outputs are git-ignored but travel to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cadr_b200", "csrc")
LIBDIR = os.path.join(ROOT, "cadr_b200", "lib")
ORACLE = os.path.join(ROOT, "oracle")
REFERENCE = "/root/reference"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--cudart", "static", "-shared",
]
CUDA_SOURCES = ["capi.cu", "process_drawables.cu", "cull_compact.cu", "cull_bounds.cu", "upload.cu", "exchange.cu", "consume_check.cu", "external.cu"]
# A/B build only (libcadr_b200_exp.so, -DCADR_B200_EXPERIMENTS): earlier / alternative long-list kernels selectable through
# CADR_B200_CULL_VARIANT, evaluation stub CADR_B200_DIAG_NOEVAL.  Never part of libcadr_b200.so.
EXPERIMENT_SOURCES = ["experiments/cull_variants.cu"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError(f"build step failed: {cmd[0]} (exit {r.returncode})")
    return r.stdout


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """libcadr_b200.so: every CUDA kernel + the C ABI, sm_100a only.  experiments=True builds the A/B library
    libcadr_b200_exp.so instead (same ABI + the experiment kernels; used by scripts/, never by tests or bench.py)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libcadr_b200_exp.so" if experiments else "libcadr_b200.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + (EXPERIMENT_SOURCES if experiments else [])]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "cull_common.cuh"), os.path.join(ROOT, "include", "cadr_b200.h")]
    if experiments:
        deps += [os.path.join(CSRC, "experiments", h) for h in ("small_staged.cuh", "list_kernels.cuh")]
    if force or _stale(out, deps):
        flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DCADR_B200_EXPERIMENTS"] if experiments else [])
        log = _run([nvcc()] + flags + ["-o", out] + srcs)
        if verbose:
            print(log)
    return out


def build_oracle(force: bool = False) -> str:
    """oracle/liboracle.so: the CPU restatement (test infrastructure; never loaded by the product)."""
    out = os.path.join(ORACLE, "libcadr_oracle.so")
    src = os.path.join(ORACLE, "cadr_oracle.c")
    if force or _stale(out, [src]):
        _run(["gcc", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c11",
              "-Wall", "-Wextra", "-o", out, src, "-lm"])
    return out


def build_ref(force: bool = False) -> str | None:
    """oracle/_ref/: artefacts compiled from the reference's own sources where they lie (only in the build
    container; the GPU box uses the prebuilt files)."""
    mk = os.path.join(ORACLE, "Makefile")
    if not os.path.isdir(REFERENCE) or not os.path.exists(mk):
        return None
    _run(["make", "-C", ORACLE, "ref"] + (["-B"] if force else []))
    return os.path.join(ORACLE, "_ref")


def build_host(force: bool = False) -> str | None:
    """C++ facade + its test programs (cadr_b200/host)."""
    mk = os.path.join(ROOT, "cadr_b200", "host", "Makefile")
    if not os.path.exists(mk):
        return None
    _run(["make", "-C", os.path.dirname(mk)] + (["-B"] if force else []))
    return os.path.dirname(mk)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_oracle(force)
    build_ref(force)
    build_host(force)


if __name__ == "__main__":
    if "--experiments" in sys.argv:
        print("built:", build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments=True))
        sys.exit(0)
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", os.path.join(LIBDIR, "libcadr_b200.so"))
