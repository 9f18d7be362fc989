"""In-tree build of libmlxaudio_cuda.so (sm_100a only) with plain nvcc.

    python mlx_audio_primitives_b200/build.py [--force] [--verbose]

The forward / inverse transform kernels are compiled once per planned n_fft
(-DMLXA_NFFT=...) so the translation units build in parallel.  The result is
``mlx_audio_primitives_b200/_lib/libmlxaudio_cuda.so``; it is git-ignored but
travels to the GPU box with the working tree.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "cuda")
OBJ = os.path.join(PKG, "csrc", "build")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libmlxaudio_cuda.so")


def _sizes(macro: str) -> tuple[int, ...]:
    """the n_fft list of a MLXA_FOR_EACH_* macro in csrc/cuda/fft_sizes.cuh (the one list of planned sizes)"""
    import re
    text = open(os.path.join(SRC, "fft_sizes.cuh")).read().replace("\\\n", " ")
    m = re.search(r"#define\s+" + macro + r"\(X\)\s+(.*)", text)
    return tuple(int(v) for v in re.findall(r"X\((\d+)\)", m.group(1)))


PLANNED_NFFT = _sizes("MLXA_FOR_EACH_NFFT")
ACF_NFFT = _sizes("MLXA_FOR_EACH_ACF_NFFT")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "--expt-relaxed-constexpr",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _newest_header() -> float:
    hs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "mlxa_cuda.h"))
    hs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hs)


def _units():
    units = [("api.o", "api.cu", []), ("util_kernels.o", "util_kernels.cu", []), ("feat_kernels.o", "feat_kernels.cu", []),
             ("bigfft.o", "bigfft.cu", [])]
    for nf in PLANNED_NFFT:
        fwd_flags = [f"-DMLXA_NFFT={nf}"]
        if os.environ.get("MLXA_TWIDDLE_POWERS"):  # experiment: rebuild pass twiddles from three table entries
            fwd_flags.append("-DMLXA_TWIDDLE_POWERS")
        units.append((f"fwd_{nf}.o", "fwd_inst.cu", fwd_flags))
        inv_flags = [f"-DMLXA_NFFT={nf}"]
        if os.environ.get(f"MLXA_INV_THREADS_{nf}"):  # experiments: threads per CTA of the inverse kernel
            inv_flags.append(f"-DMLXA_INV_THREADS={os.environ[f'MLXA_INV_THREADS_{nf}']}")
        units.append((f"inv_{nf}.o", "inv_inst.cu", inv_flags))
        if nf in ACF_NFFT:  # the pitch kernels: power-of-two transform sizes
            units.append((f"acf_{nf}.o", "acf_inst.cu", [f"-DMLXA_NFFT={nf}"]))
    extra = os.environ.get("MLXA_EXTRA_FLAGS", "").split()  # experiments: e.g. MLXA_EXTRA_FLAGS=-DMLXA_NO_PAIRED_STORE
    return [(o, src, flags + extra) for o, src, flags in units]


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_time = _newest_header()
    jobs = []
    for obj, src, defs in _units():
        o, s = os.path.join(OBJ, obj), os.path.join(SRC, src)
        stale = force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time)
        if stale:
            cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((obj, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r.returncode, r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
            for name, rc, log in ex.map(run, jobs):
                if verbose and log.strip():
                    print(f"--- {name}\n{log}")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {name}:\n{log}")
    objs = [os.path.join(OBJ, u[0]) for u in _units()]
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
