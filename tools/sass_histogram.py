#!/usr/bin/env python
"""Opcode histogram of the hot kernels in the built library (cuobjdump -sass on the per-size objects): the
committed evidence that they are sm_100a code using bulk async copies (UBLKCP), mbarriers (SYNCS.*),
packed FP32 (FADD2 / FMUL2 / FFMA2), 3-input min/max (FMNMX3), REDUX and no library calls.

    python tools/sass_histogram.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "mlx_audio_primitives_b200", "csrc", "build")
KERNELS = [  # (object, substring of the mangled kernel name, label)
    ("fwd_400.o", "mel_rows_kernel", "mel_rows_kernel<n_fft 400, POW_SQUARE, bank in smem> (C2, default)"),
    ("fwd_400.o", "mel_ws_kernel", "mel_ws_kernel<n_fft 400, POW_SQUARE, bank in smem> (C2, MLXA_MEL_WS=1)"),
    ("fwd_2048.o", "fwd_kernelILi1ELi0E", "fwd_kernel<EP_MEL, POW_SQUARE> n_fft 2048 (C3)"),
    ("fwd_4096.o", "fwd_kernelILi1ELi0E", "fwd_kernel<EP_MEL, POW_SQUARE> n_fft 4096 (C4)"),
    ("fwd_2048.o", "fwd_kernelILi0ELi0E", "fwd_kernel<EP_STFT> n_fft 2048 (C1)"),
    ("fwd_1024.o", "fwd_kernelILi2ELi0E", "fwd_kernel<EP_GL> n_fft 1024 (C5)"),
    ("inv_1024.o", "inv_kernelILb1E", "inv_kernel<full spectrum> n_fft 1024 (C5, C1 at 2048)"),
]
MARK = ["UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FMNMX3", "REDUX", "CREDUX", "LDGSTS", "MUFU", "BAR", "LDS", "STS",
        "SHFL", "HMMA", "UTCHMMA", "CALL"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            t = re.sub(r"/\*.*?\*/", "", line).split()
            if t:
                op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                body.append(op.rstrip(";"))
    if name:
        yield name, body


for obj, key, label in KERNELS:
    best = None
    for name, body in functions(obj):
        if key in name and (best is None or ("Li0ELb1" in name and "Li0ELb1" not in best[0])):
            best = (name, body)
    if best is None:
        print(f"== {label}: not found in {obj}")
        continue
    name, body = best
    full = collections.Counter(body)
    base = collections.Counter(op.split(".")[0] for op in body)
    print(f"== {label}\n   {name}\n   {len(body)} SASS instructions (static)")
    print("   marks: " + ", ".join(f"{m} {sum(v for k, v in full.items() if k.split('.')[0] == m)}" for m in MARK))
    print("   top:   " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)))
    sy = {k: v for k, v in full.items() if k.startswith("SYNCS") or k.startswith("UBLKCP")}
    if sy:
        print("   async: " + ", ".join(f"{k} {v}" for k, v in sorted(sy.items())))
