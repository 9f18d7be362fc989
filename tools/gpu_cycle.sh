#!/bin/bash
# One GPU cycle: parity tests, the contract bench, the config rows, then an ncu --set full capture of one
# kernel (only after the plain run of the same command has exited 0).  Usage: tools/gpu_cycle.sh TAG [kernel-regex] [configs]
TAG=$1; KRE=${2:-mel_rows}; CFG=${3:-c2,c3,c4}
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_tests.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_configs.py --configs $CFG > gpurun_out/${TAG}_configs.jsonl 2>&1
tail -2 gpurun_out/${TAG}_tests.log
if [ "$KRE" != "none" ]; then
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o gpurun_out/${TAG}_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
fi
