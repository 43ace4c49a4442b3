#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the dominant kernels.
# usage (from the repo root, under gpurun): bash profiles/gpu_round.sh TAG [skip-tests]
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt
nproc >> $O/gpu_$TAG.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_$TAG.txt
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log
fi
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/launches_$TAG.log 2>&1
for K in k_edges k_spmv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -f -o $O/full_${K}_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_${K}_$TAG.log 2>&1
done
# the persistent PCG kernel: launch #3 of kernel_times.py runs exactly 50 CG steps (rtol 0) -> DRAM bytes per CG step = total / 50
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pcg_persistent --launch-skip 2 -c 1 -f -o $O/full_k_pcg_persistent_$TAG \
  python profiles/kernel_times.py syn_10k_1M 50 > $O/full_k_pcg_persistent_$TAG.log 2>&1
tail -3 $O/pytest_$TAG.log; cat $O/bench_$TAG.json | head -c 3000
