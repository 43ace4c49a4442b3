#!/bin/bash
# Bench lines for the BASELINE configs that are not the headline (configs 2/3/5) + ncu captures of the covariance K1 and the
# 20M-edge PCG pass.  usage (repo root, under gpurun): bash profiles/gpu_configs.sh TAG [workloads...]
TAG=${1:-x}; shift
WL=${@:-"terrace_like piccadilly_like syn_100k_20M_cov"}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt
nproc >> $O/gpu_$TAG.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_$TAG.txt
for W in $WL; do
  timeout 900 python bench.py --workload $W > $O/bench_${W}_$TAG.json 2> $O/bench_${W}_$TAG.err
  tail -c 1500 $O/bench_${W}_$TAG.json
done
if echo "$WL" | grep -q syn_100k_20M_cov; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_edges --launch-skip 2 -c 1 -f -o $O/full_k_edges_20M_$TAG \
    python profiles/kernel_times.py syn_100k_20M_cov 3 > $O/full_k_edges_20M_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg_persistent --launch-skip 2 -c 1 -f -o $O/full_k_pcg_20M_$TAG \
    python profiles/kernel_times.py syn_100k_20M_cov 8 > $O/full_k_pcg_20M_$TAG.log 2>&1
fi
