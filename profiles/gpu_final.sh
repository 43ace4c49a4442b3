#!/bin/bash
# One GPU-box pass that leaves only SMALL files behind (gpurun copies back at most 64 MiB): bench lines of the four synthetic
# configs, the ncu launch list of the headline bench, and `ncu --set full` captures of K1 and the persistent PCG kernel at 1M and
# 20M edges -- each capture is summarised here (raw metrics, stall hot spots, dynamic opcode mix, DRAM bytes per CG step) and the
# .ncu-rep deleted.   usage (repo root, under gpurun): bash profiles/gpu_final.sh TAG [with20m]
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt
nproc >> $O/gpu_$TAG.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_$TAG.txt
timeout 900 python bench.py > $O/bench_syn_10k_1M_$TAG.json 2> $O/bench_syn_10k_1M_$TAG.err
for W in terrace_like piccadilly_like syn_100k_20M_cov; do
  timeout 900 python bench.py --workload $W > $O/bench_${W}_$TAG.json 2> $O/bench_${W}_$TAG.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-accuracy > $O/launches_$TAG.log 2>&1
python profiles/summarize_launches.py $O/launches_$TAG.csv > $O/launches_summary_$TAG.txt 2>&1
R=/tmp/ncu_$TAG; mkdir -p $R
cap() {  # name kernel-regex launch-skip workload repeats
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c 1 -f -o $R/$1 \
    python profiles/kernel_times.py $4 $5 > $O/ncu_$1_$TAG.log 2>&1
  python profiles/ncu_raw_summary.py $R/$1.ncu-rep >> $O/ncu_full_summary_$TAG.txt 2>&1
  python profiles/ncu_source_top.py $R/$1.ncu-rep $2 25 >> $O/stall_hotspots_$TAG.txt 2>&1
}
cap k_edges_1M k_edges 2 syn_10k_1M 3
python profiles/ncu_dynamic_mix.py $R/k_edges_1M.ncu-rep k_edges > $O/k_edges_dynamic_mix_$TAG.txt 2>&1
cap k_pcg_1M k_pcg_persistent 2 syn_10k_1M 50
if [ -n "$2" ]; then   # the 20M-edge captures (slow: the graph is generated three times); r02i's are of the same 6-double kernels
  cap k_edges_20M k_edges 2 syn_100k_20M_cov 3
  python profiles/ncu_dynamic_mix.py $R/k_edges_20M.ncu-rep k_edges >> $O/k_edges_dynamic_mix_$TAG.txt 2>&1
  cap k_pcg_20M k_pcg_persistent 2 syn_100k_20M_cov 8
  python profiles/ncu_traffic.py syn_100k_20M_cov:1 $R/k_pcg_20M.ncu-rep 8 >> $O/ncu_traffic_$TAG.log 2>&1
fi
cp profiles/ncu_traffic.json /tmp/ncu_traffic_before.json
python profiles/ncu_traffic.py syn_10k_1M:1 $R/k_pcg_1M.ncu-rep 50 >> $O/ncu_traffic_$TAG.log 2>&1
cp profiles/ncu_traffic.json $O/ncu_traffic_$TAG.json
rm -rf $R
python - <<XEOF
import json,glob
for f in sorted(glob.glob("$O/bench_*_$TAG.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, "value %.4e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "pcg/step", d["pcg_iterations_per_step"], "cg %.5f"%r["ms_per_launch"], "frac %.3f"%r["frac"], "k1 %.5f"%r["k1"]["ms_per_launch"], "e2e %.4e"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
XEOF
du -sh $O
