#!/bin/bash
# Profiling pass of a round (run under gpurun, ONE GPU).  Outputs -> gpurun_out/<tag>_*; tools/profile_summary.py and
# tools/launch_list_summary.py turn them into the tracked summaries under profiles/.
#   1. plain runs first (a number printed under ncu is never a bench value)
#   2. launch lists (metrics-only ncu): the C2 bench step, the C4 Q* step, a small-batch KAT search (bf16)
#   3. ncu --set full: the hidden-layer tcgen05 GEMM launches of the last C2 iteration; the one-launch small network kernel;
#      the radix-sort / merge kernels of a late Q* iteration
tag=${1:-r2}
out=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-configs --no-hda --no-kat --no-cpu-baseline"
$B > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err || exit 1
python tools/latency_probe.py > $out/${tag}_latency.log 2>&1
python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_q_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_c2_launches.csv $B > $out/${tag}_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_q_launches.csv python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_ncu_q.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1400 --csv --log-file $out/${tag}_kat_launches.csv python tools/latency_probe.py bf16 > $out/${tag}_ncu_kat.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm_bf16_tc<\(int\)256, \(int\)2, \(int\)0, \(bool\)0, \(bool\)0>' -s 246 -c 6 -o $out/${tag}_gemm_full -f $B > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_gemm_full.ncu-rep --page raw --csv > $out/${tag}_gemm_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_mlp_small -s 300 -c 2 -o $out/${tag}_small_full -f python tools/latency_probe.py bf16 > $out/${tag}_ncu_small.log 2>&1
ncu -i $out/${tag}_small_full.ncu-rep --page raw --csv > $out/${tag}_small_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:k_small_tail_v|k_small_filter_v2' -s 600 -c 4 -o $out/${tag}_blocks_full -f python tools/latency_probe.py bf16 > $out/${tag}_ncu_blocks.log 2>&1
ncu -i $out/${tag}_blocks_full.ncu-rep --page raw --csv > $out/${tag}_blocks_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:k_radix_onesweep|k_merge|k_radix_prepass' -s 330 -c 12 -o $out/${tag}_qsort_full -f python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_ncu_qsort.log 2>&1
ncu -i $out/${tag}_qsort_full.ncu-rep --page raw --csv > $out/${tag}_qsort_full_raw.csv 2>/dev/null
ls -la $out | tail -30
