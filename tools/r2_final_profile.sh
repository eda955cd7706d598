#!/bin/bash
# Final profiling pass of round 2 (run under gpurun, ONE GPU) after the two-run OPEN for big batches and the sort key cut:
#   1. the full bench line without ncu (-> profiles/r2_bench_n1.json)
#   2. launch lists (metrics-only ncu) of the C2 bench step and the C4 Q* step
#   3. ncu --set full of the Q* sort / merge kernels of a late iteration
# The GEMM / small-iteration captures of tools/r2_profile.sh are unchanged by those commits and not repeated.
tag=${1:-r2f}
out=gpurun_out
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || { tail -5 $out/${tag}_bench_n1.err; exit 1; }
B="python bench.py --steps 1 --warmup 1 --no-configs --no-hda --no-kat --no-cpu-baseline"
python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_q_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_c2_launches.csv $B > $out/${tag}_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_q_launches.csv python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_ncu_q.log 2>&1
ncu --set full --clock-control none -k 'regex:k_radix_onesweep|k_merge|k_radix_prepass|k_sort_fixup|k_young' -s 620 -c 20 -o $out/${tag}_qsort_full -f python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_ncu_qsort.log 2>&1
ncu -i $out/${tag}_qsort_full.ncu-rep --page raw --csv > $out/${tag}_qsort_full_raw.csv 2>/dev/null
rm -f $out/${tag}_qsort_full.ncu-rep
ls -la $out | tail -12
