#!/bin/bash
# Profiling pass of a round (run under gpurun, ONE GPU): launch lists (metrics-only ncu) of the C2 bench step, the C4 Q* step and a
# small-batch KAT search, plus one `ncu --set full` capture of the GEMM launches of the last C2 iteration.  Outputs -> gpurun_out/<tag>_*.
tag=${1:-r2}
out=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-configs --no-hda --no-kat --no-cpu-baseline"
$B > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err || exit 1
for pe in 1 4 8; do echo "POLL_EVERY=$pe"; DXB200_POLL_EVERY=$pe python tools/latency_probe.py; done > $out/${tag}_latency.log 2>&1
python -m pytest tests/test_gpu_api.py tests/test_gpu_engine.py -m gpu -x -q > $out/${tag}_tests.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_c2_launches.csv $B > $out/${tag}_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_q_launches.csv python tools/config_probe.py cube3 0 q 10000 8 40 > $out/${tag}_ncu_q.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_kat_launches.csv python tools/latency_probe.py > $out/${tag}_ncu_kat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16_tc -s 320 -c 16 -o $out/${tag}_gemm_full -f $B > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_gemm_full.ncu-rep --page raw --csv > $out/${tag}_gemm_full_raw.csv 2>/dev/null
ls -la $out | tail -20
