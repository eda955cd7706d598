#!/bin/bash
# small-batch (solve-time regime) profiling pass: live per-iteration latency, ncu launch list of a bf16 KAT search, and one
# `ncu --set full` capture of the one-launch network kernel.  Outputs -> gpurun_out/<tag>_*.
tag=${1:-r2s}
out=gpurun_out
python tools/latency_probe.py > $out/${tag}_latency.log 2>&1
for mt in 1 2; do echo "DXB200_SMALL_NET_MT=$mt"; DXB200_SMALL_NET_MT=$mt python tools/latency_probe.py fp32,bf16; done >> $out/${tag}_latency.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2000 --csv --log-file $out/${tag}_kat_launches.csv python tools/latency_probe.py bf16 > $out/${tag}_ncu_kat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mlp_small -s 300 -c 2 -o $out/${tag}_small_full -f python tools/latency_probe.py bf16 > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_small_full.ncu-rep --page raw --csv > $out/${tag}_small_full_raw.csv 2>/dev/null
ncu -i $out/${tag}_small_full.ncu-rep --page details > $out/${tag}_small_full_details.txt 2>/dev/null
cat $out/${tag}_latency.log
