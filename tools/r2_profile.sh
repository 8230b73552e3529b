#!/bin/bash
# tools/r2_profile.sh -- the ncu evidence of round 2, run on the GPU box AFTER the same bench command has exited 0 without ncu
# (B200_PROFILING.md): 1) launch list of the default bench command (configs[2]), 2) --set full capture of the hot kernels on the
# 40 Mb twin of configs[2] (same kernels and K; ncu's save/restore of a 30 GB state per replay pass does not finish in
# time on the full-size job), 3) per-kernel summaries + DRAM traffic per launch (profiles/r2_traffic.json).
# Every step runs under its own timeout; the big .ncu-rep never travels back.
set -x
out=gpurun_out
# (the same command without ncu is the bench run of this round: profiles/r2_bench_n1.json)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/r2_launches.csv \
    python bench.py --secondary ecoli --steps 1 --warmup 1 --cpu-seconds 1 --unitigs 0 --e2e-steps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bin_aggregate|merge_tiles|roll_scatter|split_bins|denoise_tiles|layout_scan|image_tiles|nl_write" -s 50 -c 16 \
    -o /tmp/r2_full python bench.py --workload celegans20 --secondary "" --steps 1 --warmup 0 --cpu-seconds 1 --unitigs 0 --e2e-steps 1 > $out/r2_full.log 2>&1
ls -la /tmp/r2_full.ncu-rep
timeout 300 python tools/ncu_summary.py /tmp/r2_full.ncu-rep $out/r2_full.pkl --traffic $out/r2_traffic.json --workload celegans20 > $out/r2_full.summary.txt 2>&1
for i in 0 2 4 6 8 10 12 14 16 18 20 22 24 26 28 30; do timeout 60 python tools/line_profile.py $out/r2_full.pkl $i sh-assembly_b200/libshq.so 16 >> $out/r2_line_profiles.txt 2>&1; done
cuobjdump -sass sh-assembly_b200/libshq.so | grep -E "Function :|UBLKCP|SYNCS|UTMALDG" | grep -B1 -E "UBLKCP|SYNCS|UTMALDG" | head -60 > $out/r2_sass_tma.txt
du -sh $out; ls -la $out | tail -12
find $out -size +25M -print -delete
