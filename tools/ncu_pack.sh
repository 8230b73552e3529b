#!/bin/bash
# tools/ncu_pack.sh <report.ncu-rep> : summarise an ncu report on the GPU box (per-kernel table + pickled source pages
# for tools/line_profile.py) and drop the report itself when it is too large to travel back (gpurun_out <= 64 MiB).
rep="$1"; base="${rep%.ncu-rep}"
python tools/ncu_summary.py "$rep" "$base.pkl" > "$base.summary.txt" 2>&1
sz=$(stat -c %s "$rep")
if [ "$sz" -gt 20000000 ]; then rm -f "$rep"; echo "dropped $rep ($sz bytes), kept summary + source pickle"; fi
ls -la "$base"*
