#!/bin/bash
# tools/r2_exp.sh "<ENV=... ENV=...>" ... : one short configs[2] bench per environment variant (tuning experiments)
out=gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python bench.py --steps 3 --warmup 2 --cpu-seconds 1 --unitigs 0 --secondary "" --e2e-steps ${E2E:-1} > $out/exp_$i.json 2> $out/exp_$i.err
  echo "== $v rc=$?" >> $out/exp.log
  python tools/show_bench.py $out/exp_$i.json 2>&1 | head -4 >> $out/exp.log
done
cat $out/exp.log
