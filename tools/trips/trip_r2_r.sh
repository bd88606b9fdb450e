#!/usr/bin/env bash
# round-2 trip R (1 GPU): cache-blocked order -- bytes of rows per window / item block (B200_BPR_PART_MB) on the whole configs[2] model
mkdir -p gpurun_out
rm -f gpurun_out/bpr_part.log
for mb in 40 24 32 56 80 128; do
  echo "== B200_BPR_PART_MB=$mb" >> gpurun_out/bpr_part.log
  B200_BPR_PART_MB=$mb timeout -s KILL 400 python bench.py --steps 4 --warmup 3 --no-rank --no-mf --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.4g ms %.1f frac %.4f order %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['sample_order'][:60]))
" >> gpurun_out/bpr_part.log 2>&1
done
cat gpurun_out/bpr_part.log
