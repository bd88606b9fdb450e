#!/usr/bin/env bash
# GPU trip 8: rank kernel v2 (branch-free appends, 8 epilogue warps, merge exclusion)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 400 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
timeout -s KILL 600 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc' -c 14 --csv --log-file gpurun_out/launches_rank.csv python tools/tune_rank.py > gpurun_out/tune_rank_ncu.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_rank_tc3 python tools/tune_rank.py > gpurun_out/ncu_rank_full.log 2>&1
tail -15 gpurun_out/pytest_tc.log; cat gpurun_out/tune_rank.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank.csv')) if len(r)>14 and r[12]=='gpu__time_duration.sum']
for r in rows[:20]: print(r[4][:50], r[8], r[14])
PY
