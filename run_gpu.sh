#!/usr/bin/env bash
# GPU trip 6: rank throughput + per-kernel breakdown, new tests (WBPR), ncu of the rank kernel
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 600 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
B200_RANK_TC=0 timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank_exact.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_|base_pad|score_|topk_' -c 60 --csv --log-file gpurun_out/launches_rank.csv python tools/tune_rank.py > gpurun_out/tune_rank_ncu.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_rank_tc python tools/tune_rank.py > gpurun_out/ncu_rank_full.log 2>&1
cat gpurun_out/tune_rank.log gpurun_out/tune_rank_exact.log; tail -6 gpurun_out/pytest.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank.csv')) if len(r)>14 and r[12]=='gpu__time_duration.sum']
for r in rows[:60]: print(r[4][:70], r[8], r[14])
PY
