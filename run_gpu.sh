#!/usr/bin/env bash
# GPU trip 18: base folded into the MMA (no bias ring) + MMMF kernels
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
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_rank_tc_gpu.py --deselect tests/test_rank_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_' -s 120 -c 6 --csv --log-file gpurun_out/launches_rank_c5.csv python tools/tune_rank.py > gpurun_out/tune_rank_ncu.log 2>&1
tail -12 gpurun_out/pytest_tc.log; cat gpurun_out/tune_rank.log; tail -5 gpurun_out/pytest.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank_c5.csv')) if len(r)>14 and r[12]=='gpu__time_duration.sum']
for r in rows[:6]: print(r[4][:50], r[8], r[14])
PY
