#!/usr/bin/env bash
# GPU trip 9: rank kernel v3 (V ring released by the tensor pipe, bias ring, batched scans) + full tests + bench
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
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rank_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_rank_tc4 python tools/tune_rank.py > gpurun_out/ncu_rank_full.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_rank_tc_gpu.py --deselect tests/test_rank_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_tc.log; cat gpurun_out/tune_rank.log; tail -5 gpurun_out/pytest.log; cat gpurun_out/bench.json | cut -c1-3000; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank.csv')) if len(r)>14 and r[12]=='gpu__time_duration.sum']
for r in rows[:12]: print(r[4][:50], r[8], r[14])
PY
