#!/usr/bin/env bash
# GPU trip 16: finish-kernel fix check + full test suite + bench (final state of the rank path)
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 600 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_|base_pad' -s 140 -c 7 --csv --log-file gpurun_out/launches_rank_c5.csv python tools/tune_rank.py > gpurun_out/tune_rank_ncu.log 2>&1
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bpr_|mf_|score_|topk_|delta_|rank_tc|pack_|norm_|base_pad' -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank_c5.csv')) if len(r)>14 and r[12]=='gpu__time_duration.sum']
for r in rows[:7]: print(r[4][:50], r[8], r[14])
PY
cat gpurun_out/bench.json
