#!/usr/bin/env bash
# GPU trip 32: joint (row-level) threshold raises
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 600 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
export TUNE_ONLY=c5
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_|scale_' -c 16 --csv --log-file gpurun_out/launches_rank_c5.csv python tools/tune_rank.py > /dev/null 2>&1
tail -8 gpurun_out/pytest.log; cat gpurun_out/tune_rank.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_rank_c5.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:9]:
    print(r[ki][:50], r[vi])
PY
