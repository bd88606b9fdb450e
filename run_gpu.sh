#!/usr/bin/env bash
# GPU trip 40: full validation of HEAD + records for profiles/
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1500 python -m pytest tests -q -m gpu ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout -s KILL 300 python tools/tune_rank.py > gpurun_out/tune_rank.log 2>&1
TUNE_ONLY=c5 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rank_tc|pack_|norm_|scale_' -c 16 --csv --log-file gpurun_out/launches_rank_c5.csv python tools/tune_rank.py > /dev/null 2>&1
timeout -s KILL 600 python bench.py --workload c3shard --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3shard.json 2> gpurun_out/bench_c3shard.err
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/tune_rank.log
python - <<'PY'
import csv, json
rows=[r for r in csv.reader(open('gpurun_out/launches_rank_c5.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:9]:
    print(r[ki][:50], r[vi])
d=json.load(open('gpurun_out/bench_c3shard.json')); print(d['value'], d['roofline']['frac'], d['rank']['value'], d['rank']['tflops'], d['mf']['value'])
PY
