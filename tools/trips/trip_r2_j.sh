#!/usr/bin/env bash
# round-2 trip J (1 GPU): rank after batch-predicated list scans + final raise -- parity tests, timings, launch list
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 420 python -m pytest tests/test_rank_tc_gpu.py tests/test_rank_gpu.py -q -x --timeout 200 ) > gpurun_out/pytest_rank.log 2>&1
echo "rank exit $?" >> gpurun_out/pytest_rank.log
export TUNE_ONLY=c5
rm -f gpurun_out/rank_j.log
for cfg in "1 2 0 100 100" "1 4 0 100 100" "2 2 0 100 100" "2 4 0 100 100" "2 4 0 10 100" "2 2 0 10 100"; do
  set -- $cfg
  echo "== CTA=$1 STRIPS=$2 DEBUG=$3 topk=$4 excl=$5" >> gpurun_out/rank_j.log
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 B200_RANK_DEBUG=$3 TUNE_TOPK=$4 TUNE_EXCL=$5 timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_j.log 2>&1
done
for v in "2 2" "2 4"; do
  set -- $v
  B200_RANK_CTA=$1 B200_RANK_STRIPS=$2 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches_j_cta$1_st$2.csv python tools/tune_rank.py > /dev/null 2>&1
done
tail -4 gpurun_out/pytest_rank.log; grep -E "^==|^rank" gpurun_out/rank_j.log; grep -E "rank_tc|finish" gpurun_out/rank_launches_j_cta2_st4.csv | tail -6 | cut -c1-200
