#!/usr/bin/env bash
# round-2 trip E (1 GPU): rank v4 (interleaved lists + one-pass raises + early hand-back), scheduled replay, eval paths, example experiment
mkdir -p gpurun_out
python -c "
import torch, sys
sys.path.insert(0, '.')
torch.zeros(1).cuda(); torch.cuda.synchronize()
from cornac_b200 import _lib; _lib.load(); print('warm ok')
" > gpurun_out/warm.log 2>&1
( time timeout -s KILL 1200 python -m pytest tests -q -m gpu --timeout 600 ) > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
export TUNE_ONLY=c5
for cta in 1 2; do
  for st in 2 4; do
    for dbg in 0 4; do
      echo "== B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg" >> gpurun_out/rank_v4.log
      B200_RANK_CTA=$cta B200_RANK_STRIPS=$st B200_RANK_DEBUG=$dbg timeout -s KILL 200 python tools/tune_rank.py >> gpurun_out/rank_v4.log 2>&1
    done
  done
done
for cta in 1 2; do
  B200_RANK_CTA=$cta B200_RANK_STRIPS=4 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rank_launches_v4_cta$cta.csv python tools/tune_rank.py > /dev/null 2>&1
done
unset TUNE_ONLY
timeout -s KILL 300 python tools/tune_replay.py > gpurun_out/tune_replay_r2.log 2>&1
( time timeout -s KILL 600 python examples/bpr_experiment.py ) > gpurun_out/example.log 2>&1
tail -12 gpurun_out/pytest.log; cat gpurun_out/rank_v4.log; cat gpurun_out/tune_replay_r2.log; tail -14 gpurun_out/example.log
