#!/usr/bin/env bash
# where does the BPR epoch time go?  (dev tool, GPU box)
export B200_TUNE_EXPERIMENT=1
for skip in 0 1 2 4 6 7; do
  echo "== B200_BPR_DEBUG_SKIP=$skip (1=U 2=V+ 4=V- scatter off)"; B200_BPR_DEBUG_SKIP=$skip python tools/tune_bpr.py --k 64 2>&1 | grep "G samples"
done
echo "== uniform item popularity"; python tools/tune_bpr.py --k 64 --uniform 1 2>&1 | grep "G samples"
echo "== uniform, no scatter"; B200_BPR_DEBUG_SKIP=7 python tools/tune_bpr.py --k 64 --uniform 1 2>&1 | grep "G samples"
