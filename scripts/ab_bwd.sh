#!/bin/bash
# A/B the backward kernel knobs (ring depth, K-chunk staggering) on the bench workload
for ring in 3 6; do for st in 0 1; do
  echo -n "ring=$ring mcast=$st: "
  RCNN_BWD_RING=$ring RCNN_MCAST=$st python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels']['lstm_bwd'], d['kernels']['lstm_fwd'])"
done; done
