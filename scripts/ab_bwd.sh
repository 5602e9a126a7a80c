#!/bin/bash
# A/B the backward kernel knobs on the bench workload
for two in 0 1; do for ring in 3 6; do
  echo -n "2sm=$two ring=$ring: "
  RCNN_BWD_2SM=$two RCNN_BWD_RING=$ring python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels']['lstm_bwd'], d['kernels']['lstm_fwd'])"
done; done
