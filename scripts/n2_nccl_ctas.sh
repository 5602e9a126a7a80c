#!/bin/bash
# 2-GPU train step against NCCL's CTA budget: the recurrent kernels are cooperative launches of 128 CTAs (1 per SM)
for ctas in default 4 8 16; do
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-attention > gpurun_out/n2_ctas_$ctas.json 2> gpurun_out/n2_ctas_$ctas.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/n2_ctas_$ctas.json').read().strip().splitlines()[-1])
print("NCCL_MAX_CTAS=$ctas", "ms_per_step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
