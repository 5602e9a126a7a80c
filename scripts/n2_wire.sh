#!/bin/bash
# data-parallel step with fp32 and with bf16 gradients on the wire (RCNN_DP_WIRE=bf16), N = $1 GPUs
N=${1:-2}
for w in fp32 bf16; do
  RCNN_DP_WIRE=$w python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 30 --warmup 5 --no-extras --no-attention > gpurun_out/wire_${N}_$w.json 2> gpurun_out/wire.err
  python -c "
import json
d=json.loads(open('gpurun_out/wire_${N}_$w.json').read().strip().splitlines()[-1])
print('N=$N wire $w: ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'loss', d['loss_first_last'])"
done
