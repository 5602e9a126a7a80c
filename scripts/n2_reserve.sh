#!/bin/bash
for cfg in "8 20" "16 20" "16 32" "4 8"; do set -- $cfg
  NCCL_MAX_CTAS=$1 RCNN_RESERVE_SMS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-attention > gpurun_out/n2f_$1_$2.json 2> gpurun_out/n2f.err
  python -c "
import json
d=json.loads(open('gpurun_out/n2f_$1_$2.json').read().strip().splitlines()[-1])
print('ctas $1 reserve $2 ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
NCCL_MAX_CTAS=16 RCNN_RESERVE_SMS=20 TRACE_ORDER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/step_trace.py > gpurun_out/n2_trace3.txt 2> gpurun_out/n2_trace3.err
