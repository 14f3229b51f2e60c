#!/bin/bash
N=${1:-2}
L=gpurun_out/ddp_graph_probe_n$N.log
: > $L
run() { echo "=== N=$N $*" >> $L
env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --mode train --gpus $N --steps 10 --warmup 3 2>gpurun_out/ddp_graph_err.log | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print(json.dumps({k: d.get(k) for k in ('n_gpus', 'ms_per_step', 'activation_kernels_ms_per_step', 'nccl_kernels_ms_per_step', 'all_kernels_ms_per_step', 'loss')} | {'graph': d['config'].get('cuda_graph_step'), 'ddp': d['config'].get('ddp')}))
" >> $L
grep -E "Error" gpurun_out/ddp_graph_err.log | tail -2 >> $L; }
run AFA_TRAIN_GRAPH=0 AFA_DDP_BROADCAST_BUFFERS=1
run AFA_TRAIN_GRAPH=1 AFA_DDP_BROADCAST_BUFFERS=1
run AFA_TRAIN_GRAPH=0 AFA_DDP_BROADCAST_BUFFERS=1 AFA_DDP_BUCKET_MB=25 AFA_DDP_STATIC=0
cat $L
