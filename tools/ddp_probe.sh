#!/bin/bash
# BASELINE config 5 under DDP on N GPUs with different bucket / graph settings
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/ddp_probe_n$N.log
: > $L
run() { echo "=== N=$N $*" >> $L; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --mode train --gpus $N --steps 10 --warmup 3 2>>gpurun_out/ddp_probe_err.log | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print(json.dumps({k: d.get(k) for k in ('n_gpus', 'ms_per_step', 'activation_kernels_ms_per_step', 'nccl_kernels_ms_per_step', 'all_kernels_ms_per_step', 'loss')} | {'ddp': d['config'].get('ddp')}))
" >> $L; }
echo "=== N=1" >> $L
timeout 300 python bench.py --mode train --steps 10 --warmup 3 2>>gpurun_out/ddp_probe_err.log | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print(json.dumps({k: d.get(k) for k in ('n_gpus', 'ms_per_step', 'activation_kernels_ms_per_step', 'all_kernels_ms_per_step', 'loss')}))
" >> $L
run AFA_DDP_BUCKET_MB=25 AFA_DDP_STATIC=0
run AFA_DDP_BUCKET_MB=100 AFA_DDP_STATIC=1
run AFA_DDP_BUCKET_MB=500 AFA_DDP_STATIC=1
run AFA_DDP_BUCKET_MB=100 AFA_DDP_STATIC=1 AFA_DDP_BF16_HOOK=1
run AFA_DDP_BUCKET_MB=100 AFA_DDP_STATIC=1 NCCL_MAX_NCHANNELS=4
run AFA_DDP_BUCKET_MB=100 AFA_DDP_STATIC=1 NCCL_ALGO=NVLS
cat $L
