#!/bin/bash
# N=2 check of the bench's column-sharded path (run with gpurun --gpus 2)
O=gpurun_out/e33; mkdir -p $O
nvidia-smi -L; nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -5 | tee $O/scale_n2.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -3 | cut -c1-300
python - <<'PY'
import torch
try:
    import torch.distributed._symmetric_memory as sm
    print("symm_mem ok", [n for n in dir(sm) if not n.startswith('_')][:40])
except Exception as e:
    print("symm_mem missing", e)
print(torch.cuda.can_device_access_peer(0,1))
PY
