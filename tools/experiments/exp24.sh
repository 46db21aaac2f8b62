#!/bin/bash
O=gpurun_out/e24; mkdir -p $O
for sh in 4096x4096 4096x11008; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_awq_lean -s 3 -c 1 -o $O/ncu_lean_$sh -f python tools/microbench.py --m 1 --iters 1 --layouts GEMM --shapes $sh > $O/ncu_lean_$sh.log 2>&1
done
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 or prefill" 2>&1 | tail -3
echo "== microbench gemm (graph)"; timeout 600 python tools/microbench.py --m 128,512,1024,2048 --layouts GPTQ --iters 50 --graph 2>&1 | tee $O/mb_gemm_graph.log | cut -c1-170
echo "== microbench gemm (eager)"; timeout 600 python tools/microbench.py --m 512,1024,2048 --layouts GPTQ --iters 50 --shapes 4096x4096 2>&1 | tee $O/mb_gemm_eager.log | cut -c1-170
ls -la $O
