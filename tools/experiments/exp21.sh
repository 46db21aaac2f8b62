#!/bin/bash
O=gpurun_out/e21; mkdir -p $O
echo "== timeline (graph) stream"; for sh in 4096x4096 4096x11008; do timeout 200 python tools/timeline.py --layout GEMM --shape $sh --launches 8 2>&1 | tail -9 | tee -a $O/timeline_stream.txt; done
echo "== timeline (graph) stream, default carveout"; B200Q_CARVEOUT=0 timeout 200 python tools/timeline.py --layout GEMM --shape 4096x4096 --launches 8 2>&1 | tail -9 | tee -a $O/timeline_stream_nocarve.txt
echo "== timeline (graph) legacy"; for sh in 4096x4096 4096x11008; do B200Q_GEMV=rp timeout 200 python tools/timeline.py --layout GEMM --shape $sh --launches 8 2>&1 | tail -9 | tee -a $O/timeline_legacy.txt; done
echo "== microbench stream cluster 8"; B200Q_ST_CLUSTER=8 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM --shapes 4096x4096,11008x4096 2>&1 | tee $O/mb_stream_c8.log | cut -c1-150
echo "== microbench legacy carveout0"; B200Q_CARVEOUT=0 B200Q_GEMV=rp timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM 2>&1 | tee $O/mb_legacy_nocarve.log | cut -c1-150
echo "== microbench legacy carveout"; B200Q_GEMV=rp timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM 2>&1 | tee $O/mb_legacy_carve.log | cut -c1-150
echo "== gemm timeline"; timeout 200 python tools/gemm_timeline.py 512 2>&1 | tee $O/gemm_timeline_512.txt
timeout 200 python tools/gemm_timeline.py 2048 2>&1 | tail -6 | tee $O/gemm_timeline_2048.txt
