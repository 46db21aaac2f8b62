#!/bin/bash
O=gpurun_out/e43; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest.txt
echo "== act-order microbench (Llama-2-13B shapes)"
timeout 300 python tools/microbench.py --m 1 --graph --iters 300 --layouts GPTQ,GPTQ_ACT --shapes 5120x5120,5120x13824,13824x5120 2>&1 | tee $O/mb_actorder_decode.log | cut -c1-200
timeout 300 python tools/microbench.py --m 512 --iters 30 --layouts GPTQ,GPTQ_ACT --shapes 5120x5120,5120x13824,13824x5120 2>&1 | tee $O/mb_actorder_gemm.log | cut -c1-200
B200Q_ACTORDER_RELAYOUT=0 timeout 300 python tools/microbench.py --m 1 --iters 30 --layouts GPTQ_ACT --shapes 5120x5120,5120x13824 2>&1 | tee $O/mb_actorder_generic.log | cut -c1-200
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | head -3 | cut -c1-250 | tee $O/gemm_timeline_head.txt
