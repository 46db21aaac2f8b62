#!/bin/bash
# GPU experiment batch 2: integer-MMA decode kernel + tcgen05 GEMM bring-up. Output -> gpurun_out/
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not tcgen05" 2>&1 | tail -15 > $O/pytest_decode.log; echo "== pytest decode (v3 smem): $(tail -1 $O/pytest_decode.log)"; grep -E "FAILED|Error" $O/pytest_decode.log | head
B200Q_GEMV=v2 timeout 900 python -m pytest tests -m gpu -q -k "decode or full_size_prop or workspace" 2>&1 | tail -8 > $O/pytest_decode_regs.log; echo "== pytest decode (regs): $(tail -1 $O/pytest_decode_regs.log)"
timeout 600 python -m pytest tests -m gpu -q -k "tcgen05" 2>&1 | tail -40 > $O/pytest_gemm.log; echo "== pytest gemm:"; tail -25 $O/pytest_gemm.log
python - <<'PY'
import qllm_b200, ctypes
print("gemm err flag (0 = no mbarrier timeout):", "n/a")
PY
for v in v3 v2; do
  B200Q_GEMV=$v timeout 300 python tools/microbench.py --m 1 --graph --iters 400 > $O/mb2_$v.log 2>&1
  echo "== microbench $v"; cat $O/mb2_$v.log
done
B200Q_GEMV=v3 timeout 300 python tools/microbench.py --m 2,8 --graph --iters 200 --layouts GEMM --shapes 4096x4096 > $O/mb2_m.log 2>&1; cat $O/mb2_m.log
timeout 300 python tools/microbench.py --m 16,64,128,512,2048 --iters 50 --layouts GPTQ --shapes 4096x4096,4096x11008 > $O/mb2_gemm.log 2>&1; echo "== gemm microbench"; cat $O/mb2_gemm.log
B200Q_GEMV=v3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 20 -c 2 -o $O/ncu2_gemv_v3 -f \
   python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096 > $O/ncu2_v3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o $O/ncu2_gemm -f \
   python tools/microbench.py --m 512 --iters 4 --layouts GPTQ --shapes 4096x4096 > $O/ncu2_gemm.log 2>&1
ls $O | head -50
