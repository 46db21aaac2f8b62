#!/bin/bash
# GPU experiment batch: decode-kernel variants (tests + microbench + ncu). Output -> gpurun_out/
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
for v in v2 v3 v1; do
  B200Q_GEMV=$v timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_$v.log
  echo "== pytest $v: $(tail -1 $O/pytest_$v.log)"
done
for v in v2 v3; do
  B200Q_GEMV=$v timeout 300 python tools/microbench.py --m 1 --graph --iters 400 > $O/mb_$v.log 2>&1
  echo "== microbench $v"; cat $O/mb_$v.log
done
B200Q_GEMV=v3 B200Q_SLICE_KB=24 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 > $O/mb_v3_s24.log 2>&1; echo "== v3 slice24"; cat $O/mb_v3_s24.log
B200Q_GEMV=v3 B200Q_SLICE_KB=72 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 > $O/mb_v3_s72.log 2>&1; echo "== v3 slice72"; cat $O/mb_v3_s72.log
B200Q_GEMV=v2 B200Q_MAX_CLUSTER=4 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ > $O/mb_v2_c4.log 2>&1; echo "== v2 cluster4"; cat $O/mb_v2_c4.log
B200Q_GEMV=v2 timeout 300 python tools/microbench.py --m 2,4,8 --graph --iters 200 --layouts GEMM --shapes 4096x4096,4096x11008 > $O/mb_v2_m.log 2>&1; echo "== v2 M=2,4,8"; cat $O/mb_v2_m.log
# eager (no graph) for launch-overhead comparison
B200Q_GEMV=v2 timeout 300 python tools/microbench.py --m 1 --iters 400 --layouts GEMM > $O/mb_v2_eager.log 2>&1; echo "== v2 eager"; cat $O/mb_v2_eager.log
# ncu: launch list + full sections for the decode kernels
for v in v2 v3; do
  B200Q_GEMV=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 20 -c 4 -o $O/ncu_gemv_$v -f \
     python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096,4096x11008 > $O/ncu_$v.log 2>&1
  echo "== ncu $v done: $(tail -2 $O/ncu_$v.log | head -1)"
done
B200Q_GEMV=v2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemv -c 60 --csv --log-file $O/launches_gemv_v2.csv \
   python tools/microbench.py --m 1 --iters 8 --layouts GEMM,GPTQ,MARLIN > /dev/null 2>&1
ls -la $O
