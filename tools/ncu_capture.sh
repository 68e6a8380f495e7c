#!/bin/bash
# ncu evidence of the final tree (run under gpurun, one GPU):  bash tools/ncu_capture.sh <tag>
# 1. launch list (every launch of one forward with its device time), 2. `--set full` of one launch of every hot kernel
# at the bench workload, 3. the tcgen05 attention kernel at the TACoS shape, 4. the tensor-pipe counter of a plain cuBLAS
# bf16 GEMM as the calibration of what "100 %" reads on this counter.
set -u
TAG=${1:-r02}
O=gpurun_out
CMD="python bench.py --kernel-only --steps 2 --warmup 3"
$CMD > $O/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 156 -c 104 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_list_$TAG.log 2>&1
for k in layer_kernel gemm_pair_kernel mlp_chain_kernel gemm_group_kernel inproj_kernel attn_video_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o $O/prof_${TAG}_$k $CMD > $O/ncu_${TAG}_$k.log 2>&1
done
CMD2="python bench.py --preset tacos --kernel-only --steps 2 --warmup 3"
$CMD2 > $O/plain_${TAG}_tacos.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 4 -c 1 -f -o $O/prof_${TAG}_attn_tc_kernel $CMD2 > $O/ncu_${TAG}_attn_tc.log 2>&1
python - > $O/plain_${TAG}_cublas.log 2>&1 <<PY
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): c = a @ b
torch.cuda.synchronize()
PY
cat > /tmp/cublas_cal.py <<PY
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): c = a @ b
torch.cuda.synchronize()
PY
ncu --set full --clock-control none -k regex:nvjet -s 1 -c 1 -f -o $O/prof_${TAG}_cublas_bf16_8192 python /tmp/cublas_cal.py > $O/ncu_${TAG}_cublas.log 2>&1
ls -la $O/prof_${TAG}_*.ncu-rep
