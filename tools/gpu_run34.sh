set -x
mkdir -p gpurun_out
{
for r in 0 1 0 1; do
echo "SILU_TANH=$r"
IPDM_SILU_TANH=$r timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
IPDM_SILU_TANH=$r timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
IPDM_SILU_TANH=$r timeout 60 python tools/one_conv.py 128 16 512 512 64 3 5 64 3 0
done
} > gpurun_out/r2_fold34.txt 2>&1
grep -v "^+" gpurun_out/r2_fold34.txt
timeout 300 python -m pytest tests/test_unet_kernels_gpu.py tests/test_unet_gpu.py -q -x -m gpu -s > gpurun_out/r2_t34.log 2>&1
grep -E "passed|failed|bf16" gpurun_out/r2_t34.log | grep -E "passed|failed|fused|full|proj|img" | tail -20
IPDM_SILU_TANH=0 timeout 600 python -m pytest tests/test_teacher_forced_gpu.py -q -s -m gpu -k "bf16" 2>&1 | grep "teacher-forced" | cut -c1-260
IPDM_SILU_TANH=1 timeout 600 python -m pytest tests/test_teacher_forced_gpu.py -q -s -m gpu -k "bf16" 2>&1 | grep -E "teacher-forced|passed|failed" | cut -c1-260
