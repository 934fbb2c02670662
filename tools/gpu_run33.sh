set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu > gpurun_out/r2_t33.log 2>&1
tail -2 gpurun_out/r2_t33.log
{
for r in 0 1 0 1; do
echo "RAW3=$r"
IPDM_RAW3=$r timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
IPDM_RAW3=$r timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5 64 3 0
IPDM_RAW3=$r timeout 60 python tools/one_conv.py 128 16 512 512 64 3 5 64 3 0
done
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 4
} > gpurun_out/r2_fold33.txt 2>&1
grep -v "^+" gpurun_out/r2_fold33.txt
