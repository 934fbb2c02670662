set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu > gpurun_out/r2_t28.log 2>&1
tail -2 gpurun_out/r2_t28.log
{
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
timeout 60 python tools/one_conv.py 128 16 500 228 128 1 5
timeout 60 python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1
timeout 60 python tools/one_conv.py 16 16 1000 456 16 1 6 0 3 1
timeout 60 python tools/one_conv.py 16 16 2000 912 8 1 6 8 3 0
timeout 60 python tools/one_conv.py 128 16 1000 456 16 1 6 16 3 0
} > gpurun_out/r2_fold28.txt 2>&1
grep -v "^+" gpurun_out/r2_fold28.txt
