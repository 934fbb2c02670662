set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu -k "width_folded or fused" > gpurun_out/r2_t24.log 2>&1
tail -2 gpurun_out/r2_t24.log
{
timeout 60 python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1
timeout 60 python tools/one_conv.py 16 16 1000 456 16 1 6 0 3 1
timeout 60 python tools/one_conv.py 16 16 2000 912 8 1 6 8 3 0
timeout 60 python tools/one_conv.py 128 16 1000 456 16 1 6 16 3 0
timeout 60 python tools/one_conv.py 4 16 2000 912 8 1 6 0 3 0
} > gpurun_out/r2_fold24.txt 2>&1
grep -v "^+" gpurun_out/r2_fold24.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_fold_fused5_r02 python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1 > gpurun_out/r2_ncu24.log 2>&1
