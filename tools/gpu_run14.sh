set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu -k "width_folded" > gpurun_out/r2_t14.log 2>&1
tail -5 gpurun_out/r2_t14.log
ncu --set full --clock-control none --import-source on -k regex:conv_halo_persistent --launch-skip 3 -c 1 -f -o gpurun_out/prof_fold_plain_r02 python tools/one_conv.py 8 16 2000 912 8 1 7 0 3 1 > gpurun_out/r2_ncu14.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_fold_fused_r02 python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1 >> gpurun_out/r2_ncu14.log 2>&1
python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1
python tools/one_conv.py 16 16 1000 456 16 1 6 0 3 1
python tools/one_conv.py 4 16 2000 912 8 1 6 0 3 0
