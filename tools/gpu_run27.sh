set -x
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_fused_dense128_r02 python tools/one_conv.py 128 16 500 228 128 3 5 > gpurun_out/r2_ncu27.log 2>&1
IPDM_TIME_STATS=1 timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
IPDM_TIME_STATS=1 timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
