set -x
mkdir -p gpurun_out
for dbg in 0 1 3 4 7; do IPDM_WARP_DBG=$dbg python tools/one_warp_conv.py 8 0 8 3 16 2000 912 5; done > gpurun_out/r2_warp_dbg.txt 2>&1
IPDM_WARP_DBG=0 python tools/one_warp_conv.py 16 0 16 3 16 1000 456 5 >> gpurun_out/r2_warp_dbg.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_warp_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_conv_warp_r02 python tools/one_warp_conv.py 8 0 8 3 16 2000 912 2 > gpurun_out/r2_ncu_warp.log 2>&1
grep -v "^+" gpurun_out/r2_warp_dbg.txt
