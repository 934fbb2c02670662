set -x
mkdir -p gpurun_out
for dbg in 0 2 3; do
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 128 16 500 228 128 3 5
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 64 16 512 512 64 3 5
done > gpurun_out/r2_fuse_dbg.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -o gpurun_out/prof_conv_fused_bf16_r02 python tools/one_conv.py 128 16 500 228 128 3 5 > gpurun_out/r2_ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -o gpurun_out/prof_conv_fused_tf32_r02 python tools/one_conv.py 128 16 500 228 128 1 5 >> gpurun_out/r2_ncu_fused.log 2>&1
cat gpurun_out/r2_fuse_dbg.txt
