set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py -q -s -m gpu -k "fused_groupnorm or warp_mma" > gpurun_out/r2_t8.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t8.log
for dbg in 0 3; do
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 128 16 500 228 128 3 5
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 64 16 512 512 64 3 5
done > gpurun_out/r2_fuse_dbg3.txt 2>&1
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_guided_gpu.py -q -s -m gpu -x > gpurun_out/r2_t8b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t8b.log
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
rm -f gpurun_out/trace_r2b.txt
IPDM_OP_TRACE=gpurun_out/trace_r2b.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd8.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2b.txt 2 > gpurun_out/r2_trace8_proj.txt 2>&1
python tools/op_trace.py gpurun_out/trace_r2b.txt 5 > gpurun_out/r2_trace8_img.txt 2>&1
tail -5 gpurun_out/r2_t8.log gpurun_out/r2_t8b.log; grep -v "^+" gpurun_out/r2_fuse_dbg3.txt
