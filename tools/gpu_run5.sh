set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py -q -s -m gpu -k "fused_groupnorm" > gpurun_out/r2_t5.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t5.log
for dbg in 0 2 3; do
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 128 16 500 228 128 3 5
  IPDM_FUSE_DBG=$dbg python tools/one_conv.py 64 16 512 512 64 3 5
done > gpurun_out/r2_fuse_dbg2.txt 2>&1
timeout 600 python tools/bench_conv.py 16 4,5 > gpurun_out/r2_bench_conv2.md 2>&1
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -3 gpurun_out/r2_t5.log; grep -v "^+" gpurun_out/r2_fuse_dbg2.txt
