set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py tests/test_guided_gpu.py -q -s -m gpu -k "fused_groupnorm or adaptive_schedule or philox or groupnorm_stats" > gpurun_out/r2_t3.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t3.log
timeout 600 python tools/bench_conv.py 16 4,5 > gpurun_out/r2_bench_conv.md 2>&1
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -5 gpurun_out/r2_t3.log
