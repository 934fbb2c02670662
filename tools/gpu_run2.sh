set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py -q -s -m gpu -k "fused_groupnorm or halo" > gpurun_out/r2_t2.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t2.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_guided_gpu.py -q -s -m gpu > gpurun_out/r2_t2b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t2b.log
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
IPDM_GN_FUSE=0 timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench2_nofuse.json 2> gpurun_out/r2_bench2_nofuse.err
rm -f gpurun_out/trace_r2.txt
IPDM_OP_TRACE=gpurun_out/trace_r2.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2.txt 2 > gpurun_out/r2_trace_proj.txt 2>&1
python tools/op_trace.py gpurun_out/trace_r2.txt 5 > gpurun_out/r2_trace_img.txt 2>&1
tail -5 gpurun_out/r2_t2.log gpurun_out/r2_t2b.log
