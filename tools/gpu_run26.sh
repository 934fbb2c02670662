set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu -k "upsample_conv_phases" > gpurun_out/r2_t26.log 2>&1
tail -12 gpurun_out/r2_t26.log | cut -c1-220
timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -m gpu > gpurun_out/r2_t26b.log 2>&1
tail -3 gpurun_out/r2_t26b.log | cut -c1-220
rm -f gpurun_out/trace_r2g.txt
IPDM_OP_TRACE=gpurun_out/trace_r2g.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd26.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2g.txt 2 > gpurun_out/r2_trace26_proj.txt
python tools/op_trace.py gpurun_out/trace_r2g.txt 5 > gpurun_out/r2_trace26_img.txt
head -9 gpurun_out/r2_trace26_proj.txt; grep "1000x456x128\|500x228x128(+0) dst 1000" gpurun_out/r2_trace26_proj.txt; head -9 gpurun_out/r2_trace26_img.txt; grep "512x512x128(+0)\|256x256x128(+0) dst 512" gpurun_out/r2_trace26_img.txt
