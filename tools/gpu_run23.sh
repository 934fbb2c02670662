set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_unet_kernels_gpu.py -q -x -m gpu -s > gpurun_out/r2_t23.log 2>&1
grep -E "passed|failed|rel-L2|Error|error" gpurun_out/r2_t23.log | tail -30
rm -f gpurun_out/trace_r2f.txt
IPDM_OP_TRACE=gpurun_out/trace_r2f.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd23.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2f.txt 2 > gpurun_out/r2_trace23_proj.txt
python tools/op_trace.py gpurun_out/trace_r2f.txt 5 > gpurun_out/r2_trace23_img.txt
head -12 gpurun_out/r2_trace23_proj.txt; head -12 gpurun_out/r2_trace23_img.txt
