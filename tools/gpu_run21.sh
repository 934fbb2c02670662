set -x
mkdir -p gpurun_out
rm -f gpurun_out/trace_r2e.txt
IPDM_OP_TRACE=gpurun_out/trace_r2e.txt timeout 600 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd21.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2e.txt 2 > gpurun_out/r2_trace21_proj.txt
python tools/op_trace.py gpurun_out/trace_r2e.txt 5 > gpurun_out/r2_trace21_img.txt
