set -x
mkdir -p gpurun_out
for idle in 0 1 0 1; do
rm -f gpurun_out/trace_tmp.txt
IPDM_ATTN_IDLE=$idle IPDM_OP_TRACE=gpurun_out/trace_tmp.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > /dev/null 2>&1
echo "idle=$idle" >> gpurun_out/r2_attn29.txt
python tools/op_trace.py gpurun_out/trace_tmp.txt 2 | grep -E "^forward|attention" >> gpurun_out/r2_attn29.txt
python tools/op_trace.py gpurun_out/trace_tmp.txt 5 | grep -E "^forward|attention" >> gpurun_out/r2_attn29.txt
done
cat gpurun_out/r2_attn29.txt
