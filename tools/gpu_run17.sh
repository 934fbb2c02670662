set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py -q -x -m gpu -s > gpurun_out/r2_t17.log 2>&1
tail -12 gpurun_out/r2_t17.log
rm -f gpurun_out/trace_r2d.txt
IPDM_OP_TRACE=gpurun_out/trace_r2d.txt timeout 600 python tools/one_forward.py 3 bf16 16 proj > gpurun_out/r2_fwd17.log 2>&1
python tools/op_trace.py gpurun_out/trace_r2d.txt 2 > gpurun_out/r2_trace17_proj.txt
head -60 gpurun_out/r2_trace17_proj.txt
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench17.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']);[print(k,v) for k,v in d['kernel_families'].items()]"
