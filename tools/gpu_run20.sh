set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t20.log 2>&1
tail -3 gpurun_out/r2_t20.log
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
python - <<'P'
import json;d=json.loads(open('gpurun_out/r2_bench20.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac']);[print(k,v) for k,v in d['kernel_families'].items()]
P
