set -x
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
python - <<'P'
import json;d=json.loads(open('gpurun_out/r2_bench32.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline'].get('mixed'))
P
tail -3 gpurun_out/r2_bench32.err
