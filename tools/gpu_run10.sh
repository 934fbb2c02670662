set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
IPDM_THIN=0 timeout 600 python bench.py --steps 2 --warmup 2 --skip_extras --skip_cpu_baseline > gpurun_out/r2_bench10_nothin.json 2> gpurun_out/r2_bench10_nothin.err
timeout 900 python -m pytest tests/test_teacher_forced_gpu.py -q -s -m gpu -k "proj-bf16 or img-bf16" > gpurun_out/r2_t10.log 2>&1
IPDM_THIN=0 timeout 900 python -m pytest tests/test_teacher_forced_gpu.py -q -s -m gpu -k "proj-bf16" > gpurun_out/r2_t10_nothin.log 2>&1
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -m gpu > gpurun_out/r2_t10b.log 2>&1
grep -h "teacher-forced" gpurun_out/r2_t10.log gpurun_out/r2_t10_nothin.log
