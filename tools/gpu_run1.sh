set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_smi.txt
nproc >> gpurun_out/r2_smi.txt
timeout 1500 python -m pytest tests/test_teacher_forced_gpu.py tests/test_progressive_gpu.py -q -s -m gpu > gpurun_out/r2_t1.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t1.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_teacher_forced_gpu.py --deselect tests/test_progressive_gpu.py > gpurun_out/r2_t1b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t1b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'moments|apply_kernel|fbp_' -c 16 -o gpurun_out/prof_sampler_fbp_r02 python tools/sampler_fbp_once.py > gpurun_out/r2_ncu_sf.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2_t1.log
