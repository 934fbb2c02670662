set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t31.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t31.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke31.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench31.json 2> gpurun_out/r2_bench31.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench31_ref.json 2> gpurun_out/r2_bench31_ref.err
tail -3 gpurun_out/r2_t31.log; tail -2 gpurun_out/r2_smoke31.log
