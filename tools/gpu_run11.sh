set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t11.log 2>&1
echo "rc=$?" >> gpurun_out/r2_t11.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
tail -3 gpurun_out/r2_t11.log; cat gpurun_out/r2_smoke.log | tail -2
