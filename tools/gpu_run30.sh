set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fwd_bf16_b16.csv python tools/one_forward.py 2 bf16 16 both > gpurun_out/r2_ncu30a.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_fused_dense128 python tools/one_conv.py 128 16 500 228 128 3 5 > gpurun_out/r2_ncu30b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_fused_fold8 python tools/one_conv.py 8 16 2000 912 8 1 6 0 3 1 >> gpurun_out/r2_ncu30b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo_persistent --launch-skip 3 -c 1 -f -o gpurun_out/prof_r02_halo_pers128 python tools/one_conv.py 128 16 500 228 128 3 4 >> gpurun_out/r2_ncu30b.log 2>&1
rm -f gpurun_out/trace_r02_final.txt
IPDM_OP_TRACE=gpurun_out/trace_r02_final.txt timeout 300 python tools/one_forward.py 3 bf16 16 both > gpurun_out/r2_fwd30.log 2>&1
python tools/op_trace.py gpurun_out/trace_r02_final.txt 2 > gpurun_out/r02_trace_proj.txt
python tools/op_trace.py gpurun_out/trace_r02_final.txt 5 > gpurun_out/r02_trace_img.txt
timeout 300 python tools/bench_conv.py 16 4,5 > gpurun_out/r02_bench_conv.md 2>&1 || true
