set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_kernels_gpu.py -q -x -m gpu -k "width_folded" > gpurun_out/r2_t13.log 2>&1
tail -15 gpurun_out/r2_t13.log
{
for v in 6 7; do
python tools/one_conv.py 8 16 2000 912 8 1 $v 0 3 1
python tools/one_conv.py 16 16 1000 456 16 1 $v 0 3 1
python tools/one_conv.py 16 16 2000 912 8 1 $v 8 3 0
python tools/one_conv.py 128 16 1000 456 16 1 $v 16 3 0
python tools/one_conv.py 4 16 2000 912 8 1 $v 0 3 0
done
python tools/one_conv.py 16 16 2000 912 8 1 7 8 1 0
python tools/one_conv.py 8 16 2000 912 8 1 7 8 1 0
python tools/one_conv.py 128 16 1000 456 16 1 7 16 1 0
} > gpurun_out/r2_fold13.txt 2>&1
cat gpurun_out/r2_fold13.txt
