set -x
{
timeout 60 python tools/one_conv.py 128 16 1000 456 128 3 4 0 3 0
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 4
timeout 60 python tools/one_conv.py 128 16 500 228 128 3 5
timeout 60 python tools/one_conv.py 64 16 512 512 64 3 5
timeout 60 python tools/one_conv.py 128 16 500 228 128 1 5
timeout 60 python tools/one_conv.py 16 16 2000 912 16 1 7 0 3 0
} > gpurun_out/r2_fold22.txt 2>&1
grep -v "^+" gpurun_out/r2_fold22.txt
