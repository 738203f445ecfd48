mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rowstream -s 2 -c 1 -o gpurun_out/r02_rowstream_c5shard python tools/rowpass_one.py 8192 65536 complex64 4 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowstream -s 2 -c 1 -o gpurun_out/r02_rowstream_c2 python tools/rowpass_one.py 16384 65536 float32 4 > gpurun_out/p2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260 --csv --log-file gpurun_out/r02_launches_bench_c5_n1.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/p3.log 2>&1
tail -2 gpurun_out/p1.log gpurun_out/p3.log
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_multi.py -q -m gpu -k "c2_full or group" 2>&1 | tail -4
SAN_TIMEOUT=600 bash tools/sanitize.sh
