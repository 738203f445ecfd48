timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -25 > gpurun_out/n2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/ref_n2.json 2> gpurun_out/ref_n2.err
tail -25 gpurun_out/n2_tests.log; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/ref_n2.err; cat gpurun_out/ref_n2.json
