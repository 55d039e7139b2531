set -x
O=gpurun_out/r2mg2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -q -m gpu -x -k "rank1 or callers_device or multi" 2>&1 | tail -5 > $O/pytest_multi.txt; cat $O/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cat $O/bench_2gpu.json | head -c 700; tail -3 $O/bench_2gpu.err
