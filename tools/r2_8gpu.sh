set -x
O=gpurun_out/r2g8
mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; cat $O/bench_8gpu.json | cut -c1-3000; tail -3 $O/bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > $O/bench_ref_8gpu.json 2> $O/bench_ref_8gpu.err; cut -c1-400 $O/bench_ref_8gpu.json
