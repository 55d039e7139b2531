#!/bin/bash
# pair shape (one ciphertext on a cluster of two SMs): parity tests, then latency against the single-CTA shapes
O=gpurun_out/r2pair; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 120 python tools/pair_probe.py 3 2>&1 | tail -5 | tee $O/probe.txt
grep -q "pair probe ok" $O/probe.txt || { echo "probe failed"; exit 1; }
timeout 120 python tools/pair_probe.py 80 2>&1 | tail -4 | tee -a $O/probe.txt
timeout 900 python -m pytest tests/test_gpu_gates.py tests/test_gpu_api.py -m gpu -x -q \
  -k "all_cta_shapes or rare_path or nand32 or mux_matches or truth_table or gate_mux or uint_min or empty" 2>&1 | tail -8 > $O/pytest_subset.txt
cat $O/pytest_subset.txt
NUFHE_B200_PAIR_ASYNC=0 timeout 600 python -m pytest tests/test_gpu_gates.py -m gpu -x -q -k "all_cta_shapes or rare_path" 2>&1 | tail -3 | sed "s/^/barrier-variant /" | tee $O/pytest_barrier_variant.txt
NUFHE_B200_VERBOSE=1 SWEEP_BATCHES=1,16,37,64,74 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_pair.json 2>&1 | grep -o "nufhe_b200:.*\|'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/pair_async /"
NUFHE_B200_PAIR_ASYNC=0 SWEEP_BATCHES=1,16,37,64,74 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_pair_barrier.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/pair_barrier /"
NUFHE_B200_PAIR_MAX=0 SWEEP_BATCHES=1,16,37,64,74 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_nopair.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/nopair /"
NUFHE_B200_PAIR_MAX=1000 SWEEP_BATCHES=80,100,148 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_pairall.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/pair_forced /"
