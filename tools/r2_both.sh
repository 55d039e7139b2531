#!/bin/bash
# forward pass 1 with both digits per thread (shared rotation): parity subset and throughput
O=gpurun_out/r2both; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gates.py tests/test_gpu_kernels.py -m gpu -x -q \
  -k "all_cta_shapes or rare_path or nand32 or mux_matches or more_than_one_wave or time_sliced_launches or external or blind_rotate" 2>&1 | tail -4 | tee $O/pytest_subset.txt
for i in 1 2; do timeout 300 python tools/profile_target.py 4096 4096 2>&1 | grep TIMES | sed "s/^/both_digits /" | tee -a $O/times.txt; done
NUFHE_B200_LIB=$PWD/tools/variants/r2_final_pair.so timeout 300 python tools/profile_target.py 4096 4096 2>&1 | grep TIMES | sed "s/^/before /" | tee -a $O/times.txt
SWEEP_BATCHES=592,1024,4096,16384 SWEEP_NO_MUX=1 timeout 600 python tools/sweep.py $O/sweep.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/both_digits /"
