#!/bin/bash
# pair shape + key staging in the 512-thread shape: parity subset, latency sweeps, one ncu capture of the pair kernel
O=gpurun_out/r2pair2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 120 python tools/pair_probe.py 3 2>&1 | tail -5 | tee $O/probe.txt
grep -q "pair probe ok" $O/probe.txt || { echo "probe failed"; exit 1; }
timeout 900 python -m pytest tests/test_gpu_gates.py tests/test_gpu_api.py tests/test_gpu_kernels.py -m gpu -x -q \
  -k "all_cta_shapes or rare_path or nand32 or mux_matches or truth_table or gate_mux or uint_min or empty or external or blind_rotate or time_sliced_mux" 2>&1 | tail -8 > $O/pytest_subset.txt
cat $O/pytest_subset.txt
NUFHE_B200_VERBOSE=1 SWEEP_BATCHES=1,8,16,24,37,48,55 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_pair.json 2>&1 | grep -o "nufhe_b200:.*\|'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/pair /"
NUFHE_B200_PAIR_MAX=1000 SWEEP_BATCHES=56,64,74 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_pair_forced.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/pair_forced /"
NUFHE_B200_PAIR_MAX=0 SWEEP_BATCHES=1,48,56,64,74 SWEEP_NO_MUX=1 timeout 300 python tools/sweep.py $O/sweep_nopair.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/nopair /"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate" -s 2 -c 1 -o $O/r2_pair_b1 python tools/profile_target.py 1 64 > $O/prof_b1.log 2>&1
python tools/ncu_summary.py $O/r2_pair_b1.ncu-rep > $O/r2_pair_b1_summary.txt 2>&1
cat $O/r2_pair_b1_summary.txt
