# Round-2 GPU call 1: parity tests on the new kernel, A/B of the arithmetic variants, batch sweep of the scheduler.
set -x
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
nproc >> $O/gpu.txt
for v in tools/variants/*.so; do
  n=$(basename $v .so)
  NUFHE_B200_LIB=$PWD/$v timeout 300 python tools/profile_target.py 4096 16384 > $O/var_$n.txt 2>&1
done
grep -H "TIMES\|checksum" $O/var_*.txt
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -25 > $O/pytest.txt
cat $O/pytest.txt
SWEEP_BATCHES=1,64,256,296,400,592,600,768,1024,1536,2048,4096,16384 timeout 600 python tools/sweep.py $O/sweep.json > $O/sweep.log 2> $O/sweep.err
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep.log
NUFHE_B200_MAX_CHUNKS=1 SWEEP_BATCHES=600,768,1024,1536,2048 timeout 300 python tools/sweep.py $O/sweep_nochunks.json > $O/sweep_nochunks.log 2>&1
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep_nochunks.log
tail -5 $O/sweep.err
