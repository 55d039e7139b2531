# Round-2 GPU call 3: FIFO ready queue, CTA shapes with more resident warps, 3 CTAs/SM for the stand-alone transform.
set -x
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -8 > $O/pytest.txt
cat $O/pytest.txt
timeout 300 python tools/profile_target.py 4096 16384 2>&1 | grep "TIMES\|checksum" | sed "s/^/main /" >> $O/shapes.txt
for v in tools/variants/s*.so; do
  n=$(basename $v .so)
  NUFHE_B200_LIB=$PWD/$v timeout 300 python tools/profile_target.py 4096 16384 2>&1 | grep "TIMES\|checksum" | sed "s/^/$n /" >> $O/shapes.txt
done
cat $O/shapes.txt
SWEEP_BATCHES=1,64,256,296,400,592,600,768,1024,1536,2048,4096 timeout 600 python tools/sweep.py $O/sweep.json > $O/sweep.log 2> $O/sweep.err
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep.log
grep -o "'transforms.*" $O/sweep.log
NUFHE_B200_LIB=$PWD/tools/variants/ntt3.so SWEEP_BATCHES=64 timeout 300 python tools/sweep.py $O/sweep_ntt3.json 2>&1 | grep -o "'transforms.*" | sed "s/^/ntt3 /"
tail -3 $O/sweep.err
