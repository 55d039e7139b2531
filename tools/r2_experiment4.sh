set -x
O=gpurun_out/r2d
mkdir -p $O
for i in 1 2; do
timeout 300 python tools/profile_target.py 4096 16384 2>&1 | grep "TIMES\|checksum" | sed "s/^/lazyadd1 /" >> $O/lazyadd.txt
NUFHE_B200_LIB=$PWD/tools/variants/lazyadd0.so timeout 300 python tools/profile_target.py 4096 16384 2>&1 | grep "TIMES\|checksum" | sed "s/^/lazyadd0 /" >> $O/lazyadd.txt
done
cat $O/lazyadd.txt
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -5 > $O/pytest.txt
cat $O/pytest.txt
SWEEP_BATCHES=1,592,4096 timeout 300 python tools/sweep.py $O/sweep.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*\|'transforms': [0-9]*, 'fwd_ms': [0-9.]*, 'inv_ms': [0-9.]*, 'fwd_gbs': [0-9.]*, 'inv_gbs': [0-9.]*"
