set -x
O=gpurun_out/r2e
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -5 > $O/pytest.txt; cat $O/pytest.txt
SWEEP_BATCHES=1,16,64,148,200,256,296,400,500,592,4096 timeout 400 python tools/sweep.py $O/sweep.json > $O/sweep.log 2> $O/sweep.err
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep.log | sed "s/^/default /"
NUFHE_B200_WIDE2_MAX=0 SWEEP_BATCHES=1,16,64,148 timeout 300 python tools/sweep.py $O/sweep_nowide2.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" | sed "s/^/wide2_off /"
NUFHE_B200_WIDE2_MAX=1000000 SWEEP_BATCHES=200,296 timeout 300 python tools/sweep.py $O/sweep_wide2all.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" | sed "s/^/wide2_all /"
for i in 1 2 3; do timeout 100 python tools/profile_target.py 1 64 2>&1 | grep TIMES; done
timeout 200 python examples/uint_min_graph.py 64 2>&1 | tail -1
timeout 200 python examples/uint_min_graph.py 512 2>&1 | tail -1
