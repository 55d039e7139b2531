set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -4 > $O/pytest.txt; cat $O/pytest.txt
SWEEP_BATCHES=1,16,64,148,200,222,256 timeout 300 python tools/sweep.py $O/sweep.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*"
for i in 1 2; do timeout 100 python tools/profile_target.py 1 64 2>&1 | grep TIMES; done
export NUFHE_B200_FORCE_CHUNKS=3
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_target.py 2>&1 | tail -2
SANITIZE_BATCH=310 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 2000 python tools/sanitize_target.py > $O/racecheck_all.txt 2>&1
tail -2 $O/racecheck_all.txt
grep -o "at nb::[a-z_0-9]*" $O/racecheck_all.txt | sort | uniq -c | sort -rn | head -5
