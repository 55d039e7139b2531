set -x
O=gpurun_out/r2chk
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -4 > $O/pytest.txt; cat $O/pytest.txt
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_time.txt; cat $O/bench_time.txt | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2chk/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['traffic_source'], d['parity_checked'])
for r in d['batch_sweep']: print(r)
print(d['mux']); print(d['ntt'])
PY
tail -3 $O/bench_default.err
