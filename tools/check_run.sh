# GPU tests, then the batch sweep and one bench line (used after a kernel change)
set -x
O=gpurun_out/chk
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > $O/pytest.txt
python tools/sweep.py $O/sweep.json > /dev/null 2> $O/sweep.err
python bench.py --steps 5 --warmup 3 > $O/bench_nand.json 2> $O/bench_nand.err
NUFHE_B200_WIDE_MAX=0 python tools/sweep.py $O/sweep_nowide.json > /dev/null 2> $O/sweep_nowide.err
cat $O/pytest.txt $O/bench_nand.json
