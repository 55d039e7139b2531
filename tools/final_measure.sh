# Round-end measurement on one B200: GPU tests, bench lines, sweep, ncu launch list + full capture.
set -x
O=gpurun_out/fin
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > $O/pytest.txt
python bench.py --steps 5 --warmup 3 > $O/bench_nand.json 2> $O/bench_nand.err
python bench.py --steps 5 --warmup 3 --gate mux > $O/bench_mux.json 2> $O/bench_mux.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python tools/sweep.py $O/sweep.json > /dev/null 2> $O/sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate|keyswitch_kernel|ntt_forward|ntt_inverse" -c 4 -o $O/r1b_fin python tools/profile_target.py 592 16384 > $O/prof.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
cat $O/pytest.txt $O/bench_nand.json $O/bench_mux.json $O/bench_ref.json $O/smoke.txt
