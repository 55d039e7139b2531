set -x
mkdir -p gpurun_out/fin
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/fin/pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/fin/bench_nand.json 2> gpurun_out/fin/bench_nand.err
python bench.py --steps 5 --warmup 3 --gate mux > gpurun_out/fin/bench_mux.json 2> gpurun_out/fin/bench_mux.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin/bench_ref.json 2> gpurun_out/fin/bench_ref.err
python tools/sweep.py gpurun_out/fin/sweep.json > /dev/null 2> gpurun_out/fin/sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/fin/launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate|keyswitch_kernel|ntt_forward|ntt_inverse" -c 4 -o gpurun_out/fin/r1_fin python tools/profile_target.py 592 16384 > gpurun_out/fin/prof.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin/smoke.txt 2>&1
cat gpurun_out/fin/pytest.txt gpurun_out/fin/bench_nand.json gpurun_out/fin/bench_mux.json gpurun_out/fin/bench_ref.json
