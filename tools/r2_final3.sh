# Final round-2 measurement on one B200 (final build): all GPU tests, bench lines,
# sweep, ncu launch list, DRAM traffic at batch 4096, full captures at batch 592 and batch 1.
set -x
O=gpurun_out/r2fin3
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --durations=6 2>&1 | tail -14 > $O/pytest.txt
cat $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; cat $O/smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_nand.json 2> $O/bench_nand.err; cat $O/bench_nand.json; tail -2 $O/bench_nand.err
timeout 600 python bench.py --steps 5 --warmup 3 --gate mux --no-extras > $O/bench_mux.json 2> $O/bench_mux.err; cat $O/bench_mux.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
SWEEP_BATCHES=1,16,48,64,148,256,592,768,1024,1536,2048,4096,16384,65536 timeout 900 python tools/sweep.py $O/sweep.json > $O/sweep.log 2> $O/sweep.err
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep.log; grep -o "'transforms.*" $O/sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/launch_bench.log 2>&1
grep -c "gpu__time_duration" $O/launches.csv
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"blind_rotate|keyswitch_kernel" -s 4 -c 2 --csv --log-file $O/traffic_b4096.csv python tools/profile_target.py 4096 4096 > $O/traffic.log 2>&1
python tools/ncu_traffic.py $O/traffic_b4096.csv 4096 $O/r2_traffic.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate|keyswitch_kernel|ntt_forward|ntt_inverse" -s 8 -c 4 -o $O/r2_final python tools/profile_target.py 592 16384 > $O/prof.log 2>&1
python tools/ncu_summary.py $O/r2_final.ncu-rep > $O/r2_final_summary.txt 2>&1
cat $O/r2_final_summary.txt
# the lowest-latency shape alone on the machine: one ciphertext on a cluster of two SMs (pair shape)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate" -s 2 -c 1 -o $O/r2_final_b1 python tools/profile_target.py 1 64 > $O/prof_b1.log 2>&1
python tools/ncu_summary.py $O/r2_final_b1.ncu-rep > $O/r2_final_b1_summary.txt 2>&1
cat $O/r2_final_b1_summary.txt
