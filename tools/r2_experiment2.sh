# Round-2 GPU call 2: chunk overhead, stagger, thresholds, new stand-alone transform, bench line, ncu captures.
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "not full_4096 and not 65536" 2>&1 | tail -5 > $O/pytest.txt
cat $O/pytest.txt
for c in 1 2 5 10 25 50; do
  NUFHE_B200_FORCE_CHUNKS=$c timeout 200 python tools/profile_target.py 1184 4096 2>&1 | grep TIMES | sed "s/^/chunks=$c /" >> $O/chunks.txt
done
cat $O/chunks.txt
for st in 0 12000 24000 36000; do
  for b in 296 592; do
    NUFHE_B200_STAGGER=$st NUFHE_B200_WIDE_MAX=296 timeout 200 python tools/profile_target.py $b 4096 2>&1 | grep TIMES | sed "s/^/stagger=$st /" >> $O/stagger.txt
  done
done
cat $O/stagger.txt
for wm in 296 100000; do
  NUFHE_B200_WIDE_MAX=$wm SWEEP_BATCHES=340,400,450,500,550,592 timeout 300 python tools/sweep.py $O/sweep_wide$wm.json 2>&1 | grep -o "'batch': [0-9]*, 'ms': [0-9.]*" | sed "s/^/wide_max=$wm /" >> $O/wide.txt
done
cat $O/wide.txt
SWEEP_BATCHES=1,64,256,296,400,592,600,768,1024,1536,2048,4096 timeout 600 python tools/sweep.py $O/sweep.json > $O/sweep.log 2> $O/sweep.err
grep -o "'batch': [0-9]*, 'ms': [0-9.]*, 'ms_per_gate': [0-9.e-]*, 'gates_per_s': [0-9.]*" $O/sweep.log
grep -o "'transforms.*" $O/sweep.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_nand.json 2> $O/bench_nand.err
cat $O/bench_nand.json; tail -3 $O/bench_nand.err
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"blind_rotate|keyswitch_kernel" -s 4 -c 2 --csv --log-file $O/traffic_b4096.csv python tools/profile_target.py 4096 4096 > $O/traffic.log 2>&1
python tools/ncu_traffic.py $O/traffic_b4096.csv 4096 $O/r2_traffic.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate|keyswitch_kernel|ntt_forward|ntt_inverse" -s 8 -c 4 -o $O/r2b_full python tools/profile_target.py 592 16384 > $O/prof.log 2>&1
python tools/ncu_summary.py $O/r2b_full.ncu-rep > $O/r2b_full_summary.txt 2>&1
cat $O/r2b_full_summary.txt
