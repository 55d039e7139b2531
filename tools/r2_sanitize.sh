set -x
O=gpurun_out/r2san
mkdir -p $O
export NUFHE_B200_FORCE_CHUNKS=3
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py > $O/$tool.txt 2>&1
  tail -3 $O/$tool.txt
done
SANITIZE_BATCH=620 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 2000 python tools/sanitize_target.py > $O/racecheck_all.txt 2>&1
tail -3 $O/racecheck_all.txt
grep -o "in [a-z_0-9:]*kernel[^ (]*\|at nb::[a-z_0-9]*" $O/racecheck_all.txt | sort | uniq -c | sort -rn | head -10
