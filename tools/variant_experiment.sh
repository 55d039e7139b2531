# One GPU call: pipe microbenchmarks, then every build in tools/variants/ on the same seeded workload
# (tools/profile_target.py prints a checksum of the outputs: equal checksums = same bits).
set -x
mkdir -p gpurun_out/var

for v in tools/variants/*.so; do
  n=$(basename $v .so)
  NUFHE_B200_LIB=$PWD/$v timeout 300 python tools/profile_target.py 4096 16384 > gpurun_out/var/$n.b4096.txt 2>&1
done

grep -H "TIMES\|checksum" gpurun_out/var/*.b*.txt
