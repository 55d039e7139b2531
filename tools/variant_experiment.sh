# One GPU call: pipe microbenchmarks, then every build in tools/variants/ on the same seeded workload
# (tools/profile_target.py prints a checksum of the outputs: equal checksums = same bits), then the GPU tests.
set -x
mkdir -p gpurun_out/var
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/var/gpu.txt
./tools/microbench/alu > gpurun_out/var/alu.txt 2>&1
./tools/microbench/bfly > gpurun_out/var/bfly.txt 2>&1
for v in tools/variants/*.so; do
  n=$(basename $v .so)
  NUFHE_B200_LIB=$PWD/$v timeout 300 python tools/profile_target.py 4096 16384 > gpurun_out/var/$n.b4096.txt 2>&1
  NUFHE_B200_LIB=$PWD/$v timeout 300 python tools/profile_target.py 592 16384 > gpurun_out/var/$n.b592.txt 2>&1
done
grep -h "TIMES\|checksum" gpurun_out/var/*.b*.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/var/pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/var/bench_nand.json 2> gpurun_out/var/bench_nand.err
cat gpurun_out/var/alu.txt gpurun_out/var/bfly.txt gpurun_out/var/pytest.txt gpurun_out/var/bench_nand.json
grep -H "TIMES\|checksum" gpurun_out/var/*.b*.txt
