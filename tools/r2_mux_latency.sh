O=gpurun_out/r2mux; mkdir -p $O
timeout 200 python tools/mux_latency.py $O/mux_latency.json 2>&1 | grep batch
NUFHE_B200_PAIR_MAX=0 timeout 200 python tools/mux_latency.py $O/mux_latency_nopair.json 2>&1 | grep batch | sed "s/^/nopair /"
