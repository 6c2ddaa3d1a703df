set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r2n_plain.json 2> gpurun_out/r2n_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r2n_ncu1.log 2>&1; echo "launchlist rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^scan_kernel -s 40 -c 1 -f -o gpurun_out/r2_scan_f32cos_full python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-parity > gpurun_out/r2n_ncu2.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r2_scan_f32cos_full.ncu-rep --page raw --csv > gpurun_out/r2_scan_f32cos_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_scan_f32cos_full.ncu-rep --page details > gpurun_out/r2_scan_f32cos_full_details.txt 2>/dev/null
timeout 300 python tools/hnsw_latency.py 200000 > gpurun_out/r2n_lat_plain.txt 2>&1; echo "lat rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_cta_kernel -s 3 -c 1 -f -o gpurun_out/r2_hnsw_cta_full python tools/hnsw_latency.py 200000 > gpurun_out/r2n_ncu3.log 2>&1; echo "cta rc=$?"
ncu -i gpurun_out/r2_hnsw_cta_full.ncu-rep --page raw --csv > gpurun_out/r2_hnsw_cta_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_hnsw_cta_full.ncu-rep --page details > gpurun_out/r2_hnsw_cta_full_details.txt 2>/dev/null
ls -la gpurun_out | tail -20
