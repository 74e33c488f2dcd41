set -x
# (1) ncu --set full: the four cast kernels on the bench scene (second scan), and the config-1 chain's kernels
ncu --set full --clock-control none --import-source on -k regex:k_cast_ -s 4 -c 4 -o gpurun_out/r02_cast python tools/profile_cast.py 710 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_tsdf_|k_mesh_' -s 14 -c 14 -o gpurun_out/r02_chain python tools/profile_chain.py 2 1 > /dev/null 2>&1
# (2) launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --scans-per-step 64 > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bench_launches.csv
