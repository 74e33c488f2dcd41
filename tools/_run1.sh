python -m pytest tests/test_trace_gpu.py tests/test_trace_edge_gpu.py tests/test_cast_gpu.py -m gpu -q -x 2>&1 | tail -3
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_sparse_tsdf_gpu.py tests/test_trace_gpu.py tests/test_mesh_gpu.py -m gpu -q -x -k "(label0 or odd or sensor_inside or persistent or four_sweeps) and not 2048" > gpurun_out/r02_racecheck.log 2>&1; echo racecheck rc=$?; tail -3 gpurun_out/r02_racecheck.log
python tools/trace_modes.py 20 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --method lbvh --scans-per-step 256 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lbvh value', d['value']); print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['stages'].items()})"
