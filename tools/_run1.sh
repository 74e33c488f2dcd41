python bench.py --steps 3 --warmup 3 --no-cpu-baseline --scans-per-step 256 --manifest-scans 200 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; echo rc=$?; tail -5 gpurun_out/bench_r02_c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c.json')); print(d['pipeline_sharded']); print(d['e2e_deform']); print(d['value'], d['e2e']['value'])"
