python -m pytest tests/test_cast_gpu.py tests/test_reference_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/ab_units.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/ab_units.json')); print('value', round(d['value'],1), round(d['ms_per_step'],2)); print({k:round(v['ms_per_launch']*1e3,1) for k,v in d['stages'].items()})"
