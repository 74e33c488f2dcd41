for v in 1 0 1 0; do VL_CAST_REARM=$v python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; tail -2 gpurun_out/ab_$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/ab_$v.json')); print('rearm $v', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['clocks'])"; done
