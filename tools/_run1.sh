python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo rc=$?; tail -2 gpurun_out/r02_bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_8gpu.json')); print(d['value'], d['e2e']['value'], d['e2e_pipelined']['value'], d['pipeline_sharded']['scans_per_s'], d['numa'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_4gpu.json')); print(d['value'], d['e2e']['value'], d['e2e_pipelined']['value'], d['pipeline_sharded']['scans_per_s'], d['numa'])"
nproc; nvidia-smi topo -m | head -12
