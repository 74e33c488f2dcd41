python -m pytest tests -m gpu -q 2>&1 | tail -25
for k in 0 1 2 4; do VLIDAR_COPY_THREADS=$k python tools/ctrace_bench.py 710 20 2>&1 | tail -1; done
nproc; lscpu | grep -E "Model name|Socket|NUMA"
