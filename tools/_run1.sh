python -m pytest tests/test_reference_driver_gpu.py -m gpu -q -rs -s > gpurun_out/drv.log 2>&1
grep -n "passed\|failed\|^FAILED\|scan [0-9]: voxels\|SKIP\|Error\|integrations\|^E  " gpurun_out/drv.log | cut -c1-400 | head -60
