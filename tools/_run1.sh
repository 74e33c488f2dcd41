python tools/deform_timing.py 4 2>&1 | tail -45
python -m pytest tests/test_dropin_gpu.py tests/test_reference_driver_gpu.py tests/test_project_tsdf_gpu.py tests/test_chain_gpu.py -m gpu -q -x 2>&1 | tail -5
