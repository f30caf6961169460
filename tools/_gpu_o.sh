timeout 300 python -m pytest tests/test_gpu_voxgrid.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python tools/bench_edit.py 2>&1 | tail -4
