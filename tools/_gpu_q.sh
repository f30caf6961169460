timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 200 python tools/bench_edit.py 10000000 dyn 2>&1 | tail -2
