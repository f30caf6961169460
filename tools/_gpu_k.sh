timeout 100 python tools/bench_agg.py --iters 8 2>&1 | tail -1; echo "rc=${PIPESTATUS[0]}"
timeout 100 python tools/bench_agg.py --iters 8 --semantic 2>&1 | tail -1; echo "rc=${PIPESTATUS[0]}"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
