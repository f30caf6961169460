timeout 300 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_parity_configs.py -m gpu -q -x 2>&1 | tail -4
timeout 120 python tools/bench_agg.py --iters 4 2>&1 | tail -1
SGN_TC_DEBUG=32 timeout 120 python tools/bench_agg.py --iters 2 2>&1 | grep -E "colour mma issuer" | tail -1
