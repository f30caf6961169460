python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_parity_configs.py -m gpu -q -x 2>&1 | tail -4
python tools/bench_agg.py --iters 4 2>&1 | tail -2
SGN_TC_DEBUG=32 python tools/bench_agg.py --iters 2 2>&1 | grep -E "mma issuer|epi warp 0:|epi warp 4:" | tail -3
