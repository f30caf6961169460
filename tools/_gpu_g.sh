SGN_TC_DEBUG=32 python tools/bench_agg.py --iters 2 2>&1 | grep -E "colour mma issuer" | tail -2
