for i in 1 2 3; do timeout 90 python tools/bench_agg.py --iters 6 2>&1 | tail -1; echo "rc=$?"; done
