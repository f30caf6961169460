SGN_TC_DEBUG=512 timeout 90 python tools/bench_agg.py --iters 3 2>&1 | tail -1; echo "rc512=${PIPESTATUS[0]}"
timeout 60 python tools/bench_agg.py --iters 2 --width 320 --height 240 2>&1 | tail -1; echo "rc_small=${PIPESTATUS[0]}"
