python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-strong --no-reference-gpu --no-cpu-baseline --no-semantic-variant > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 500 gpurun_out/bench_r2d.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2d.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print("hbm", {k:(v["ms"],v["frac"]) for k,v in d["hbm_stages"].items() if isinstance(v,dict)})
print("train", d["train_step"])
PY
