python -m pytest tests -m gpu -q -x 2>&1 | tail -4; python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 800 gpurun_out/bench_r2b.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2b.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print("hbm", {k:(v["ms"],v["frac"]) for k,v in d["hbm_stages"].items() if isinstance(v,dict)})
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print("cold", d["config"]["cold_frame_ms"], "sem", d["config"]["semantic_variant"])
print("train", d["train_step"])
print("strong", json.dumps(d["strong"])[:1500])
print("refgpu", json.dumps(d["reference_gpu"])[:800])
print("cpu", json.dumps(d.get("cpu_baseline"))[:800])
PY
