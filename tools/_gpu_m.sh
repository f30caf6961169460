python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2e_n2.json 2> gpurun_out/bench_r2e_n2.err; tail -c 600 gpurun_out/bench_r2e_n2.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2e_n2.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "n", d["n_gpus"])
print("train", d["train_step"])
print("strong", json.dumps(d["strong"])[:1600])
PY
