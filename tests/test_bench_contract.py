"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own PointAggregator + ray_march from baseline/_ref -- the oracle
restatement when that is not staged -- after the sequential C query, on the host cores) prints one JSON line with the keys the bench
contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                   "--cpu-sample", "64"], cwd=ROOT, timeout=600).decode()
    assert len(out.strip().splitlines()) == 1, "stdout must carry the JSON line only"
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
