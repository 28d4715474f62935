"""bench.py's reference arm (CPU): the JSON line carries the contract's keys, and -- in this container, where the reference's
sources are reachable -- it is the UNMODIFIED reference that is timed (kind == "reference"), at the GPU arm's K."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "stage2_reranked_triplets_per_s" and d["unit"] == "triplets/s"
    assert d["vs_baseline"] is None and d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "K=100" in d["config"]["workload"] and "K=100" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["kind"] == ("reference" if os.path.exists("/root/reference/src/blip_stage2.py") else "port")
