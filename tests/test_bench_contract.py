"""bench.py's driver contract, the parts that need no GPU: the reference arm (`--impl reference`) prints ONE JSON line
with the agreed keys, only rank 0 prints under a multi-rank launch, and the product arm refuses to run without a
CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.pop("WORLD_SIZE", None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=e, timeout=timeout)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample-batch", "256"])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1      # exactly K timed steps after W warm-ups
    for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "B=256" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank0_prints():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample-batch", "128"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on the GPU box the product arm runs for real (bench itself is the driver's job there)
    r = _run(["--steps", "1", "--warmup", "3", "--no-extra", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
