"""bench.py's driver contract, the parts that run without a GPU: the reference arm (`--impl reference`: the oracle port
of Engine.play on the host cores) prints exactly ONE JSON line on stdout with the agreed keys, also under torchrun
where only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _one_json_line(stdout):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def _check(line, n_gpus):
    assert REQUIRED <= set(line), REQUIRED - set(line)
    assert line["impl"] == "reference" and line["n_gpus"] == n_gpus and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "env-steps/s" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "boat_race" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "boat_race" in cb["sample"]
    assert cb["one_core"]["cores"] == 1 and cb["one_core"]["value"] > 0          # both figures
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    _check(_one_json_line(res.stdout), 1)


def test_reference_arm_under_torchrun_only_rank0_works():
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    _check(_one_json_line(res.stdout), 2)
