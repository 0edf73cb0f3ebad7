"""CPU: the JSON line bench.py prints (checked on the reference arm of the quickest workload, which needs no GPU) carries
the keys the round driver reads, and the argument parser accepts every workload the docs name."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "raw_histogram"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["value"] > 0 and d["gpu_launches"] == 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and "workload" in d["config"]


def test_every_documented_workload_parses():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for w in ("pretrain", "finetune", "histogram", "event_pipeline", "raw_histogram"):
        assert f'"{w}"' in src
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "nonsense"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT)
    assert out.returncode != 0 and "invalid choice" in out.stderr
