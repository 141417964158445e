"""bench.py --impl reference (the reference-side CPU arm the driver runs beside ours): contract of its JSON line, and under
a multi-rank launch rank 0 alone works and prints while the other ranks exit 0 silently. CPU only."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(rank, world):
    env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
               MASTER_PORT="29555")
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", str(world), "--steps", "2",
                           "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line_and_rank_gating():
    r0 = _run(0, 2)
    assert r0.returncode == 0, r0.stderr[-1500:]
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["steps"] == 2
    assert line["unit"] == "stream-ticks/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["gpu_launches"] == 0 and line["config"]["workload"].startswith("cfg2")
    r1 = _run(1, 2)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
