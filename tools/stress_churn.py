"""Repeats the calm / churn runs of tests/test_gpu_batch.py::test_batch_groups_survive_rooms_leaving_and_rejoining and says
where (key, first differing tick) a run departs from the first calm run. usage: stress_churn.py <iterations> [tickers]"""
import os, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
def run(tag, churn, tickers, d):
    out = Path(d) / f"{tag}.npz"
    cmd = [sys.executable, str(ROOT / "tests" / "graph_runner.py"), "--streams", "12", "--pins", "4", "--ticks", "90",
           "--tickers", str(tickers), "--dump", str(out)] + (["--churn", churn] if churn else [])
    r = subprocess.run(cmd, env=dict(os.environ, MSB200_BATCH="16"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(np.load(out))
n, tickers = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 1
with tempfile.TemporaryDirectory() as d:
    base = run("calm0", "", tickers, d)
    bad = 0
    for it in range(n):
        for tag, churn in (("calm", ""), ("churn", "30,55")):
            got = run(f"{tag}{it}", churn, tickers, d)
            for i in range(8):
                for key in (f"spk{i}", f"out{i}"):
                    a, b = base[key], got[key]
                    if len(a) != len(b) or not np.array_equal(a, b):
                        m = min(len(a), len(b))
                        first = int(np.argmax(a[:m] != b[:m])) if (a[:m] != b[:m]).any() else m
                        print(f"iter {it} {tag} {key}: len {len(a)} vs {len(b)}, first difference at sample {first} (tick {first // 480}), "
                              f"max |d| {int(np.abs(a[:m].astype(int) - b[:m].astype(int)).max())}", flush=True)
                        bad += 1
    print(f"done: {n} iterations x (calm, churn), {bad} departures from the first calm run (AEC path {os.environ.get('MSB200_AEC_PATH', '0')})")
