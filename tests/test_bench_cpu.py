"""bench.py's contract where it can be checked without a GPU: the reference arm (the reference's ATen CPU path through
oracle/esr_oracle_torch.py) prints ONE JSON line with the keys the driver reads, on the same `config` the B200 arm
reports, and the workload presets follow BASELINE.json's configs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--size", "48", "40"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


def test_workload_presets_follow_baseline_json():
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert sorted(bench.CONFIGS) == list(range(len(base["configs"]))) == [0, 1, 2, 3, 4]
    models = [c[0] for _, c in sorted(bench.CONFIGS.items())]
    assert models == ["imdn", "rfdn", "rfdn", "rlfn", "bsrn"]
    assert bench.CONFIGS[1][1] == 1 and tuple(bench.CONFIGS[1][2]) == (256, 256) and bench.CONFIGS[1][3] == "f16"
    assert bench.CONFIGS[2][2] == "div2k" and bench.CONFIGS[3][1] == 8 and bench.CONFIGS[4][1] == 16 and tuple(bench.CONFIGS[4][2]) == (270, 480)
    shapes = bench.div2k_shapes(32)
    assert len(shapes) == 32 and all(max(s) == 510 for s in shapes)
