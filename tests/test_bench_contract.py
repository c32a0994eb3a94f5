"""bench.py's contract, as far as it can be checked without a GPU: the workload tables are consistent, both arms describe a
workload with the same `config` object, and the reference arm runs here (it needs no GPU) and prints one JSON line with the keys the
driver reads."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_tables_are_consistent():
    b = _bench()
    assert set(b.SECONDARY) <= set(b.WORKLOADS) and set(b.E2E_SECONDARY) <= set(b.SECONDARY)
    assert len(set(b.SECONDARY)) == len(b.SECONDARY) and "fft4096_f32" not in b.SECONDARY  # the top-level workload is not repeated
    for name, spec in b.WORKLOADS.items():
        assert spec["kind"] in ("fft", "iir", "pipeline"), name
        assert spec["precision"] in ("f32", "f64") and spec["bytes_per_sample"] in (8, 12, 16, 20, 32), name
        cfg = b.config_of(name, spec)
        assert cfg["workload"] == name and "model" not in cfg and "kind" not in cfg
        assert cfg == b.config_of(name, dict(spec)) and json.loads(json.dumps(cfg)) == cfg  # deterministic, serialisable
    for name in b.NCU_TRAFFIC:
        assert name in b.WORKLOADS, name
    # every BASELINE config is on the default line: config 2 (top level + fp64), 3, 4, 5
    for must in ("fft4096_f64", "iir16384_f32", "iirscan_f64", "pipeline_cfg5_f32", "fft65536_f32"):
        assert must in b.SECONDARY


@pytest.mark.parametrize("workload", ["fft4096_f32", "iir16384_f32"])
def test_reference_arm_prints_the_contract_line(workload):
    from oracle import oracle as O

    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    b = _bench()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", workload],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"] == b.config_of(workload, b.WORKLOADS[workload])  # the same object the b200 arm emits
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_exits_quietly():
    """Under torchrun only rank 0 times the reference; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
