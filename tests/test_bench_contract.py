"""bench.py bookkeeping that does not need a GPU: the algorithmic bytes of the decode loop are SURVEY.md §8d's figures,
and the reference arm's JSON line carries the contract's keys."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_decode_bytes_match_survey():
    b = _bench()
    # 21 002 572 B of weights per step + 145 728 B per clip (T=29, minT=4); C2 = 25.67 MB/step, 7.70 GB per 300 steps
    assert b.algorithmic_decode_bytes(32) == 300 * (21_002_572 + 32 * 145_728)
    assert abs(b.algorithmic_decode_bytes(32) / 300 / 1e6 - 25.67) < 0.01
    assert b.algorithmic_decode_bytes(1) == 300 * (21_002_572 + 145_728)
    # T=75, minT=10: 346 432 B per clip
    assert b.algorithmic_decode_bytes(32, t=75, min_t=10) == 300 * (21_002_572 + 32 * 346_432)


def test_committed_bench_line_has_contract_keys():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r1g_bench_line.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert line["gpu_launches"] > 0 and "workload" in line["config"] and "model" not in line["config"]
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9


def test_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU port of the reference's path, timed on host cores) needs no GPU and prints the
    contract's JSON line with impl / cpu_baseline / zero-copy e2e."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "mel-frames/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["value"] > 0
