"""bench.py contract checks that need no GPU: the reference arm's JSON line, rank handling, and the refusal of the
product arm to produce a number without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--width", "48", "--height", "27", "--spp", "4", "--ref-spp", "2", "--depth", "5"]


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1"] + SMALL)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "cornell-box-scene.json 48x27 4spp depth5" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"] + SMALL,
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = run(["--steps", "1", "--warmup", "1", "--no-cpu"] + SMALL)
    assert r.returncode != 0
    assert not any(ln.startswith("{") and '"value"' in ln for ln in r.stdout.splitlines())
