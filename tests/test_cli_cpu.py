"""`nr-ray-tracer render` (csrc/cli_main.cpp) argument and error handling — everything that happens before the GPU is
touched, plus the refusal to run without one (no CPU fallback)."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nr_ray_tracer_b200", "bin", "nr-ray-tracer")
QUADS = os.path.join(ROOT, "scenes", "quads.toml")


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([CLI, *args], capture_output=True, text=True, cwd=ROOT, env=e, timeout=120)


def test_usage_and_argument_errors(tmp_path):
    r = run("--help")
    assert r.returncode == 0 and "Usage: nr-ray-tracer render" in r.stdout and "--bvh reference|sah" in r.stdout
    assert run("render", "--help").returncode == 0
    for args, needle in [(("render",), "missing <SCENE>"),
                         (("render", QUADS, "-o", str(tmp_path / "x.bmp")), "unsupported output format"),
                         (("render", str(tmp_path / "nope.toml"), "-o", str(tmp_path / "a.png")), "cannot read"),
                         (("render", QUADS, "-W", "10", "-H", "10", "--aspect-ratio", "2", "-o", str(tmp_path / "b.png")),
                          "exactly two of"),
                         (("render", QUADS, "--bvh", "fast", "-o", str(tmp_path / "c.png")), "--bvh must be"),
                         (("render", QUADS, "--samples-per-pixel"), "missing value"),
                         (("frobnicate",), "unknown command")]:
        r = run(*args)
        assert r.returncode != 0 and needle in (r.stderr + r.stdout), (args, r.stderr, r.stdout)


def test_numeric_options_are_validated_like_clap(tmp_path):
    """Negative, overflowing, fractional or empty numbers are errors (clap's u32 / f64 / u64 parsers), an unknown
    option is reported without swallowing the argument after it, and nothing is written."""
    out = str(tmp_path / "n.png")
    for args, needle in [(("-W", "-1"), "invalid"), (("--width", "4294967296"), "invalid value"),
                         (("--samples-per-pixel", "1e3"), "invalid value"), (("--ray-max-bounces", ""), "invalid value"),
                         (("--field-of-view", "wide"), "invalid value"), (("--gamma-value", "x"), "--gamma-value"),
                         (("--seed", "-3"), "--seed"), (("--device", "zero"), "--device"), (("--gpus", "0"), "--gpus"),
                         (("--gpus", "two"), "--gpus"), (("--frobnicate", "-W"), "unknown option '--frobnicate'"),
                         (("-x",), "unknown option '-x'")]:
        r = run("render", QUADS, "-o", out, *args)
        assert r.returncode != 0 and needle in (r.stderr + r.stdout), (args, r.stderr, r.stdout)
        assert not os.path.exists(out)
    r = run("render", QUADS, "-o", out, env={"NR_RT_GPUS": "many"})
    assert r.returncode != 0 and "NR_RT_GPUS" in (r.stderr + r.stdout)
    r = run("render", QUADS, "-o", out, env={"NR_RT_CAMERA_WIDTH": "-5"})
    assert r.returncode != 0 and "NR_RT_CAMERA_WIDTH" in (r.stderr + r.stdout)


def test_output_file_is_not_overwritten_without_force(tmp_path):
    out = tmp_path / "keep.png"
    out.write_bytes(b"precious")
    r = run("render", QUADS, "-o", str(out))
    assert r.returncode != 0 and "use -f to overwrite" in (r.stderr + r.stdout)   # create_new (cli.rs:140-154)
    assert out.read_bytes() == b"precious"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_render_without_a_gpu_fails_loudly(tmp_path):
    r = run("render", QUADS, "-o", str(tmp_path / "q.png"), "-W", "16", "-H", "9")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_create_subcommand_hands_over_to_the_generators(tmp_path):
    """The binary keeps the reference's two subcommands: `create` runs the Python generators (create.py)."""
    from nr_ray_tracer_b200 import create
    out = tmp_path / "cube.json"
    r = run("create", "cube", "-o", str(out), env={"NRRT_PYTHON": __import__("sys").executable})
    assert r.returncode == 0, r.stderr
    assert out.read_text() == create.dumps(create.cube(), "json") + "\n"
    r = run("create", "quads", env={"NRRT_PYTHON": __import__("sys").executable})
    assert r.returncode == 0 and r.stdout == create.dumps(create.quads(), "toml") + "\n"
    assert run("create", "nonsense", env={"NRRT_PYTHON": __import__("sys").executable}).returncode != 0
