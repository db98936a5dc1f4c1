"""Mutation fuzz of the native scene loader (csrc/scene_loader.cpp: own JSON and TOML parsers) against the Python
one (json / tomllib, i.e. parsers as strict as the reference's serde_json / toml): corrupted files must be rejected
with an error — never a crash — and the two loaders must agree on what is acceptable and on the scene it describes."""
import os
import random

import pytest

from nr_ray_tracer_b200 import api, create
from nr_ray_tracer_b200.scene_config import load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {(kind, fmt): create.dumps(create.GENERATORS[kind](), fmt)
        for kind in ("cornell-box", "noise", "quads", "simple-lights", "earth") for fmt in ("json", "toml")}
TOKENS = b'{}[]",:=.-0123456789eE \n\tabcxyz\\#\''


def _native(path):
    try:
        return api.NativeScene(path, base_dir=ROOT), ""
    except api.NrrtError as e:
        return None, str(e)


def _python(path):
    try:
        return load_scene(path, base_dir=ROOT), ""
    except Exception as e:   # SceneError, or a ValueError from a field conversion
        return None, f"{type(e).__name__}: {e}"


@pytest.mark.parametrize("seed", [1, 2])
def test_gentle_mutations_loaders_agree(seed, tmp_path):
    rng = random.Random(seed)
    disagreements = []
    accepted = 0
    for it in range(250):
        (kind, fmt), text = rng.choice(sorted(BASE.items()))
        b = bytearray(text.encode())
        for _ in range(rng.randint(1, 3)):
            pos = rng.randrange(len(b))
            if chr(b[pos]).isdigit():
                b[pos] = ord(rng.choice("0123456789"))
            elif rng.random() < 0.3:
                b[pos:pos] = rng.choice([b" ", b"\n", b"0", b"-", b".5", b"e2", b","])
            elif rng.random() < 0.3:
                del b[pos]
        path = str(tmp_path / f"f{it}.{fmt}")
        with open(path, "wb") as f:
            f.write(bytes(b))
        (ns, nmsg), (g, pmsg) = _native(path), _python(path)
        if (ns is None) != (g is None):
            disagreements.append((bytes(b)[:0], kind, fmt, it, nmsg[:80], pmsg[:80]))
        elif ns is not None:
            accepted += 1
            try:
                a, c = api.HostScene(ns).desc, api.HostScene(g).desc
            except api.NrrtError:
                continue
            assert (a.n_nodes, a.n_spheres, a.n_planes, a.n_materials, a.n_textures) == \
                   (c.n_nodes, c.n_spheres, c.n_planes, c.n_materials, c.n_textures), (kind, fmt, it)
    assert not disagreements, disagreements[:5]
    assert accepted > 20   # the mutations are gentle enough to leave many files well-formed


def test_heavy_corruption_is_rejected_not_crashed(tmp_path):
    rng = random.Random(7)
    rejected = 0
    for it in range(300):
        (kind, fmt), text = rng.choice(sorted(BASE.items()))
        b = bytearray(text.encode())
        for _ in range(rng.randint(1, 6)):
            op, pos = rng.random(), rng.randrange(len(b))
            if op < 0.4:
                b[pos] = rng.choice(TOKENS)
            elif op < 0.7:
                del b[pos:pos + rng.randint(1, 20)]
            else:
                b[pos:pos] = bytes(rng.choice(TOKENS) for _ in range(rng.randint(1, 8)))
        path = str(tmp_path / f"h{it}.{fmt}")
        with open(path, "wb") as f:
            f.write(bytes(b))
        ns, msg = _native(path)
        if ns is None:
            assert msg          # an error message, not a silent failure
            rejected += 1
        else:
            try:
                api.HostScene(ns)
            except api.NrrtError:
                pass
    assert rejected > 200


def test_recursive_scene_includes_are_an_error_not_a_crash(tmp_path):
    """A scene that includes itself — directly or through another file — must come back as a loader error from both
    loaders (the reference recurses until its stack overflows, scene_config.rs:335-340; the native loader must not take
    the process down with it)."""
    from nr_ray_tracer_b200 import api
    from nr_ray_tracer_b200.scene_config import SceneError, load_scene
    a, b = tmp_path / "a.json", tmp_path / "b.json"
    a.write_text('{"camera": {}, "scene": [{"Scene": {"path": "b.json"}}]}')
    b.write_text('{"camera": {}, "scene": [{"Scene": {"path": "a.json"}}]}')
    selfinc = tmp_path / "self.json"
    selfinc.write_text('{"camera": {}, "scene": [{"Sphere": {"center": [0,0,0], "radius": 1}}, {"Scene": {"path": "self.json"}}]}')
    for f in (a, selfinc):
        with pytest.raises(api.NrrtError) as e:
            api.NativeScene(str(f), base_dir=str(tmp_path))
        assert "include" in str(e.value)
        with pytest.raises(SceneError):
            load_scene(str(f), base_dir=str(tmp_path))
    # a legitimate diamond (two includes of the same file) still loads
    leaf = tmp_path / "leaf.json"
    leaf.write_text('{"camera": {}, "scene": [{"Sphere": {"center": [0,0,0], "radius": 1}}]}')
    top = tmp_path / "top.json"
    top.write_text('{"camera": {}, "scene": [{"Scene": {"path": "leaf.json"}}, {"Translate": {"offset": [3,0,0], '
                   '"object": {"Scene": {"path": "leaf.json"}}}}]}')
    ns = api.NativeScene(str(top), base_dir=str(tmp_path))
    assert ns.graph.n_objects >= 4
    assert load_scene(str(top), base_dir=str(tmp_path)).count_primitives() == 2
