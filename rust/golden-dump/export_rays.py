#!/usr/bin/env python3
"""tests/golden/golden.npz -> tests/golden/rust_dump/<scene>.rays.bin (n x 6 little-endian f64) plus the sample points
of the texture dump, for the Rust harness (which should not need an .npz reader)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.make_golden import GOLDEN_SCENES  # noqa: E402
from tests.test_oracle import rust_dump_points  # noqa: E402

out = os.path.join(ROOT, "tests", "golden", "rust_dump")
os.makedirs(out, exist_ok=True)
G = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
for name in GOLDEN_SCENES:
    key = name.split(".")[0].replace("-", "_")
    rays = np.ascontiguousarray(G[f"{key}__rays"], dtype="<f8")
    rays.tofile(os.path.join(out, f"{name}.rays.bin"))
    print(name, rays.shape)
np.ascontiguousarray(rust_dump_points(), dtype="<f8").tofile(os.path.join(out, "texture_points.bin"))
