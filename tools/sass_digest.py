"""md5 of the device code (SASS listing of every sm_100a cubin) inside the built library.

Host-only changes must leave it unchanged: that is how a change made without a GPU at hand is shown not to touch what
was validated on the B200 (python tools/sass_digest.py before and after the rebuild)."""
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(lib: str) -> str:
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
        h = hashlib.md5()
        for f in sorted(os.listdir(d)):
            h.update(subprocess.run(["cuobjdump", "-sass", f], cwd=d, check=True, capture_output=True).stdout)
        return h.hexdigest()


if __name__ == "__main__":
    lib = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, "nr_ray_tracer_b200", "libnrrt_b200.so")
    print(digest(lib), lib)
