#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counts to CUDA source lines (developer tool).

ncu's CSV source page is SASS-only; this joins it (in instruction order) with `nvdisasm -g` line info of the
same kernel in the in-tree library, and prints the hottest source lines.
Usage: ncu_lines.py rep.ncu-rep <kernel-regex> [top_n] [mangled-regex]
The optional mangled-regex picks the template instantiation in the library (e.g. k_render_fusedILj6ELi4E) when
the demangled kernel regex given to ncu matches several.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nr_ray_tracer_b200", "libnrrt_b200.so")


def run(args, cwd=None):
    return subprocess.run(args, capture_output=True, text=True, cwd=cwd).stdout


def disasm_lines(kernel_regex):
    """[(opcode text, file, line)] for the first function whose name matches."""
    tmp = tempfile.mkdtemp()
    run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp)
    cubin = [f for f in os.listdir(tmp) if f.startswith("nrrt_device.")][0]
    text = run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)])
    out, cur_fn, take, file, line = [], None, False, "?", 0
    for ln in text.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            take = re.search(kernel_regex, cur_fn) is not None and not out
            continue
        if not take:
            if out and re.match(r"\s*\.section", ln):
                break
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            file, line = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(1).strip(), file, line))
    return out


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    src = list(csv.reader(run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}",
                               "--launch-skip", "0", "--launch-count", "1"]).splitlines()))
    h = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
    sh = src[h]
    si, ii, ti, wi = sh.index("Source"), sh.index("Instructions Executed"), sh.index(
        "Thread Instructions Executed"), sh.index("# Samples")
    rows = []
    for r in src[h + 1:]:
        try:
            rows.append((r[si].strip(), int(r[ii]), int(r[ti]), int(r[wi])))
        except (ValueError, IndexError):
            pass
    dis = disasm_lines(sys.argv[4] if len(sys.argv) > 4 else kre)
    if len(dis) != len(rows):
        print(f"warning: {len(rows)} profiled instructions vs {len(dis)} disassembled (library rebuilt since the "
              f"profile?) — attribution by position may be off", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (txt, n, t, s), (_op, f, l) in zip(rows, dis):
        a = agg[(f, l)]
        a[0] += n
        a[1] += t
        a[2] += s
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    cache = {}

    def source(f, l):
        for d in ("nr_ray_tracer_b200/csrc", "include"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in cache:
                    cache[p] = open(p, errors="ignore").read().splitlines()
                return cache[p][l - 1].strip()[:100] if 0 < l <= len(cache[p]) else ""
        return ""
    print(f"total warp instructions {tot}, stall samples {tots}")
    print(f"{'%inst':>6} {'thr':>5} {'%stall':>6}  location")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0] / tot * 100:6.2f} {a[1] / max(a[0], 1):5.1f} {a[2] / tots * 100:6.2f}  {f}:{l}  {source(f, l)}")


if __name__ == "__main__":
    main()
