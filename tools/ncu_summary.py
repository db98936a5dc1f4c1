#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + SASS opcode mix) as markdown.  Usage: ncu_summary.py rep.ncu-rep [title]"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    rows = list(csv.reader(run(['ncu', '-i', rep, '--page', 'raw', '--csv']).splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = rows[2:]
    out = [f"# {title}\n", f"Source: `{rep.split('/')[-1]}` (ncu --set full --clock-control none)\n"]
    names = [r[idx['Kernel Name']].split('(')[0] for r in launches]
    out.append("| metric | unit | " + " | ".join(names) + " |")
    out.append("|---|---|" + "---|" * len(names))
    for k in KEYS:
        if k in idx:
            out.append(f"| {k} | {rows[1][idx[k]]} | " + " | ".join(r[idx[k]] for r in launches) + " |")
    seen = set()
    for li, name in enumerate(names):
        if name in seen:
            continue
        seen.add(name)
        src = list(csv.reader(run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{name}',
                                   '--launch-skip', '0', '--launch-count', '1']).splitlines()))
        h = None
        for i, r in enumerate(src):
            if r and r[0] == 'Address':
                h = i
                break
        if h is None:
            continue
        sh = src[h]
        si, ii, ti, wi = sh.index('Source'), sh.index('Instructions Executed'), sh.index(
            'Thread Instructions Executed'), sh.index('# Samples')
        ops, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
        tot = 0
        for r in src[h + 1:]:
            try:
                n, t, s = int(r[ii]), int(r[ti]), int(r[wi])
            except (ValueError, IndexError):
                continue
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
            op = m.group(2).split('.')[0] if m else '?'
            ops[op] += n
            thr[op] += t
            smp[op] += s
            tot += n
        out.append(f"\n## SASS opcode mix: {name} (first profiled launch, {tot} warp instructions)\n")
        out.append("| opcode | % of warp instr | avg active threads | % of stall samples |")
        out.append("|---|---|---|---|")
        st = max(sum(smp.values()), 1)
        for op, n in ops.most_common(22):
            out.append(f"| {op} | {n / tot * 100:.1f} | {thr[op] / max(n, 1):.1f} | {smp[op] / st * 100:.1f} |")
    print("\n".join(out))


if __name__ == "__main__":
    main()
