#!/usr/bin/env python
"""Static evidence per kernel, no GPU needed: registers / spills / shared memory from `ptxas -v`, and the SASS mnemonics of
the built library that show which hardware path a kernel uses (B200_PROFILING.md: `UTC*MMA` = tcgen05.mma, `LDTM`/`STTM` =
tcgen05.ld/st, `UTMALDG`/`UTMASTG`/`UBLKCP` = TMA, `HMMA` = legacy mma.sync).  Writes a markdown table to stdout.

    python scripts/static_report.py > profiles/r1_static_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xagents_b200 import _build  # noqa: E402

KEYS = ['UTC', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDG.E.128', 'STG.E.128', 'LDGSTS', 'SYNCS', 'ELECT', 'SHFL', 'DADD']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    short = {}
    for raw, full in zip(names, out):
        m = re.match(r'(?:void )?(?:\(anonymous namespace\)::)?([\w:]+)(<.*>)?\(', full)
        short[raw] = (m.group(1) + (m.group(2) or '')) if m else full
    return short


def ptxas_table():
    rows = {}
    for src in _build.SOURCES:
        path = os.path.join(_build.CSRC, src)
        cmd = [_build.nvcc_path(), *_build.ARCH_FLAGS, '-lineinfo', '-O3', '-std=c++17', '-Xptxas=-v', '-c', '-I', os.path.join(ROOT, 'include'),
               '-o', os.devnull, path]
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
        cur = None
        for line in err.splitlines():
            m = re.search(r"Compiling entry function '(\w+)'", line)
            if m:
                cur = m.group(1)
                rows[cur] = {'file': src}
            m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
            if m and cur:
                rows[cur].update(stack=int(m.group(1)), spill=int(m.group(2)) + int(m.group(3)))
            m = re.search(r'Used (\d+) registers', line)
            if m and cur:
                rows[cur]['regs'] = int(m.group(1))
                sm = re.search(r'(\d+) bytes smem', line)
                rows[cur]['smem'] = int(sm.group(1)) if sm else 0
    return rows


def sass_counts():
    lib = _build.build()
    text = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in text.splitlines():
        m = re.search(r'Function : (\w+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
        if m:
            op = m.group(1)
            counts[cur]['total'] += 1
            for key in KEYS:
                if op.startswith(key) or (key in ('LDG.E.128', 'STG.E.128') and op.startswith(key[:3]) and '.128' in op):
                    counts[cur][key] += 1
    return counts


def main():
    regs, sass = ptxas_table(), sass_counts()
    names = demangle(sorted(sass))
    print('# Static evidence per kernel (ptxas -v, cuobjdump -sass of the shipped library; no GPU involved)\n')
    print('`UTC*` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st (TMEM), `UTMALDG`/`UTMASTG` = TMA tensor copies, `UBLKCP` = TMA bulk copies,\n'
          '`SYNCS` = mbarrier operations, `HMMA` = legacy mma.sync (expected: 0 everywhere).  Spills are bytes of spill loads + stores.\n')
    print('| kernel | file | regs | static smem B | spill B | SASS instr | ' + ' | '.join(KEYS) + ' |')
    print('|---|---|---|---|---|---|' + '---|' * len(KEYS))
    for raw in sorted(sass, key=lambda r: (regs.get(r, {}).get('file', ''), names[r])):
        r, c = regs.get(raw, {}), sass[raw]
        print(f"| `{names[raw][:110]}` | {r.get('file', '?')} | {r.get('regs', '?')} | {r.get('smem', '?')} | {r.get('spill', '?')} | {c['total']} | "
              + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')


if __name__ == '__main__':
    main()
