#!/usr/bin/env python
"""Attribute an ncu source-page CSV (SASS level) to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv localuf_b200/csrc/libluf_b200.so k_mc_uf [--top 40]

ncu's own CUDA view only shows the .cu file; the kernel phases live in an
inlined header, so the line table is taken from nvdisasm and joined with ncu's
per-instruction counters by instruction order.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def section_tables(so):
    """{section name: [(file, line) per SASS instruction]} for every .text section of the library's cubin."""
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    tables, cur, name = {}, None, None
    for ln in txt.splitlines():
        m = re.match(r'\s*//-+ \.text\.(\S+)', ln)
        if m:
            name = m.group(1)
            tables[name] = []
            cur = None
            continue
        if name is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
            tables[name].append(cur)
    return tables


def line_table(so, kernel, n_instructions=None):
    """The line table of ONE kernel instance: the .text section whose name contains `kernel` (a mangled
    instance name such as k_mc_ufILi0E picks a template instance) and, when `n_instructions` is given (the
    number of SASS rows ncu profiled), whose instruction count equals it -- several template instances
    share a name prefix, and joining ncu's rows with another instance's table mislabels every row."""
    cands = {n: t for n, t in section_tables(so).items() if kernel in n}
    if n_instructions is not None:
        exact = {n: t for n, t in cands.items() if len(t) == n_instructions}
        if exact:
            cands = exact
        elif cands:
            print(f'warning: no instance of {kernel} has {n_instructions} instructions '
                  f'({ {n: len(t) for n, t in cands.items()} }); the library differs from the profiled one', file=sys.stderr)
    if not cands:
        raise SystemExit(f'no .text section matches {kernel}')
    if len(cands) > 1:
        print(f'warning: {len(cands)} sections match {kernel}: {sorted(cands)}; taking the shortest name', file=sys.stderr)
    name = min(cands, key=len)
    print(f'[ncu_lines] section .text.{name}: {len(cands[name])} instructions', file=sys.stderr)
    return cands[name]


def main():
    src_csv, so, kernel = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    table = line_table(so, kernel, len(body))
    if len(table) != len(body):
        print(f'warning: {len(table)} disassembled vs {len(body)} profiled instructions', file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for k, r in enumerate(body):
        key = table[k] if k < len(table) else ('?', 0)
        vals = (int(r[col['Instructions Executed']] or 0), int(r[col['Thread Instructions Executed']] or 0),
                int(r[col['# Samples']] or 0))
        for j in range(3):
            agg[key][j] += vals[j]; tot[j] += vals[j]
    print(f'total: warp-inst {tot[0]:,}  thread-inst {tot[1]:,} (avg {tot[1] / max(tot[0], 1):.1f} lanes)  samples {tot[2]:,}')
    print(f'{"file:line":34s} {"warp-inst%":>10s} {"lanes":>6s} {"samples%":>9s}')
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f'{str(key[0]) + ":" + str(key[1]):34s} {100 * v[0] / tot[0]:10.2f} {v[1] / max(v[0], 1):6.1f} {100 * v[2] / max(tot[2], 1):9.2f}')


if __name__ == '__main__':
    main()
