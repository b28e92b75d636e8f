"""Per-source-line view of an `ncu --set full --import-source on` capture, without the GUI.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [--top 40] [--file render_stream.cu]

`ncu --page source --csv` lists SASS instructions with executed counts and stall samples but no line numbers; `nvdisasm -g` lists the same
instructions of the cubin with `//## File "...", line N` markers.  Both are in program order, so they are joined by position (the opcode
text is cross-checked) and aggregated per source line: warp instructions executed, stall samples by reason."""
import argparse
import collections
import csv
import re
import subprocess
import sys


def sass_with_lines(cubin, kernel, outer=False):
    out = subprocess.run(['nvdisasm', '-gi' if outer else '-g', '-c', cubin], capture_output=True, text=True, check=True).stdout.splitlines()
    res, cur, active = [], (None, 0), False
    for ln in out:
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel in m.group(1)
            continue
        if ln.startswith('\t.section') or re.match(r'\s*\.section', ln):
            active = False if '.text.' not in ln else active
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:                                  # with -gi an inlined instruction carries its whole chain; the LAST marker is the outermost call site
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep'); ap.add_argument('cubin'); ap.add_argument('kernel')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--file', default=None)
    ap.add_argument('--outer', action='store_true', help='attribute inlined code to its outermost call site (nvdisasm -gi)')
    ap.add_argument('--launch', type=int, default=0, help='index of the kernel launch inside the report')
    args = ap.parse_args()
    txt = subprocess.run(['ncu', '-i', args.rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(txt.splitlines()):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(row)
    blocks = [b for b in blocks if args.kernel in b['name']]
    if not blocks:
        sys.exit('kernel not in report')
    b = blocks[min(args.launch, len(blocks) - 1)]
    hdr, rows = b['rows'][0], b['rows'][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    sass = sass_with_lines(args.cubin, args.kernel.split('<')[0].split('(')[0], args.outer)
    if len(sass) != len(rows):
        print(f'warning: {len(sass)} SASS instructions in the cubin vs {len(rows)} in the report (joined by position up to the shorter)', file=sys.stderr)
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for (addr, op, loc), r in zip(sass, rows):
        a = agg[loc]
        a['inst'] += int(r[ix['Instructions Executed']] or 0)
        a['samples'] += int(r[ix['# Samples']] or 0)
        for h in reasons:
            a[h] += int(r[ix[h]] or 0)
        tot['inst'] += int(r[ix['Instructions Executed']] or 0); tot['samples'] += int(r[ix['# Samples']] or 0)
    print(f"kernel {b['name'][:100]}  total warp-instructions {tot['inst']}  samples {tot['samples']}")
    items = [(loc, a) for loc, a in agg.items() if args.file is None or loc[0] == args.file]
    print('--- by stall samples'); show(sorted(items, key=lambda x: -x[1]['samples'])[:args.top], tot, reasons)
    print('--- by instructions executed'); show(sorted(items, key=lambda x: -x[1]['inst'])[:args.top], tot, reasons)


def show(items, tot, reasons):
    for loc, a in items:
        top = sorted(((a[h], h) for h in reasons if a[h]), reverse=True)[:3]
        print(f"{loc[0]}:{loc[1]:<5} inst {a['inst']:>10} ({100.0 * a['inst'] / max(tot['inst'], 1):5.1f}%)  samples {a['samples']:>7} "
              f"({100.0 * a['samples'] / max(tot['samples'], 1):5.1f}%)  " + ' '.join(f'{h[6:]}={n}' for n, h in top))


if __name__ == '__main__':
    main()
