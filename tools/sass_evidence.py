"""Writes profiles/r2_sass_evidence.md: per kernel of libr3dp_b200.so, the counts of the SASS mnemonics that prove the Blackwell-native path
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async)
plus a short excerpt around the first tensor-core instruction of the three tensor-core kernels.  Runs without a GPU (cuobjdump)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'real3dportrait_b200', 'lib', 'libr3dp_b200.so')
PAT = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKPF', 'USETMAXREG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA2', 'MUFU', 'LDG.E.128', 'ELECT']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout.splitlines()
    funcs, cur = collections.OrderedDict(), None
    for ln in sass:
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    out = ['# r2 - SASS evidence (cuobjdump -sass real3dportrait_b200/lib/libr3dp_b200.so, built by __graft_entry__.build())', '',
           'Counts of instructions per kernel (`tools/sass_evidence.py`).  `UTCHMMA` = `tcgen05.mma` (`.2CTA` = `cta_group::2`), `LDTM` = `tcgen05.ld`,',
           '`UTMALDG` = `cp.async.bulk.tensor` (TMA), `UBLKPF` = `cp.async.bulk.prefetch.L2`, `USETMAXREG` = `setmaxnreg`, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, `LDGSTS` = `cp.async`; `HMMA` (legacy `mma.sync`) must be 0.', '',
           '| kernel | ' + ' | '.join(PAT) + ' |', '|---|' + '---:|' * len(PAT)]
    for name, lines in funcs.items():
        body = '\n'.join(lines)
        cnt = []
        for p in PAT:
            if p == 'UTCHMMA':
                cnt.append(len(re.findall(r'\bUTCHMMA\b', body)))
            elif p == 'HMMA':
                cnt.append(len(re.findall(r'\bHMMA\b', body)))
            else:
                cnt.append(body.count(p))
        if sum(cnt[:9]) or 'render' in name or 'conv' in name:
            out.append('| `' + name.replace('r3dp::', '') + '` | ' + ' | '.join(str(c) for c in cnt) + ' |')
    for key in ('conv_tc3_kernel<2>', 'render_stream_kernel<3, false>', 'fir_tma_kernel'):
        for name, lines in funcs.items():
            if key in name:
                idx = next((i for i, ln in enumerate(lines) if 'UTCHMMA' in ln or 'UTMALDG' in ln), None)
                if idx is not None:
                    out += ['', f'## `{name}`: first tensor-core / TMA instruction in context', '', '```']
                    out += [re.sub(r'/\*[0-9a-fx]+\*/\s*$', '', ln).rstrip() for ln in lines[max(0, idx - 6):idx + 10]]
                    out += ['```']
                break
    path = os.path.join(ROOT, 'profiles', 'r2_sass_evidence.md')
    open(path, 'w').write('\n'.join(out) + '\n')
    print(path, len(funcs), 'kernels')


if __name__ == '__main__':
    sys.exit(main())
