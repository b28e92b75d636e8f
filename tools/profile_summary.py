"""Turns the captures of tools/gpu_profile.sh (gpurun_out/prof_*) into the committed evidence: copies the reports and the launch list into profiles/,
prints the per-launch SR table, the step breakdown, and rewrites profiles/ncu_traffic.json (keyed to the current sr_tc.cu hash and commit)."""
import csv
import hashlib
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def raw(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    for src, dst in (('prof_sr.ncu-rep', 'r2_sr.ncu-rep'), ('prof_render_stream.ncu-rep', 'r2_render_stream.ncu-rep'), ('prof_sample.ncu-rep', 'r2_sample.ncu-rep'),
                     ('prof_launches.csv', 'r2_launches_warmcache_nograph.csv')):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    hdr, units, rows = raw(os.path.join(P, 'r2_sr.ncu-rep'))
    ix = {k: i for i, k in enumerate(hdr)}
    keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max.per_second']
    tot, nconv = 0.0, 0
    print('| kernel | time us | tensor % | DRAM rd MB | DRAM wr MB | DRAM % | L2 % | issue % | GHz |')
    for r in rows:
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
        v = [float(r[ix[k]]) for k in keys]
        print(f'| {name} | ' + ' | '.join(f'{x:.1f}' if i != 7 else f'{x:.2f}' for i, x in enumerate(v)) + ' |')
        if 'conv_tc3' in name:
            tot += v[2] + v[3]; nconv += 1
    sha = hashlib.sha256(open(os.path.join(ROOT, 'real3dportrait_b200', 'csrc', 'sr_tc.cu'), 'rb').read()).hexdigest()[:16]
    commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    d = {'sr_tc_cu_sha16': sha, 'commit': commit, 'bytes_per_launch': tot * 1e6 / max(nconv, 1),
         'what': f'dram__bytes_read.sum + dram__bytes_write.sum averaged over the {nconv} conv_tc3_kernel launches of one step (batch 4, {tot:.1f} MB in total), '
                 'ncu --set full --clock-control none',
         'file': 'profiles/r2_sr.ncu-rep (summary: profiles/r2_sr_ncu.md)'}
    json.dump(d, open(os.path.join(P, 'ncu_traffic.json'), 'w'), indent=1)
    print(json.dumps(d))
    rows = list(csv.reader(open(os.path.join(P, 'r2_launches_warmcache_nograph.csv'))))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    ix = {k: i for i, k in enumerate(rows[h])}
    data = rows[h + 1:]
    names = [r[ix['Kernel Name']][:60] for r in data]
    vals = [float(r[ix['Metric Value']]) / 1000 for r in data]
    starts = [i for i, n in enumerate(names) if 'mlp_to_tc' in n]
    s = starts[4] - 1
    print('one steady-state step (us):')
    t = 0.0
    for n, v in zip(names[s:s + 12], vals[s:s + 12]):
        print(f'  {v:8.1f}  {n}'); t += v
    print(f'  {t:8.1f}  total')
    for rep, kern in (('r2_render_stream.ncu-rep', 'render_stream'), ('r2_sample.ncu-rep', 'triplane_sample')):
        hdr, units, rws = raw(os.path.join(P, rep))
        ix = {k: i for i, k in enumerate(hdr)}
        print(kern, {k.split('.')[0]: rws[0][ix[k]] for k in ('gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                                                               'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                                               'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio') if k in ix})


if __name__ == '__main__':
    main()
