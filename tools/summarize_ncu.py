"""Turn gpurun_out ncu artefacts into the committed summaries under profiles/.

    python tools/summarize_ncu.py --tag r1 --launches gpurun_out/launches_r1.csv --rep gpurun_out/k2_r1.ncu-rep \
        --workload C3 --dtype f64 --nt 16 --cmd "python bench.py --steps 2 --warmup 1 --nt-local 16 --no-cpu --no-e2e"
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

KEEP = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'smsp__cycles_active.avg',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tag', required=True)
    ap.add_argument('--launches')
    ap.add_argument('--rep')
    ap.add_argument('--kernel', default='k2')
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--dtype', default='f64')
    ap.add_argument('--nt', type=int, default=16)
    ap.add_argument('--cmd', default='')
    ap.add_argument('--fused', action='store_true')
    a = ap.parse_args()
    os.makedirs('profiles', exist_ok=True)
    md = [f'# ncu summary {a.tag}', '', f'command: `{a.cmd}` (one B200, --clock-control none)', '']
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 5]
        hdr, rows = rows[0], rows[1:]
        ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
        agg = collections.OrderedDict()
        for r in rows:
            name = re.sub(r'\(.*', '', r[ki])
            name = re.sub(r'<unnamed>::', '', name)[:90]
            x = agg.setdefault(name, [0, 0.0])
            x[0] += 1
            x[1] += float(r[vi].replace(',', ''))
        step = {k: v for k, v in agg.items() if 'k2_edgeflux' in k or 'k3_integrate' in k or 'k23_fused' in k
                or 'k_reduce_subrows' in k}
        tot_step = sum(v[1] for v in step.values())
        md += ['## launch list (gpu__time_duration.sum, ns; cold-cache, serialised: compare shares)', '',
               '| kernel | launches | total ns | ns/launch | share of the step kernels |', '|---|---|---|---|---|']
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sh = f'{100 * t / tot_step:.1f} %' if k in step else ''
            md.append(f'| `{k}` | {c} | {t:.0f} | {t / c:.0f} | {sh} |')
        md += ['', 'step kernels = K2 + K3 (what bench.py times); the torch elementwise kernels are the synthetic-input '
               'generator, the other nfx kernels are the one-off locator / computeWeights (K1).', '']
        open(f'profiles/{a.tag}_launches.csv', 'w').write(open(a.launches).read())
    if a.rep:
        out = subprocess.run(['ncu', '-i', a.rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        md += [f'## ncu --set full, {a.kernel} (per launch)', '']
        traffic = []
        for r in rows:
            md.append(f'kernel `{r[hdr.index("Kernel Name")][:100]}`')
            md += ['', '| metric | value | unit |', '|---|---|---|']
            for k in KEEP:
                if k in hdr:
                    md.append(f'| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |')
            md.append('')

            def tobytes(key):
                v = float(r[hdr.index(key)].replace(',', ''))
                u = units[hdr.index(key)].lower()
                return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}[u]
            traffic.append(tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'))
        with open(f'profiles/{a.tag}_{a.kernel}_raw.csv', 'w') as f:
            w = csv.writer(f)
            keep_idx = [i for i, h in enumerate(hdr) if h in KEEP or h in ('Kernel Name', 'ID')]
            for r in [hdr, units] + rows:
                w.writerow([r[i] for i in keep_idx])
        if a.kernel == 'k2':
            per_step = sum(traffic) / len(traffic) / a.nt
            md += [f'DRAM traffic per launch (read+write, mean of {len(traffic)} launches): '
                   f'{sum(traffic) / len(traffic):.4e} B for {a.nt} time steps = {per_step:.4e} B per time step.', '']
            tpath = 'profiles/k2_traffic.json'
            tr = json.load(open(tpath)) if os.path.exists(tpath) else {}
            key = f'{a.workload}_{a.dtype}' + ('_fused' if a.fused else '')
            tr[key] = {'dram_bytes_per_timestep': per_step,
                       'source': f'profiles/{a.tag}_k2_raw.csv (ncu --set full, nt={a.nt})'}
            json.dump(tr, open(tpath, 'w'), indent=1)
    open(f'profiles/{a.tag}_ncu_summary.md', 'w').write('\n'.join(md) + '\n')
    print('\n'.join(md))


if __name__ == '__main__':
    main()
