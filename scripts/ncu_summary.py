"""Summaries of ncu captures for profiles/ (run in the build container on the .ncu-rep / .csv files gpurun brought back).

    python scripts/ncu_summary.py full  gpurun_out/<name>.ncu-rep  profiles/<out>.json [--edges-per-launch E --version "<dndm_version>"]
    python scripts/ncu_summary.py list  gpurun_out/<launches>.csv  profiles/<out>_summary.csv

``full``: the metrics the roofline discussion uses, per profiled launch (raw page).  With --edges-per-launch it also writes
profiles/gcl_traffic.json -- DRAM bytes per edge of edge_pair_kernel<GCL> -- which bench.py reads for ``roofline.traffic``.
``list``: per kernel name the launch count, total and share of device time (ncu --metrics gpu__time_duration.sum)."""
import csv
import json
import os
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']


def full(rep, out, edges=None, version=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        o = {'kernel': d['Kernel Name'], 'grid': d.get('Grid Size'), 'block': d.get('Block Size')}
        for m in METRICS:
            if m in d:
                o[m] = float(d[m].replace(',', '')) if d[m] else None
                o[m + ' [unit]'] = units[hdr.index(m)]
        res.append(o)
    json.dump({'source': os.path.basename(rep), 'command': 'ncu --set full --clock-control none --import-source on', 'launches': res},
              open(out, 'w'), indent=1)
    if edges:
        gcl = [o for o in res if 'edge_pair_kernel<1' in o['kernel'] or 'edge_pair_kernel<(bool)1' in o['kernel']]
        if gcl:
            to_b = lambda o, k: o[k] * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}[o[k + ' [unit]']]
            tot = sum(to_b(o, 'dram__bytes_read.sum') + to_b(o, 'dram__bytes_write.sum') for o in gcl) / len(gcl)
            json.dump({'kernel_version': version, 'dram_bytes_per_edge': tot / float(edges), 'dram_bytes_per_launch': tot,
                       'edges_per_launch': float(edges), 'source': f'profiles/{os.path.basename(out)} (dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(gcl)} launches)'},
                      open(os.path.join(os.path.dirname(out), 'gcl_traffic.json'), 'w'), indent=1)


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    name_i, val_i = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= val_i:
            continue
        n = r[name_i].split('(')[0]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[val_i].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write('kernel,launches,total_us,share_pct\n')
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{n}",{c},{t / 1e3:.1f},{100 * t / tot:.1f}\n')


if __name__ == '__main__':
    if sys.argv[1] == 'full':
        kw = {}
        if '--edges-per-launch' in sys.argv:
            kw['edges'] = float(sys.argv[sys.argv.index('--edges-per-launch') + 1])
        if '--version' in sys.argv:
            kw['version'] = sys.argv[sys.argv.index('--version') + 1]
        full(sys.argv[2], sys.argv[3], **kw)
    else:
        launch_list(sys.argv[2], sys.argv[3])
