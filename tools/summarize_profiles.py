# -*- coding: utf-8 -*-
''' Turn the raw artefacts of one scripts/gpu_round.sh session (gpurun_out/*_<tag>.*) into the
    small, tracked summaries under profiles/:

        python tools/summarize_profiles.py <tag>

    profiles/<tag>_bench.json            the bench line (own arm) and the reference-arm line
    profiles/<tag>_launches.txt          ncu launch list of the bench command, aggregated per kernel
    profiles/<tag>_ncu_c1_integrate.txt  key counters + stall reasons of the integrator (ncu --set full, C1)
    profiles/<tag>_ncu_c2_counters.txt   DRAM traffic / FP64 pipe / lane efficiency of the integrator on C2
    profiles/<tag>_sass_profile.txt      executed instructions per source function (ncu source page)
'''
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')


def launches(tag):
    path = os.path.join(OUT, f'launches_{tag}.csv')
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name, v = r[ik][:70], float(r[iv].replace(',', ''))
        if name.startswith('sonic_integrate_kernel') and v < 1e6:
            name = 'sonic_integrate_kernel [placement probe, returns at once]'
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, f'{tag}_launches.txt'), 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n')
        fh.write('# (cold-cache, serialised launches: compare shares, not absolutes)\n')
        fh.write(f'{"kernel":70s} {"launches":>8s} {"total ms":>12s} {"share %":>8s} {"mean ms":>10s}\n')
        for k, v in agg.items():
            fh.write(f'{k:70s} {len(v):8d} {sum(v) / 1e6:12.3f} {100 * sum(v) / tot:8.2f} {sum(v) / len(v) / 1e6:10.3f}\n')


KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__icc_request_hit_rate.pct', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__warps_eligible.avg.per_cycle_active']


def ncu_full(tag):
    rep = os.path.join(OUT, f'prof_c1_{tag}.ncu-rep')
    if not os.path.isfile(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f'{tag}_ncu_c1_integrate.txt'), 'w') as fh:
        fh.write('# ncu --set full --clock-control none --import-source on -k regex:sonic_integrate  '
                 'python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f'kernel: {d["Kernel Name"]}\n')
            for k in KEYS:
                if k in d:
                    fh.write(f'  {k:70s} {d[k]:>18s} {units[hdr.index(k)]}\n')
            st = {k: float(v.replace(',', '')) for k, v in d.items()
                  if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and v not in ('', 'n/a')}
            tot = sum(st.values()) or 1
            fh.write('  warp stall reasons (pc sampling):\n')
            for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
                fh.write(f'    {k.replace("smsp__pcsamp_warps_issue_stalled_", ""):30s} {100 * v / tot:6.1f} %\n')
    # right-hand sides evaluated by one C1 launch (from the plain run of the same command)
    nticks = '1'
    plain = os.path.join(OUT, f'plain_c1_{tag}.log')
    if os.path.isfile(plain):
        for ln in open(plain):
            if ln.startswith('{'):
                nticks = str(json.loads(ln)['roofline']['rhs_evaluations'])
    sass = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'sass_profile.py'), rep,
                           os.path.join(ROOT, 'pysonic_b200', 'libsonic_b200.so'),
                           '_Z22sonic_integrate_kernelILb0EEv8SonicJob', nticks], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, f'{tag}_sass_profile.txt'), 'w') as fh:
        fh.write('# executed warp-instructions per source function of sonic_integrate_kernel (ncu source page x nvdisasm line info)\n')
        fh.write('\n'.join(sass.splitlines()[:45]) + '\n')


def c2_counters(tag):
    path = os.path.join(OUT, f'c2_counters_{tag}.csv')
    if not os.path.isfile(path):
        return None
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    im, iu, iv = hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
    vals = {r[im]: (r[iv], r[iu]) for r in rows[1:]}
    with open(os.path.join(PROF, f'{tag}_ncu_c2_counters.txt'), 'w') as fh:
        fh.write('# ncu --metrics ... -k regex:sonic_integrate -c 1  python bench.py --steps 1 --warmup 0 --no-cpu-baseline  (C2, one launch)\n')
        for k, (v, u) in vals.items():
            fh.write(f'  {k:70s} {v:>18s} {u}\n')
    try:
        return float(vals['dram__bytes_read.sum'][0].replace(',', '')) + float(vals['dram__bytes_write.sum'][0].replace(',', ''))
    except KeyError:
        return None


def workload_counters(tag):
    ''' integrator / averaging counters of the C3 (STN), C4 (SWnode) and C5 (TC) workloads. '''
    with open(os.path.join(PROF, f'{tag}_ncu_workloads.txt'), 'w') as fh:
        fh.write('# ncu --metrics ... -k regex:"sonic_integrate|sonic_average"  python tools/gpu_workload.py <W>\n')
        for w in ('STN', 'SWnode', 'TC'):
            path = os.path.join(OUT, f'counters_{w}_{tag}.csv')
            if not os.path.isfile(path):
                continue
            rows = [r for r in csv.reader(open(path)) if len(r) > 10]
            hdr = rows[0]
            ik, ii = hdr.index('Kernel Name'), hdr.index('ID')
            im, iu, iv = hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
            plain = os.path.join(OUT, f'plain_{w}_{tag}.log')
            if os.path.isfile(plain):
                fh.write(f'## {w}: ' + open(plain).read().strip().splitlines()[-1] + '\n')
            last = None
            for r in rows[1:]:
                if (r[ii], r[ik]) != last:
                    last = (r[ii], r[ik])
                    fh.write(f'kernel {r[ik][:60]} (launch {r[ii]})\n')
                fh.write(f'  {r[im]:70s} {r[iv]:>18s} {r[iu]}\n')


def avg_full(tag):
    rep = os.path.join(OUT, f'prof_avg_stn_{tag}.ncu-rep')
    if not os.path.isfile(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = KEYS + ['l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
                   'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                   'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed']
    with open(os.path.join(PROF, f'{tag}_ncu_average_stn.txt'), 'w') as fh:
        fh.write('# ncu --set full --clock-control none --import-source on -k regex:sonic_average -c 1  python tools/gpu_workload.py STN\n')
        fh.write('# (C3 shape: STN, 7 344 trajectories x 100 coverage fractions x 19 tables)\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f'kernel: {d["Kernel Name"]}\n')
            for k in keys:
                if k in d:
                    fh.write(f'  {k:70s} {d[k]:>18s} {units[hdr.index(k)]}\n')
            st = {k: float(v.replace(',', '')) for k, v in d.items()
                  if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and v not in ('', 'n/a')}
            tot = sum(st.values()) or 1
            fh.write('  warp stall reasons (pc sampling):\n')
            for k, v in sorted(st.items(), key=lambda x: -x[1])[:6]:
                fh.write(f'    {k.replace("smsp__pcsamp_warps_issue_stalled_", ""):30s} {100 * v / tot:6.1f} %\n')


def extras(tag):
    ''' parity report, workload probes, kernel probe: small JSON files copied as they are. '''
    import shutil
    for src, dst in ((f'parity_{tag}.json', f'{tag}_parity.json'), (f'configs_{tag}.json', f'{tag}_configs_c3_c4_c5.json'),
                     (f'offnode_{tag}.json', f'{tag}_offnode.json'), (f'overtones_{tag}.json', f'{tag}_overtones_lookup.json'),
                     (f'kprobe_{tag}.json', f'{tag}_kprobe.json'), (f'ticks_{tag}.jsonl', f'{tag}_tick_calibration.jsonl'),
                     (f'workloads_{tag}.json', f'{tag}_workloads.json')):
        p = os.path.join(OUT, src)
        if os.path.isfile(p) and os.path.getsize(p) > 0:
            shutil.copyfile(p, os.path.join(PROF, dst))
    workload_counters(tag)
    avg_full(tag)


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    lines = {}
    for name, key in ((f'bench_{tag}.json', 'own'), (f'bench_ref_{tag}.json', 'reference')):
        p = os.path.join(OUT, name)
        if os.path.isfile(p):
            txt = [ln for ln in open(p).read().splitlines() if ln.startswith('{')]
            if txt:
                lines[key] = json.loads(txt[-1])
    if lines:
        with open(os.path.join(PROF, f'{tag}_bench.json'), 'w') as fh:
            json.dump(lines, fh, indent=1)
    launches(tag)
    ncu_full(tag)
    traffic = c2_counters(tag)
    if traffic:
        with open(os.path.join(PROF, 'traffic.json'), 'w') as fh:
            json.dump({'c2': {'integrate_dram_bytes_per_launch': traffic, 'source': f'profiles/{tag}_ncu_c2_counters.txt'}}, fh, indent=1)
    for f in ('pytest_gpu', 'smi'):
        for ext in ('log', 'txt'):
            p = os.path.join(OUT, f'{f}_{tag}.{ext}')
            if os.path.isfile(p):
                with open(p) as src, open(os.path.join(PROF, f'{tag}_{f}.txt'), 'w') as dst:
                    dst.write(''.join(src.readlines()[-15:]))
    extras(tag)
    print('profiles/:', sorted(x for x in os.listdir(PROF) if x.startswith(tag)))


if __name__ == '__main__':
    main()
