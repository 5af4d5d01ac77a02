# -*- coding: utf-8 -*-
''' Join an ncu source-page dump (per-SASS-instruction execution counts) with nvdisasm line
    info of the same binary, and aggregate executed warp-instructions per source line and per
    enclosing function of sonic_core.h / sonic_b200.cu.

    usage: python tools/sass_profile.py <report.ncu-rep> <libsonic_b200.so> [kernel-mangled-name] [n_ticks]
'''
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    out, cur, active = [], ('?', 0), False
    for line in txt.splitlines():
        if line.startswith('//-') and '.text.' in line:
            active = ('.text.' + kernel) in line
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def function_map(path):
    ''' line -> enclosing function name (top-level SONIC_HD / __global__ definitions). '''
    fmap, cur = {}, '<top>'
    with open(path) as fh:
        for i, line in enumerate(fh, 1):
            m = re.match(r'^(?:SONIC_HD|static|template|__global__|SONIC_HDM)?.*?\b(sonic_\w+)\s*\(', line)
            if m and not line.startswith(' ') and not line.strip().startswith('//'):
                cur = m.group(1)
            m2 = re.match(r'^\s+// ---- (stage [A-E]\'?)', line)
            if m2 and cur.startswith('sonic_tick'):
                fmap[i] = f'sonic_tick/{m2.group(1)}'
                cur_stage = m2.group(1)
                cur = f'sonic_tick/{cur_stage}'
                continue
            fmap[i] = cur
    return fmap


def main():
    rep, so = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else '_Z22sonic_integrate_kernel8SonicJob'
    nticks = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    data = rows[2:]
    iE = hdr.index('Instructions Executed')
    iS = hdr.index('# Samples')
    iT = hdr.index('Thread Instructions Executed')
    sass = sass_lines(so, kernel)
    if len(sass) != len(data):
        print(f'WARNING: instruction count mismatch: nvdisasm {len(sass)} vs ncu {len(data)} '
              '(binary differs from the profiled one?)')
    n = min(len(sass), len(data))
    here = os.path.dirname(os.path.abspath(__file__))
    fmaps = {'sonic_core.h': function_map(os.path.join(here, '..', 'pysonic_b200', 'csrc', 'sonic_core.h')),
             'sonic_b200.cu': function_map(os.path.join(here, '..', 'pysonic_b200', 'csrc', 'sonic_b200.cu'))}
    per_line, per_fn, samples_fn, thr_fn = {}, {}, {}, {}
    total = 0
    for k in range(n):
        e = int(data[k][iE])
        sm = int(data[k][iS])
        (fname, line) = sass[k][1]
        total += e
        per_line[(fname, line)] = per_line.get((fname, line), 0) + e
        fn = fmaps.get(fname, {}).get(line, fname)
        per_fn[fn] = per_fn.get(fn, 0) + e
        samples_fn[fn] = samples_fn.get(fn, 0) + sm
        thr_fn[fn] = thr_fn.get(fn, 0) + int(data[k][iT])
    tot_s = sum(samples_fn.values()) or 1
    print(f'total executed warp-instructions: {total}  ({total / nticks:.1f} per tick)')
    print('--- per function (instructions per tick, share, stall-sample share)')
    for fn, e in sorted(per_fn.items(), key=lambda x: -x[1])[:40]:
        print(f'{fn:34s} {e / nticks:9.1f} {100 * e / total:6.1f}%  samples {100 * samples_fn[fn] / tot_s:5.1f}%  threads/inst {thr_fn[fn] / max(e, 1):5.1f}')
    print('--- top source lines')
    for (fname, line), e in sorted(per_line.items(), key=lambda x: -x[1])[:40]:
        print(f'{fname}:{line:<5d} {e / nticks:9.1f} {100 * e / total:6.1f}%')


if __name__ == '__main__':
    main()
