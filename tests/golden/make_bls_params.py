# -*- coding: utf-8 -*-
''' Extract the per-(radius, resting charge) bilayer-sonophore constants that the lookup
    path takes as *input* -- equilibrium gap Delta and the 4 Lennard-Jones fit parameters of
    the intermolecular pressure -- from the reference's on-disk cache
    (PySONIC/core/bls_lookups.json, read by bls.py:44-77) into the flat table shipped with
    this package (pysonic_b200/data/bls_params.json).

    Run in the build container only (needs /root/reference):
        python tests/golden/make_bls_params.py
'''

import json
import os

SRC = '/root/reference/PySONIC/core/bls_lookups.json'
DST = os.path.join(os.path.dirname(__file__), '..', '..', 'pysonic_b200', 'data',
                   'bls_params.json')


def main():
    with open(SRC) as fh:
        cache = json.load(fh)
    rows = []
    for akey, per_q in cache.items():
        for qkey, rec in per_q.items():
            lj = rec['LJ_approx']
            rows.append([akey, qkey, rec['Delta_eq'], lj['x0'], lj['C'], lj['nrep'], lj['nattr']])
    rows.sort(key=lambda r: (float(r[0]), float(r[1])))
    out = {
        'source': 'PySONIC/core/bls_lookups.json (reference data cache, bls.py:44-77)',
        'key_format': "a_key = f'{a*1e9:.1f}' (nm), Q_key = f'{Qm0*1e5:.2f}' (nC/cm2)",
        'columns': ['a_key', 'Q_key', 'Delta', 'x0', 'C', 'nrep', 'nattr'],
        'rows': rows,
    }
    with open(DST, 'w') as fh:
        json.dump(out, fh, separators=(',', ':'))
        fh.write('\n')
    print(f'{len(rows)} entries -> {os.path.normpath(DST)}')


if __name__ == '__main__':
    main()
