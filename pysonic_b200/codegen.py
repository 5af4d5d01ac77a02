# -*- coding: utf-8 -*-
''' Generate the per-neuron CUDA rate functions inlined by the fused cycle-averaging kernel.

    Input: the neuron specs in `neurons.py`.  Output: `csrc/generated/neuron_rates.cuh` with,
    for each neuron id, a specialisation

        template <> struct SonicRates<ID> {
            static constexpr int N = <number of rates>;
            static __device__ __forceinline__ void eval(double Vm, double* r);
        };

    plus the name tables used by the C ABI (`sonic_neuron_id`, `sonic_neuron_rate_name`).
    Run by `__graft_entry__.build()` before nvcc; the generated header is also committed so
    the kernels can be compiled without Python.
'''

import os

from .neurons import NEURON_SPECS, NEURON_ORDER, NEURON_SIM, Gate, Rate, spec_rate_names, MAX_RATES

HEADER = '''// GENERATED FILE -- do not edit.  Produced by pysonic_b200/codegen.py from pysonic_b200/neurons.py.
// Voltage-dependent rate constants (s^-1) of every supported point neuron, as device functions.
#pragma once

#define SONIC_N_NEURONS {nn}
#define SONIC_MAX_RATES {maxr}

// x / (exp(x / y) - 1): naive form of the reference (pneuron.py:351-354), 0/0 at x = 0 kept.
static __device__ __forceinline__ double vtrap(double x, double y) {{ return x / (exp(x / y) - 1); }}

template <int ID> struct SonicRates;

// Net membrane current (mA/m2) of the neurons whose SONIC simulation is supported (NS = number of gating
// states, state k <-> rates 2k and 2k + 1 of SonicRates<ID>); NS = 0: not supported.
template <int ID> struct SonicSim {{
    static constexpr int NS = 0;
    static __device__ __forceinline__ double inet(double, const double*) {{ return 0.0; }}
}};
'''


def _fmt(v):
    return repr(float(v))


def gen_neuron(nid, name):
    spec = NEURON_SPECS[name]
    lines = [f'// ---- {name} ----', f'template <> struct SonicRates<{nid}> {{',
             f'    static constexpr int N = {len(spec_rate_names(name))};',
             '    static __device__ __forceinline__ void eval(const double Vm, double* r) {']
    for k, v in spec['consts'].items():
        lines.append(f'        const double {k} = {_fmt(v)};')
    for item in spec['kin']:
        for st in item.pre:
            lines.append(f'        const double {st};')
    i = 0
    for item in spec['kin']:
        if isinstance(item, Gate):
            lines.append(f'        const double inf_{item.key} = {item.xinf};')
            lines.append(f'        const double tau_{item.key} = {item.tau};')
            lines.append(f'        r[{i}] = inf_{item.key} / tau_{item.key};')
            lines.append(f'        r[{i + 1}] = (1 - inf_{item.key}) / tau_{item.key};')
            i += 2
        else:
            assert isinstance(item, Rate)
            lines.append(f'        r[{i}] = {item.expr};')
            i += 1
    for k in spec['consts']:
        lines.append(f'        (void){k};')
    lines.append('        (void)Vm; (void)r;')
    lines += ['    }', '};', '']
    if name in NEURON_SIM:
        sim = NEURON_SIM[name]
        rates = spec_rate_names(name)
        for k, st in enumerate(sim['states']):
            assert rates[2 * k] == f'alpha{st}' and rates[2 * k + 1] == f'beta{st}', (name, st)
        assert len(rates) == 2 * len(sim['states']), name
        lines += [f'template <> struct SonicSim<{nid}> {{', f"    static constexpr int NS = {len(sim['states'])};",
                  '    static __device__ __forceinline__ double inet(const double Vm, const double* x) {']
        for k, v in sim['consts'].items():
            lines.append(f'        const double {k} = {_fmt(v)};')
        for k, st in enumerate(sim['states']):
            lines.append(f'        const double {st} = x[{k}];')
        lines += [f"        return {sim['inet']};", '    }', '};', '']
    return '\n'.join(lines)


def generate():
    maxr = max(len(spec_rate_names(n)) for n in NEURON_ORDER)
    assert maxr <= MAX_RATES
    out = [HEADER.format(nn=len(NEURON_ORDER), maxr=MAX_RATES)]
    for nid, name in enumerate(NEURON_ORDER):
        out.append(gen_neuron(nid, name))
    out.append('static const char* const SONIC_NEURON_NAMES[SONIC_N_NEURONS] = {' +
               ', '.join(f'"{n}"' for n in NEURON_ORDER) + '};')
    out.append('static const int SONIC_NEURON_NRATES[SONIC_N_NEURONS] = {' +
               ', '.join(str(len(spec_rate_names(n))) for n in NEURON_ORDER) + '};')
    out.append('static const double SONIC_NEURON_CM0[SONIC_N_NEURONS] = {' +
               ', '.join(_fmt(NEURON_SPECS[n]['Cm0']) for n in NEURON_ORDER) + '};')
    rows = []
    for n in NEURON_ORDER:
        names = spec_rate_names(n)
        names = names + [''] * (MAX_RATES - len(names))
        rows.append('    {' + ', '.join(f'"{x}"' for x in names) + '}')
    out.append('static const char* const SONIC_NEURON_RATE_NAMES[SONIC_N_NEURONS][SONIC_MAX_RATES] = {\n' +
               ',\n'.join(rows) + '\n};')
    # dispatch macro: run `CALL(ID)` for the runtime neuron id
    cases = ' '.join(f'case {i}: CALL({i}); break;' for i in range(len(NEURON_ORDER)))
    out.append(f'#define SONIC_DISPATCH_NEURON(id, CALL) switch (id) {{ {cases} default: break; }}')
    return '\n'.join(out) + '\n'


def write(path=None):
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'generated',
                            'neuron_rates.cuh')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    text = generate()
    old = None
    if os.path.isfile(path):
        with open(path) as fh:
            old = fh.read()
    if old != text:
        with open(path, 'w') as fh:
            fh.write(text)
    return path


if __name__ == '__main__':
    print(write())
