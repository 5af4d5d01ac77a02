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

// Arithmetic of the rate expressions.  The reference's formulas are compiled as they are written (neurons.py), on a
// wrapper type: a division is a multiplication by a refined hardware reciprocal (<= 2 ulp; 1 / 0 = inf, 1 / inf = 0),
// the exponential is the branch-free one of the integrator
// (sonic_exp, 1-2 ulp) inside its range and the library's outside.  The averaging kernel evaluates every rate of the
// neuron at 1000 samples x every coverage fraction of every point: on the STN coverage sweep (19 tables) this is the
// second largest kernel of the run.  (sonic_core.h is included before this header.)
struct rd {{
    double v;
    __device__ __forceinline__ rd(double x) : v(x) {{}}
}};
static __device__ __forceinline__ rd operator+(rd a, rd b) {{ return rd(a.v + b.v); }}
static __device__ __forceinline__ rd operator-(rd a, rd b) {{ return rd(a.v - b.v); }}
static __device__ __forceinline__ rd operator*(rd a, rd b) {{ return rd(a.v * b.v); }}
// (sonic_rcp is for the integrator's tame arguments: its refinement turns 1 / 0 and 1 / inf into NaN, and the rates do
// reach exp() = inf and x / 0 at their singular points: this one keeps the hardware seed there, without a branch)
static __device__ __forceinline__ double rate_rcp(double x) {{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    const double r2 = fma(r, e, r);
    return r2 == r2 ? r2 : r;          // 1 / 0 = inf and 1 / inf = 0 as the hardware seed gives them
}}
static __device__ __forceinline__ rd operator/(rd a, rd b) {{ return rd(a.v * rate_rcp(b.v)); }}
static __device__ __forceinline__ rd operator-(rd a) {{ return rd(-a.v); }}
static __device__ __forceinline__ rd operator+(rd a) {{ return a; }}
static __device__ __forceinline__ bool operator<(rd a, rd b) {{ return a.v < b.v; }}
static __device__ __forceinline__ bool operator>(rd a, rd b) {{ return a.v > b.v; }}
static __device__ __forceinline__ bool operator<=(rd a, rd b) {{ return a.v <= b.v; }}
static __device__ __forceinline__ bool operator>=(rd a, rd b) {{ return a.v >= b.v; }}
static __device__ __forceinline__ rd exp(rd x) {{ return rd(fabs(x.v) < 690.0 ? sonic_exp(x.v) : ::exp(x.v)); }}
// x / (exp(x / y) - 1): naive form of the reference (pneuron.py:351-354), 0/0 at x = 0 kept.
static __device__ __forceinline__ rd vtrap(rd x, rd y) {{ return x / (exp(x / y) - 1); }}

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
             '    static __device__ __forceinline__ void eval(const double Vm_, double* r) {',
             '        const rd Vm(Vm_);']
    for k, v in spec['consts'].items():
        lines.append(f'        const rd {k}({_fmt(v)});')
    for item in spec['kin']:
        for st in item.pre:
            lines.append(f'        const rd {st};')
    i = 0
    for item in spec['kin']:
        if isinstance(item, Gate):
            lines.append(f'        const rd inf_{item.key} = {item.xinf};')
            lines.append(f'        const rd tau_{item.key} = {item.tau};')
            lines.append(f'        r[{i}] = rd(inf_{item.key} / tau_{item.key}).v;')
            lines.append(f'        r[{i + 1}] = rd((1 - inf_{item.key}) / tau_{item.key}).v;')
            i += 2
        else:
            assert isinstance(item, Rate)
            lines.append(f'        r[{i}] = rd({item.expr}).v;')
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
