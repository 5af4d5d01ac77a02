// hostsim.cpp -- CPU build of the per-lane state machine (pysonic_b200/csrc/sonic_core.h).
//
// TEST HARNESS ONLY: lets the CPU test-suite (`-m "not gpu"`) exercise the integrator logic
// that the CUDA kernels run, and compare its step-by-step statistics with scipy's LSODA.
// It is never loaded by the pysonic_b200 package.
#include <stdlib.h>
#include <string.h>

#define SONIC_TRACE 1
#include "../../pysonic_b200/csrc/sonic_core.h"
#include "../../pysonic_b200/csrc/sonic_quad.h"

static SonicTables g_tab;
static int g_tab_ready = 0;

static int g_driver = 0;   // 0 = staged tick (sonic_tick), 1 = nested tick of a lone lane (sonic_tick_lone), 2 = sonic_lone_advance
static double* g_steplog = 0;
static long g_steplog_max = 0, g_steplog_n = 0;

extern "C" {

void hostsim_set_driver(int d) { g_driver = d; }

// optional per-step log: rows of 8 doubles [cycle, nst, tn, hu, h_next, nqu*10+mused, nfe, nq*10+meth]
void hostsim_set_steplog(double* buf, long max_rows) {
    g_steplog = buf;
    g_steplog_max = max_rows;
    g_steplog_n = 0;
}
long hostsim_steplog_rows(void) { return g_steplog_n; }

// bls = {a, Delta, x0, C, nrep, nattr, Cm0, depth}
// trace (optional, may be NULL): rows of 8 doubles per emitted sample
//   [cycle, k, nst, nfe_in_cycle, nje_in_cycle, hu, tn, nqu*10 + mused]
// returns number of ticks executed
// same with charge overtones: ov = [nov][2] (amplitude C/m2, phase rad)
long hostsim_point_ov(const double* bls, double f, double A, double Q, int nov, const double* ov,
                      double* zbuf, double* ngbuf, int* ncycles, unsigned* status, unsigned* stats,
                      double* trace, long trace_rows_max, long* trace_rows) {
    if (!g_tab_ready) {
        sonic_fill_tables(&g_tab);
        g_tab_ready = 1;
    }
    SonicBls b;
    b.a = bls[0]; b.Delta = bls[1]; b.x0 = bls[2]; b.C = bls[3]; b.nrep = bls[4];
    b.nattr = bls[5]; b.Cm0 = bls[6]; b.depth = bls[7];
    SonicPoint p;
    sonic_point_init(p, b, f, A, Q, nov, ov);
    if (p.nov) sonic_update_charge(p, 0.0);
    SonicSink sink;
    sink.zbuf = zbuf; sink.ngbuf = ngbuf;
    SonicLane s;
    memset(&s, 0, sizeof(s));
    double hist[SONIC_H_SIZE];
    memset(hist, 0, sizeof(hist));
    SonicHist H;
    H.base = hist;
    sonic_lane_init(s, H, p, f, sink);
    const double period = 1.0 / f;
    long nticks = 0, nrows = 0;
    unsigned last_nsteps = 0;
    int cyc0 = 0, k0 = 1;
    unsigned nfe_base = 0, nje_base = 0;
    while (s.phase != PH_DONE) {
        if (g_driver == 2) {
            // register-resident runs at fixed order + generic lone ticks (several ticks per call: no per-tick logs)
            const unsigned before = s.nfe - 2 * s.nje;
            if (p.nov) sonic_lone_advance<true>(s, H, &g_tab, p, sink, period);
            else sonic_lone_advance<false>(s, H, &g_tab, p, sink, period);
            nticks += (long)((s.nfe - 2 * s.nje) - before);
            continue;
        }
        double fv[3];
        if (p.nov) sonic_update_charge(p, s.tn);
        if (sonic_rhs(p, s.tn, s.y, fv)) s.status |= SONIC_ST_ZCLAMP;
        if (g_driver == 1) sonic_tick_lone(s, H, &g_tab, p, sink, period, fv);
        else sonic_tick(s, H, &g_tab, p, sink, period, fv, 0u);
        nticks++;
        if (g_steplog && s.nsteps != last_nsteps && g_steplog_n < g_steplog_max) {
            double* r = g_steplog + 8 * g_steplog_n++;
            r[0] = s.cyc; r[1] = s.nst; r[2] = s.told; r[3] = s.hu; r[4] = s.h; r[5] = s.nqu * 10 + s.mused;
            r[6] = s.nfe; r[7] = s.nq * 10 + s.meth;
            last_nsteps = s.nsteps;
        }
        if (trace) {
            // samples emitted by this tick: from (cyc0, k0) up to (s.cyc, s.kout) exclusive
            while (nrows < trace_rows_max && (cyc0 < s.cyc || (cyc0 == s.cyc && k0 < s.kout))) {
                double* r = trace + 8 * nrows++;
                r[0] = cyc0; r[1] = k0; r[2] = s.nst; r[3] = (double)(s.nfe - nfe_base);
                r[4] = (double)(s.nje - nje_base); r[5] = s.hu; r[6] = s.told;
                r[7] = s.nqu * 10 + s.mused;
                k0++;
                if (k0 > SONIC_NOUT) {
                    k0 = 1;
                    cyc0++;
                    nfe_base = s.nfe;
                    nje_base = s.nje;
                }
            }
        }
    }
    *ncycles = s.cyc;
    *status = s.status;
    stats[0] = s.nfe; stats[1] = s.nje; stats[2] = s.nsteps;
    if (trace_rows) *trace_rows = nrows;
    return nticks;
}

// elementwise check of the branch-free elementary functions of the right-hand side
// kind 0: log, 1: exp, 2: sin(2 pi u - pi), 3: reciprocal
void hostsim_math(int kind, const double* x, double* out, long n) {
    for (long i = 0; i < n; i++)
        out[i] = kind == 0 ? sonic_log(x[i]) : kind == 1 ? sonic_exp(x[i]) : kind == 2 ? sonic_sin_drive(x[i]) : sonic_rcp(x[i]);
}

long hostsim_point(const double* bls, double f, double A, double Q, double* zbuf, double* ngbuf,
                   int* ncycles, unsigned* status, unsigned* stats, double* trace,
                   long trace_rows_max, long* trace_rows) {
    return hostsim_point_ov(bls, f, A, Q, 0, 0, zbuf, ngbuf, ncycles, status, stats, trace, trace_rows_max,
                            trace_rows);
}

// charge cycle sample j of a Fourier-series charge (nbls.py:174-177)
double hostsim_charge_sample(double q0, int nov, const double* ov, int j) {
    return sonic_charge_sample(q0, nov, ov, j);
}

void hostsim_rhs(const double* bls, double f, double A, double Q, double t, const double* y,
                 double* dy) {
    SonicBls b;
    b.a = bls[0]; b.Delta = bls[1]; b.x0 = bls[2]; b.C = bls[3]; b.nrep = bls[4];
    b.nattr = bls[5]; b.Cm0 = bls[6]; b.depth = bls[7];
    SonicPoint p;
    sonic_point_init(p, b, f, A, Q);
    sonic_rhs(p, t, y, dy);
}

double hostsim_z0(const double* bls, double f, double A, double Q) {
    SonicBls b;
    b.a = bls[0]; b.Delta = bls[1]; b.x0 = bls[2]; b.C = bls[3]; b.nrep = bls[4];
    b.nattr = bls[5]; b.Cm0 = bls[6]; b.depth = bls[7];
    SonicPoint p;
    sonic_point_init(p, b, f, A, Q);
    double z0 = 0.;
    sonic_z0(p, f, &z0);
    return z0;
}

// average intermolecular pressure with the QAGS restatement (sonic_quad.h), and the bare quadrature of
// a few test integrands (kind 0: sqrt(x), 1: log(x + 1e-9), 2: 1 / sqrt(|x - 0.3| + 1e-6), 3: exp(-50 (x - 0.7)^2),
// 4: sin(30 x) / (x + 0.01), 5: |x - 0.5|^0.3) for comparison with scipy.integrate.quad
void hostsim_pmavg(double a, double Delta, long n, const double* Z, double* out, int* last) {
    for (long i = 0; i < n; i++) out[i] = sonic_pmavg_point(a, Delta, Z[i], last + i);
}

struct HostsimTestIntegrand {
    int kind;
    double operator()(double x) const {
        switch (kind) {
            case 0: return sqrt(x);
            case 1: return log(x + 1e-9);
            case 2: return 1.0 / sqrt(fabs(x - 0.3) + 1e-6);
            case 3: return exp(-50.0 * (x - 0.7) * (x - 0.7));
            case 4: return sin(30.0 * x) / (x + 0.01);
            default: return pow(fabs(x - 0.5), 0.3);
        }
    }
};

double hostsim_quad(int kind, double a, double b, int* last) {
    HostsimTestIntegrand f;
    f.kind = kind;
    return sonic_qags(f, a, b, last);
}

// Decision-equivalence of the two shortcuts of the step-size heuristics (sonic_core.h): for BDF order nq and
// error estimates (dsm, ddn, dup; dup < 0: no candidate for an order increase)
//   out[0] = 1 if sonic_select's shortcut fires ("every candidate is certainly below 1.1")
//   out[1] = the largest of the three step-size candidates, computed in full
//   out[2] = 1 if sonic_method_switch_decide's shortcut fires for (dsm, pdnorm * |h| = pdh)
//   out[3] = 1 if the full method-switch test (shortcut disabled by calling with dsm > 1 guard bypassed) says "switch"
void hostsim_shortcuts(int nq, double dsm, double ddn, double dup, double pdh, double pnorm, double* out) {
    if (!g_tab_ready) {
        sonic_fill_tables(&g_tab);
        g_tab_ready = 1;
    }
    const SonicTables* T = &g_tab;
    const int l = nq + 1;
    out[0] = (dsm > T->thr_sm[nq - 1] && (nq == 1 || ddn > T->thr_dn[nq - 1]) && (dup < 0.0 || dup > T->thr_up[nq - 1])) ? 1.0 : 0.0;
    const double rhsm = 1.0 / (1.2 * pow(dsm, 1.0 / l) + 0.0000012);
    const double rhdn = nq == 1 ? 0.0 : 1.0 / (1.3 * pow(ddn, 1.0 / nq) + 0.0000013);
    const double rhup = (dup < 0.0 || l == SONIC_MXORDS + 1) ? 0.0 : 1.0 / (1.4 * pow(dup, 1.0 / (l + 1)) + 0.0000014);
    out[1] = fmax(rhsm, fmax(rhdn, rhup));
    // method-switch test of a BDF step: the shortcut, and the full comparison it stands for
    out[2] = (dsm <= 1.0 && T->sm1[nq - 1] < 0.8 * pdh) ? 1.0 : 0.0;
    {
        const double exsm = 1.0 / l;
        double rh1 = 1.0 / (1.2 * pow(dsm * T->c21[nq - 1], exsm) + 0.0000012);
        double rh1it = 2.0 * rh1;
        if (pdh * rh1 > 0.00001) rh1it = T->sm1[nq - 1] / pdh;
        rh1 = fmin(rh1, rh1it);
        const double rh2 = 1.0 / (1.2 * pow(dsm, exsm) + 0.0000012);
        bool sw = !(rh1 * 5.0 < 5.0 * rh2);
        if (sw) {
            const double alpha = fmax(0.001, rh1);
            const double dm1 = pow(alpha, exsm) * (dsm * T->c21[nq - 1]);
            if (dm1 <= 1000.0 * SONIC_UROUND * pnorm) sw = false;
        }
        out[3] = sw ? 1.0 : 0.0;
    }
}

long hostsim_tables_size(void) { return (long)(sizeof(SonicTables) / sizeof(double)); }

void hostsim_tables(double* out) {
    if (!g_tab_ready) {
        sonic_fill_tables(&g_tab);
        g_tab_ready = 1;
    }
    memcpy(out, &g_tab, sizeof(g_tab));
}

}  // extern "C"
