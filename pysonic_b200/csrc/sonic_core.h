// sonic_core.h -- per-grid-point state machine of the SONIC lookup engine.
//
// One "lane" (= one CUDA thread) owns one grid point (a, f, A, Q) and carries the whole
// reference algorithm for that point:
//
//   * bilayer-sonophore right-hand side          (reference: PySONIC/core/bls.py:681-718)
//   * quasi-static initial deflection Z0         (bls.py:538-573,720-725; scipy brentq)
//   * a variable-order, variable-step Adams/BDF integrator with automatic stiffness
//     switching, finite-difference Jacobian and Nordsieck history -- the LSODA algorithm
//     (Hindmarsh/Petzold) that the reference reaches through scipy.integrate.odeint
//     (solvers.py:167), restated here from its published description, with odeint's call
//     pattern (fresh problem per acoustic cycle, 999 interpolated outputs per cycle,
//     rtol = atol = 1.49012e-8, <= 500 steps per output interval)
//   * the periodic-convergence bookkeeping       (solvers.py:283-365)
//
// The integrator is written as a *tick machine*: every tick performs exactly ONE right-hand
// side evaluation at a lane-specific (t, y) and then the lane-specific bookkeeping selected
// by `phase`.  On the GPU all 32 lanes of a warp therefore execute the expensive part (two
// pow, one sin, three divisions) convergently, whatever step/order/method each lane is in;
// only the cheap bookkeeping diverges.
//
// The file compiles for the device (nvcc) and, for CPU-side unit tests of the logic only
// (tests/hostsim), for the host.  The product never runs the host build.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SONIC_HD __host__ __device__ __forceinline__
#define SONIC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SONIC_HD static inline
#define SONIC_HD_NOINLINE static
#endif

#define SONIC_NEQ 3
#define SONIC_MXORDN 12
#define SONIC_MXORDS 5
#define SONIC_NYH 13           /* max order + 1 Nordsieck columns */
#define SONIC_NPC 1000         /* samples per cycle (constants.py:35) */
#define SONIC_NOUT 999         /* new samples appended per cycle (solvers.py:168-170) */
#define SONIC_NCYC_CAP 11      /* solvers.py:353-359 with NCYCLES_MAX = 10 */
#define SONIC_MXSTEP 500       /* LSODA default, odeint passes mxstep = 0 */

// status bits reported per point
#define SONIC_ST_OK 0u
#define SONIC_ST_NOCONV 1u      /* periodic criterion not met at the cycle cap (solvers.py:362) */
#define SONIC_ST_ZCLAMP 2u      /* Z < Zmin clamp hit in the RHS (bls.py:695-697) */
#define SONIC_ST_MXSTEP 4u      /* > 500 steps inside one output interval (odeint "excess work") */
#define SONIC_ST_STEPFAIL 8u    /* repeated error-test / corrector failures */
#define SONIC_ST_Z0FAIL 16u     /* no sign change for the quasi-static root (bls.py:566-572) */
#define SONIC_ST_TOLSF 32u      /* too much accuracy requested for machine precision */

// ---------------------------------------------------------------------------------------
// physical constants (bls.py:87-110, constants.py:13)
// ---------------------------------------------------------------------------------------
#define SONIC_PI 3.141592653589793
#define SONIC_RG 8.31342
#define SONIC_T 309.15
#define SONIC_DELTA0 2.0e-9
#define SONIC_RHOL 1075.0
#define SONIC_MUL 7.0e-4
#define SONIC_MUS 0.035
#define SONIC_KA 0.24
#define SONIC_ALPHA_TISSUE 7.56
#define SONIC_C0 0.62
#define SONIC_KH 1.613e5
#define SONIC_P0 1.0e5
#define SONIC_DGL 3.68e-9
#define SONIC_XI 0.5e-9
#define SONIC_EPS0 8.854e-12
#define SONIC_REL_ZMIN (-0.49)

#define SONIC_UROUND 2.220446049250313e-16
#define SONIC_RTOL 1.49012e-8
#define SONIC_ATOL 1.49012e-8
#define SONIC_CONV_THR 1e-4     /* MAX_RMSE_PTP_RATIO, constants.py:31 */

// ---------------------------------------------------------------------------------------
// Method coefficient tables (both families), filled on the host by sonic_fill_tables().
// index [meth-1][nq-1][...]
// ---------------------------------------------------------------------------------------
struct SonicTables {
    double elco[2][12][13];
    double tesco[2][12][3];
    double cm1[12];
    double cm2[5];
    double sm1[12];
};

// Per-radius constants (one entry per sonophore radius of the lookup).
struct SonicBls {
    double a;        // radius (m)
    double Delta;    // equilibrium gap (m)
    double x0, C, nrep, nattr;   // Lennard-Jones fit of the intermolecular pressure
    double Cm0;      // resting capacitance (F/m2)
    double depth;    // embedding depth d (m): kA_tissue = 2 * alpha * f * d
};

// Per-point constants derived once per lane.
struct SonicPoint {
    double a, a2, inva2, Delta, x0, Clj, nrep, nattr, Zmin;
    double V0, c_vol;          // V = V0 * (1 + Z * c_vol * (3 + Z^2 * inva2)), c_vol = 1/(3 Delta)
    double kAtot;              // kA + kA_tissue (N/m)
    double pel0;               // Q^2 / (2 eps0 epsR)
    double omega;              // 2 pi f
    double A;
    double ng0;
};

SONIC_HD void sonic_point_init(SonicPoint& p, const SonicBls& b, double f, double A, double Q) {
    p.a = b.a;
    p.a2 = b.a * b.a;
    p.inva2 = 1.0 / p.a2;
    p.Delta = b.Delta;
    p.x0 = b.x0;
    p.Clj = b.C;
    p.nrep = b.nrep;
    p.nattr = b.nattr;
    p.Zmin = SONIC_REL_ZMIN * b.Delta;
    p.V0 = SONIC_PI * b.Delta * p.a2;                   // bls.py:136
    p.c_vol = 1.0 / (3.0 * b.Delta);
    p.kAtot = SONIC_KA + 2.0 * (SONIC_ALPHA_TISSUE * f) * b.depth;   // bls.py:583-602
    p.pel0 = Q * Q / (2.0 * (SONIC_EPS0 * 1.0));        // bls.py:489-491
    p.omega = 2.0 * SONIC_PI * f;
    p.A = A;
    p.ng0 = SONIC_P0 * p.V0 / (SONIC_RG * SONIC_T);     // bls.py:137,529-536
}

// Lennard-Jones intermolecular pressure (bls.py:29-41,472-480).
SONIC_HD double sonic_pm(const SonicPoint& p, double Z) {
    const double x = p.x0 / (2.0 * Z + p.Delta);
#ifdef SONIC_FAST_POW
    const double lx = log(x);
    return p.Clj * (exp(p.nrep * lx) - exp(p.nattr * lx));
#else
    return p.Clj * (pow(x, p.nrep) - pow(x, p.nattr));
#endif
}

// Gas pressure in the cavity (bls.py:311-319,518-526).
SONIC_HD double sonic_pg(const SonicPoint& p, double Z, double ng) {
    const double V = p.V0 * (1.0 + Z * p.c_vol * (3.0 + Z * Z * p.inva2));
    return ng * (SONIC_RG * SONIC_T) / V;
}

// Right-hand side (bls.py:681-718).  Returns true if the Zmin clamp was applied.
SONIC_HD bool sonic_rhs(const SonicPoint& p, double t, const double y[3], double dy[3]) {
    const double U = y[0];
    double Z = y[1];
    const double ng = y[2];
    bool clamped = false;
    if (Z < p.Zmin) {
        Z = p.Zmin;
        clamped = true;
    }
    const double s2 = p.a2 + Z * Z;                     // S / pi
    const double inv_s2 = 1.0 / s2;
    const double invR = 2.0 * Z * inv_s2;               // 1 / curvature radius (0 at Z = 0)
    const double ainvR = fabs(invR);
    const double Pg = sonic_pg(p, Z, ng);
    const double Pm = sonic_pm(p, Z);
    const double Pac = p.A * sin(p.omega * t - SONIC_PI);             // drives.py:303-304
    const double Pv = -12.0 * U * SONIC_DELTA0 * SONIC_MUS * (invR * invR)
                      - 4.0 * U * SONIC_MUL * ainvR;                  // bls.py:613-631
    const double zr = Z / p.a;
    const double PE = -(p.kAtot * (zr * zr)) * invR;                  // bls.py:575-611
    const double Pel = -(p.a2 * inv_s2) * p.pel0;                     // bls.py:482-491
    const double Ptot = Pm + Pg - SONIC_P0 - Pac + PE + Pv + Pel;
    dy[0] = Ptot * ainvR / SONIC_RHOL - 1.5 * (U * U) * invR;         // bls.py:633-655
    dy[1] = U;
    dy[2] = 2.0 * (SONIC_PI * s2) * SONIC_DGL * (SONIC_C0 - Pg / SONIC_KH) / SONIC_XI;  // :508-516
    return clamped;
}

// Quasi-static net pressure (bls.py:538-553).
SONIC_HD double sonic_ptot_qs(const SonicPoint& p, double Z, double ng, double Pac) {
    const double s2 = p.a2 + Z * Z;
    return sonic_pm(p, Z) + sonic_pg(p, Z, ng) - SONIC_P0 - Pac - (p.a2 / s2) * p.pel0;
}

// Initial deflection: root of the quasi-static pressure on (Zmin, a) for
// Pac = A sin(2 pi f dt - pi), dt = 1/(1000 f) (bls.py:555-573,720-725).  Bracketing
// inverse-quadratic/secant/bisection root finder with the tolerances the reference passes
// to scipy.optimize.brentq (xtol = 1e-16, rtol = 4 eps, 100 iterations).
SONIC_HD bool sonic_z0(const SonicPoint& p, double f, double* z0) {
    const double dt = 1.0 / (SONIC_NPC * f);
    const double Pac = p.A * sin(p.omega * dt - SONIC_PI);
    const double xtol = 1e-16, rtol = 8.881784197001252e-16;
    double xpre = p.Zmin, xcur = p.a;
    double xblk = 0., fblk = 0., spre = 0., scur = 0.;
    double fpre = sonic_ptot_qs(p, xpre, p.ng0, Pac);
    double fcur = sonic_ptot_qs(p, xcur, p.ng0, Pac);
    if (!(fpre > 0. && 0. > fcur)) {
        *z0 = 0.;
        return false;
    }
    for (int i = 0; i < 100; i++) {
        if (fpre != 0. && fcur != 0. && ((fpre < 0.) != (fcur < 0.))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2.;
        const double sbis = (xblk - xcur) / 2.;
        if (fcur == 0. || fabs(sbis) < delta) break;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2. * fabs(stry) < fmin(fabs(spre), 3. * fabs(sbis) - delta)) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta)
            xcur += scur;
        else
            xcur += (sbis > 0. ? delta : -delta);
        fcur = sonic_ptot_qs(p, xcur, p.ng0, Pac);
    }
    *z0 = xcur;
    return true;
}

// Membrane capacitance (bls.py:330-345).
SONIC_HD double sonic_capacitance(double a2, double Delta, double Cm0, double Z) {
    if (Z == 0.0) return Cm0;
    const double Z2 = (a2 - Z * Z - Z * Delta) / (2.0 * Z);
    return Cm0 * Delta / a2 * (Z + Z2 * log((2.0 * Z + Delta) / Delta));
}

// ---------------------------------------------------------------------------------------
// Integrator lane state
// ---------------------------------------------------------------------------------------
enum SonicPhase : int {
    PH_INIT = 0,        // RHS at (t0, y0): first call of a fresh problem (one per cycle)
    PH_CORR_FIRST = 1,  // RHS at the predicted state (m = 0)
    PH_CORR_ITER = 2,   // RHS at a corrector iterate (m > 0)
    PH_JAC = 3,         // RHS at a perturbed state (finite-difference Jacobian column)
    PH_RESET = 4,       // RHS at y_n after 3+ error-test failures (order reset)
    PH_DONE = 5
};

struct SonicLane {
    // ---- integrator (LSODA) ----
    double yh[SONIC_NYH][SONIC_NEQ];   // Nordsieck history, yh[j] = h^j y^(j) / j!
    double wm[9];                      // LU factors of P = I - h el0 J (row-major)
    double ewt[3], savf[3], acor[3], y[3];
    double h, hu, tn, told, rc, el0, crate, rmax, conit;
    double pdest, pdlast, pdnorm, dsm, pnorm, del, delp, rate;
    double yj_save, jac_r0;
    int ipvt[3];
    int nq, l, meth, miter, mused, nqu, ialth, ipup, icount, irflag, jcur, kflag, m, ncf;
    int nst, nslp, nslast, jstart, lmax, tab_meth, jcol, ierpj;
    int phase;
    // ---- output / cycle bookkeeping ----
    double t0, tstop, tstep, tout;
    double ssq_z, ssq_ng, min_z, max_z, min_ng, max_ng;
    int cyc, kout;
    unsigned status;
    // ---- statistics ----
    unsigned nfe, nje, nsteps;
};

SONIC_HD double sonic_mnorm(const double v[3], const double w[3]) {
    double vm = 0.0;
    vm = fmax(vm, fabs(v[0]) * w[0]);
    vm = fmax(vm, fabs(v[1]) * w[1]);
    vm = fmax(vm, fabs(v[2]) * w[2]);
    return vm;
}

SONIC_HD void sonic_ewset(SonicLane& s) {
    for (int i = 0; i < 3; i++) s.ewt[i] = 1.0 / (SONIC_RTOL * fabs(s.yh[0][i]) + SONIC_ATOL);
}

// coefficient accessors: l-vector of the current order in the currently loaded table
#define SONIC_EL(s, T, j) ((T)->elco[(s).tab_meth - 1][(s).nq - 1][(j)])
#define SONIC_TESCO(s, T, k) ((T)->tesco[(s).tab_meth - 1][(s).nq - 1][(k)])

// Reset the order-dependent constants (order nq of the loaded family).
SONIC_HD void sonic_set_order(SonicLane& s, const SonicTables* T) {
    const double el1 = SONIC_EL(s, T, 0);
    s.rc = s.rc * el1 / s.el0;
    s.el0 = el1;
    s.conit = 0.5 / (s.nq + 2);
}

// Multiply yh by the Pascal triangle (prediction) or its inverse (retraction).
SONIC_HD void sonic_pascal(SonicLane& s, int sign) {
    // For jb = 1..nq the sweep updates columns nq-jb .. nq-1 (0-based) in ascending order.
    for (int jb = 1; jb <= s.nq; jb++) {
        for (int j = s.nq - jb; j < s.nq; j++) {
            if (sign > 0) {
                s.yh[j][0] += s.yh[j + 1][0];
                s.yh[j][1] += s.yh[j + 1][1];
                s.yh[j][2] += s.yh[j + 1][2];
            } else {
                s.yh[j][0] -= s.yh[j + 1][0];
                s.yh[j][1] -= s.yh[j + 1][1];
                s.yh[j][2] -= s.yh[j + 1][2];
            }
        }
    }
}

// Apply a step-size ratio: bound it, restrict by the Adams stability region, rescale history.
SONIC_HD void sonic_rescale(SonicLane& s, const SonicTables* T, double rh) {
    rh = fmin(rh, s.rmax);
    if (s.meth == 1) {
        s.irflag = 0;
        const double pdh = fmax(fabs(s.h) * s.pdlast, 0.000001);
        if (rh * pdh * 1.00001 >= T->sm1[s.nq - 1]) {
            rh = T->sm1[s.nq - 1] / pdh;
            s.irflag = 1;
        }
    }
    double r = 1.0;
    for (int j = 1; j < s.l; j++) {
        r *= rh;
        s.yh[j][0] *= r;
        s.yh[j][1] *= r;
        s.yh[j][2] *= r;
    }
    s.h *= rh;
    s.rc *= rh;
    s.ialth = s.l;
}

// Prediction: advance tn, apply Pascal triangle, set the evaluation point.
SONIC_HD void sonic_predict(SonicLane& s) {
    if (fabs(s.rc - 1.0) > 0.3) s.ipup = s.miter;
    if (s.nst >= s.nslp + 20) s.ipup = s.miter;
    s.tn += s.h;
    sonic_pascal(s, +1);
    s.pnorm = sonic_mnorm(s.yh[0], s.ewt);
    s.m = 0;
    s.rate = 0.0;
    s.del = 0.0;
    s.y[0] = s.yh[0][0];
    s.y[1] = s.yh[0][1];
    s.y[2] = s.yh[0][2];
    s.phase = PH_CORR_FIRST;
}

// 3x3 LU with partial pivoting (column-major semantics of the reference solver: first
// maximal pivot, multipliers stored negated) and the matching solve.
SONIC_HD int sonic_lu3(double a[9], int ipvt[3]) {
    // a[i + 3*j] = A(i, j)
    int info = 0;
    for (int k = 0; k < 2; k++) {
        int lmax_ = k;
        double dmax = fabs(a[k + 3 * k]);
        for (int i = k + 1; i < 3; i++)
            if (fabs(a[i + 3 * k]) > dmax) {
                dmax = fabs(a[i + 3 * k]);
                lmax_ = i;
            }
        ipvt[k] = lmax_;
        if (a[lmax_ + 3 * k] == 0.0) {
            info = k + 1;
            continue;
        }
        if (lmax_ != k) {
            const double t = a[lmax_ + 3 * k];
            a[lmax_ + 3 * k] = a[k + 3 * k];
            a[k + 3 * k] = t;
        }
        const double tinv = -1.0 / a[k + 3 * k];
        for (int i = k + 1; i < 3; i++) a[i + 3 * k] *= tinv;
        for (int j = k + 1; j < 3; j++) {
            double t = a[lmax_ + 3 * j];
            if (lmax_ != k) {
                a[lmax_ + 3 * j] = a[k + 3 * j];
                a[k + 3 * j] = t;
            }
            for (int i = k + 1; i < 3; i++) a[i + 3 * j] += t * a[i + 3 * k];
        }
    }
    ipvt[2] = 2;
    if (a[8] == 0.0) info = 3;
    return info;
}

SONIC_HD void sonic_lusolve3(const double a[9], const int ipvt[3], double b[3]) {
    for (int k = 0; k < 2; k++) {
        const int lp = ipvt[k];
        const double t = b[lp];
        if (lp != k) {
            b[lp] = b[k];
            b[k] = t;
        }
        for (int i = k + 1; i < 3; i++) b[i] += t * a[i + 3 * k];
    }
    for (int k = 2; k >= 0; k--) {
        b[k] /= a[k + 3 * k];
        const double t = -b[k];
        for (int i = 0; i < k; i++) b[i] += t * a[i + 3 * k];
    }
}

// Interpolate the solution at time t from the Nordsieck history (k = 0 derivative).
SONIC_HD void sonic_interp(const SonicLane& s, double t, double out[3]) {
    const double sfrac = (t - s.tn) / s.h;
    out[0] = s.yh[s.l - 1][0];
    out[1] = s.yh[s.l - 1][1];
    out[2] = s.yh[s.l - 1][2];
    for (int j = s.nq - 1; j >= 0; j--) {
        out[0] = s.yh[j][0] + sfrac * out[0];
        out[1] = s.yh[j][1] + sfrac * out[1];
        out[2] = s.yh[j][2] + sfrac * out[2];
    }
}

// Start a fresh problem at (t0, y0) for one acoustic cycle of period T (odeint call).
SONIC_HD void sonic_cycle_begin(SonicLane& s, double t0, double T, const double y0[3]) {
    s.t0 = t0;
    s.tstop = t0 + T;                                   // solvers.py:334
    s.tstep = (s.tstop - s.t0) / (double)SONIC_NOUT;    // numpy.linspace step
    s.kout = 1;
    // tout_k = k * step + start, two separately rounded operations as numpy does
#if defined(__CUDA_ARCH__)
    s.tout = __dadd_rn(__dmul_rn(1.0, s.tstep), s.t0);
#else
    s.tout = 1.0 * s.tstep + s.t0;
#endif
    s.ssq_z = s.ssq_ng = 0.0;
    s.min_z = s.min_ng = INFINITY;
    s.max_z = s.max_ng = -INFINITY;
    s.y[0] = y0[0];
    s.y[1] = y0[1];
    s.y[2] = y0[2];
    s.tn = t0;
    s.phase = PH_INIT;
}

SONIC_HD double sonic_tout_at(const SonicLane& s, int k) {
    if (k >= SONIC_NOUT) return s.tstop;
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn((double)k, s.tstep), s.t0);
#else
    volatile double prod = (double)k * s.tstep;
    return prod + s.t0;
#endif
}

// Evaluation point of the next tick.
SONIC_HD double sonic_eval_time(const SonicLane& s) { return s.tn; }

struct SonicSink {
    // where the lane stores its per-cycle samples; stride in doubles between samples
    double* zbuf;    // [1000]: zbuf[0] = Z at the start of the current cycle, zbuf[k] sample k
    double* ngbuf;   // [1000]
    long stride;
};

// --- pieces of the step controller ----------------------------------------------------------

// Fatal integrator condition: stop the lane.
SONIC_HD void sonic_fail(SonicLane& s, unsigned bit) {
    s.status |= bit;
    s.phase = PH_DONE;
}

// Choose the next order/step after a success (ialth == 0) or an error-test failure.
// Returns: 0 = accepted and finished (step OK, no redo), 1 = must redo the step (predict again)
SONIC_HD int sonic_select(SonicLane& s, const SonicTables* T, double rhup, int iredo) {
    const double exsm = 1.0 / s.l;
    double rhsm = 1.0 / (1.2 * pow(s.dsm, exsm) + 0.0000012);
    double rhdn = 0.0;
    if (s.nq != 1) {
        const double ddn = sonic_mnorm(s.yh[s.l - 1], s.ewt) / SONIC_TESCO(s, T, 0);
        const double exdn = 1.0 / s.nq;
        rhdn = 1.0 / (1.3 * pow(ddn, exdn) + 0.0000013);
    }
    double pdh = 0.0;
    if (s.meth == 1) {
        pdh = fmax(fabs(s.h) * s.pdlast, 0.000001);
        if (s.l < s.lmax) rhup = fmin(rhup, T->sm1[s.l - 1] / pdh);
        rhsm = fmin(rhsm, T->sm1[s.nq - 1] / pdh);
        if (s.nq > 1) rhdn = fmin(rhdn, T->sm1[s.nq - 2] / pdh);
        s.pdest = 0.0;
    }
    int newq;
    double rh;
    if (rhsm >= rhup) {
        if (rhsm < rhdn) {
            newq = s.nq - 1;
            rh = rhdn;
            if (s.kflag < 0 && rh > 1.0) rh = 1.0;
        } else {
            newq = s.nq;
            rh = rhsm;
        }
    } else {
        if (rhup > rhdn) {
            // order increase
            newq = s.l;
            rh = rhup;
            if (rh < 1.1) {
                s.ialth = 3;
                return 0;
            }
            const double r = SONIC_EL(s, T, s.l - 1) / s.l;
            s.yh[newq][0] = s.acor[0] * r;
            s.yh[newq][1] = s.acor[1] * r;
            s.yh[newq][2] = s.acor[2] * r;
            s.nq = newq;
            s.l = s.nq + 1;
            sonic_set_order(s, T);
            sonic_rescale(s, T, rh);
            if (iredo == 0) s.rmax = 10.0;
            return iredo != 0;
        }
        newq = s.nq - 1;
        rh = rhdn;
        if (s.kflag < 0 && rh > 1.0) rh = 1.0;
    }
    // 10 percent test, bypassed when Adams step is stability-limited
    bool bypass = false;
    if (s.meth == 1 && rh * pdh * 1.00001 >= T->sm1[newq - 1]) bypass = true;
    if (!bypass && s.kflag == 0 && rh < 1.1) {
        s.ialth = 3;
        return 0;
    }
    if (s.kflag <= -2) rh = fmin(rh, 0.2);
    if (newq != s.nq) {
        s.nq = newq;
        s.l = s.nq + 1;
        sonic_set_order(s, T);
    }
    sonic_rescale(s, T, rh);
    if (iredo == 0) s.rmax = 10.0;
    return iredo != 0;
}

// Preliminaries before attempting a step (driver level + step entry).
// Returns false if the lane stopped.
SONIC_HD bool sonic_step_begin(SonicLane& s, const SonicTables* T, bool first) {
    if (!first) {
        if (s.nst - s.nslast >= SONIC_MXSTEP) {
            sonic_fail(s, SONIC_ST_MXSTEP);
            return false;
        }
        sonic_ewset(s);
    }
    const double tolsf = SONIC_UROUND * sonic_mnorm(s.yh[0], s.ewt);
    if (tolsf > 1.0) {
        sonic_fail(s, SONIC_ST_TOLSF);
        return false;
    }
    // step entry
    s.kflag = 0;
    s.told = s.tn;
    s.ncf = 0;
    s.ierpj = 0;
    s.jcur = 0;
    s.delp = 0.0;
    if (s.jstart == 0) {
        s.lmax = SONIC_MXORDN + 1;
        s.nq = 1;
        s.l = 2;
        s.ialth = 2;
        s.rmax = 10000.0;
        s.rc = 0.0;
        s.el0 = 1.0;
        s.crate = 0.7;
        s.nslp = 0;
        s.ipup = s.miter;
        s.icount = 20;
        s.irflag = 0;
        s.pdest = 0.0;
        s.pdlast = 0.0;
        s.tab_meth = 1;
        sonic_set_order(s, T);
    } else if (s.jstart == -1) {
        s.ipup = s.miter;
        s.lmax = (s.meth == 2 ? SONIC_MXORDS : SONIC_MXORDN) + 1;
        if (s.ialth == 1) s.ialth = 2;
        if (s.meth != s.mused) {
            s.tab_meth = s.meth;
            s.ialth = s.l;
            sonic_set_order(s, T);
        }
    }
    s.jstart = 1;
    sonic_predict(s);
    return true;
}

// Output handling after a successful step: emit every sample reached, handle the end of the
// cycle (convergence test, next cycle or stop).  Returns true if the lane continues stepping
// within the same problem, false if it either finished or started a new problem (phase set).
SONIC_HD bool sonic_emit(SonicLane& s, const SonicSink& sink, double period) {
    while ((s.tn - s.tout) * s.h >= 0.0) {
        double yo[3];
        sonic_interp(s, s.tout, yo);
        const long idx = (long)s.kout * sink.stride;
        if (s.cyc >= 1) {
            const double dz = yo[1] - sink.zbuf[idx];
            const double dn = yo[2] - sink.ngbuf[idx];
            s.ssq_z += dz * dz;
            s.ssq_ng += dn * dn;
            s.min_z = fmin(s.min_z, yo[1]);
            s.max_z = fmax(s.max_z, yo[1]);
            s.min_ng = fmin(s.min_ng, yo[2]);
            s.max_ng = fmax(s.max_ng, yo[2]);
        }
        sink.zbuf[idx] = yo[1];
        sink.ngbuf[idx] = yo[2];
        if (s.kout == SONIC_NOUT) {
            // end of cycle (solvers.py:317-365)
            bool stop = false;
            if (s.cyc >= 1) {
                const double rz = sqrt(s.ssq_z / (double)SONIC_NOUT) / (s.max_z - s.min_z);
                const double rn = sqrt(s.ssq_ng / (double)SONIC_NOUT) / (s.max_ng - s.min_ng);
                const bool stable = (rz < SONIC_CONV_THR) && (rn < SONIC_CONV_THR);
                if (stable) stop = true;
                else if (s.cyc >= SONIC_NCYC_CAP - 1) {
                    stop = true;
                    s.status |= SONIC_ST_NOCONV;
                }
            }
            s.cyc++;
            if (stop) {
                s.phase = PH_DONE;
                return false;
            }
            sink.zbuf[0] = yo[1];
            sink.ngbuf[0] = yo[2];
            sonic_cycle_begin(s, s.tstop, period, yo);
            return false;
        }
        s.kout++;
        s.tout = sonic_tout_at(s, s.kout);
        s.nslast = s.nst;
    }
    return true;
}

// After a successful step (history updated, h/order chosen): driver-level bookkeeping.
SONIC_HD void sonic_after_step(SonicLane& s, const SonicTables* T, const SonicSink& sink,
                               double period) {
    if (s.meth != s.mused) {
        // method switch: force coefficient reload on the next step entry
        s.jstart = -1;
    }
    if (!sonic_emit(s, sink, period)) return;
    sonic_step_begin(s, T, false);
}

// Consider switching Adams <-> BDF after a successful step. Returns true if a switch was made
// (history rescaled, step finished).
SONIC_HD bool sonic_method_switch(SonicLane& s, const SonicTables* T) {
    if (s.meth == 1) {
        if (s.nq > 5) return false;
        double rh2;
        int nqm2;
        if (s.dsm > 100.0 * s.pnorm * SONIC_UROUND && s.pdest != 0.0) {
            const double exsm = 1.0 / s.l;
            double rh1 = 1.0 / (1.2 * pow(s.dsm, exsm) + 0.0000012);
            double rh1it = 2.0 * rh1;
            const double pdh = s.pdlast * fabs(s.h);
            if (pdh * rh1 > 0.00001) rh1it = T->sm1[s.nq - 1] / pdh;
            rh1 = fmin(rh1, rh1it);
            if (s.nq > SONIC_MXORDS) {
                nqm2 = SONIC_MXORDS;
                const int lm2 = SONIC_MXORDS + 1;
                const double exm2 = 1.0 / lm2;
                const double dm2 = sonic_mnorm(s.yh[lm2], s.ewt) / T->cm2[SONIC_MXORDS - 1];
                rh2 = 1.0 / (1.2 * pow(dm2, exm2) + 0.0000012);
            } else {
                const double dm2 = s.dsm * (T->cm1[s.nq - 1] / T->cm2[s.nq - 1]);
                rh2 = 1.0 / (1.2 * pow(dm2, exsm) + 0.0000012);
                nqm2 = s.nq;
            }
            if (rh2 < 5.0 * rh1) return false;
        } else {
            if (s.irflag == 0) return false;
            rh2 = 2.0;
            nqm2 = s.nq < SONIC_MXORDS ? s.nq : SONIC_MXORDS;
        }
        s.icount = 20;
        s.meth = 2;
        s.miter = 2;
        s.pdlast = 0.0;
        s.nq = nqm2;
        s.l = s.nq + 1;
        sonic_rescale(s, T, rh2);
        s.rmax = 10.0;
        return true;
    }
    // currently BDF: consider Adams
    const double exsm = 1.0 / s.l;
    double rh1, dm1, exm1;
    int nqm1;
    if (SONIC_MXORDN < s.nq) {
        nqm1 = SONIC_MXORDN;
        const int lm1 = SONIC_MXORDN + 1;
        exm1 = 1.0 / lm1;
        dm1 = sonic_mnorm(s.yh[lm1], s.ewt) / T->cm1[SONIC_MXORDN - 1];
        rh1 = 1.0 / (1.2 * pow(dm1, exm1) + 0.0000012);
    } else {
        dm1 = s.dsm * (T->cm2[s.nq - 1] / T->cm1[s.nq - 1]);
        rh1 = 1.0 / (1.2 * pow(dm1, exsm) + 0.0000012);
        nqm1 = s.nq;
        exm1 = exsm;
    }
    double rh1it = 2.0 * rh1;
    const double pdh = s.pdnorm * fabs(s.h);
    if (pdh * rh1 > 0.00001) rh1it = T->sm1[nqm1 - 1] / pdh;
    rh1 = fmin(rh1, rh1it);
    const double rh2 = 1.0 / (1.2 * pow(s.dsm, exsm) + 0.0000012);
    if (rh1 * 5.0 < 5.0 * rh2) return false;
    const double alpha = fmax(0.001, rh1);
    dm1 = pow(alpha, exm1) * dm1;
    if (dm1 <= 1000.0 * SONIC_UROUND * s.pnorm) return false;
    s.icount = 20;
    s.meth = 1;
    s.miter = 0;
    s.pdlast = 0.0;
    s.nq = nqm1;
    s.l = s.nq + 1;
    sonic_rescale(s, T, rh1);
    s.rmax = 10.0;
    return true;
}

// The corrector converged: local error test, history update, order/step/method selection.
SONIC_HD void sonic_converged(SonicLane& s, const SonicTables* T, const SonicSink& sink,
                              double period) {
    s.jcur = 0;
    const double tq2 = SONIC_TESCO(s, T, 1);
    s.dsm = (s.m == 0) ? s.del / tq2 : sonic_mnorm(s.acor, s.ewt) / tq2;
    if (s.dsm > 1.0) {
        // ---- error test failed ----
        s.kflag--;
        s.tn = s.told;
        sonic_pascal(s, -1);
        s.rmax = 2.0;
        if (fabs(s.h) <= 0.0) {
            sonic_fail(s, SONIC_ST_STEPFAIL);
            return;
        }
        if (s.kflag <= -3) {
            if (s.kflag == -10) {
                sonic_fail(s, SONIC_ST_STEPFAIL);
                return;
            }
            s.h *= 0.1;
            s.y[0] = s.yh[0][0];
            s.y[1] = s.yh[0][1];
            s.y[2] = s.yh[0][2];
            s.phase = PH_RESET;
            return;
        }
        sonic_select(s, T, 0.0, 2);
        sonic_predict(s);
        return;
    }
    // ---- step accepted ----
    s.kflag = 0;
    s.nst++;
    s.nsteps++;
    s.hu = s.h;
    s.nqu = s.nq;
    s.mused = s.meth;
    for (int j = 0; j < s.l; j++) {
        const double e = SONIC_EL(s, T, j);
        s.yh[j][0] += e * s.acor[0];
        s.yh[j][1] += e * s.acor[1];
        s.yh[j][2] += e * s.acor[2];
    }
    s.icount--;
    bool switched = false;
    if (s.icount < 0) switched = sonic_method_switch(s, T);
    if (!switched) {
        s.ialth--;
        if (s.ialth == 0) {
            double rhup = 0.0;
            if (s.l != s.lmax) {
                double d[3];
                d[0] = s.acor[0] - s.yh[s.lmax - 1][0];
                d[1] = s.acor[1] - s.yh[s.lmax - 1][1];
                d[2] = s.acor[2] - s.yh[s.lmax - 1][2];
                const double dup = sonic_mnorm(d, s.ewt) / SONIC_TESCO(s, T, 2);
                const double exup = 1.0 / (s.l + 1);
                rhup = 1.0 / (1.4 * pow(dup, exup) + 0.0000014);
            }
            sonic_select(s, T, rhup, 0);
        } else if (s.ialth <= 1 && s.l != s.lmax) {
            s.yh[s.lmax - 1][0] = s.acor[0];
            s.yh[s.lmax - 1][1] = s.acor[1];
            s.yh[s.lmax - 1][2] = s.acor[2];
        }
    }
    sonic_after_step(s, T, sink, period);
}

// The corrector iteration failed to converge.
SONIC_HD void sonic_corrector_failed(SonicLane& s, const SonicTables* T) {
    if (s.miter != 0 && s.jcur != 1) {
        // retry with a fresh Jacobian at the predicted state
        s.ipup = s.miter;
        s.m = 0;
        s.rate = 0.0;
        s.del = 0.0;
        s.y[0] = s.yh[0][0];
        s.y[1] = s.yh[0][1];
        s.y[2] = s.yh[0][2];
        s.phase = PH_CORR_FIRST;
        return;
    }
    s.ncf++;
    s.rmax = 2.0;
    s.tn = s.told;
    sonic_pascal(s, -1);
    if (fabs(s.h) <= 0.0 || s.ncf == 10) {
        sonic_fail(s, SONIC_ST_STEPFAIL);
        return;
    }
    s.ipup = s.miter;
    sonic_rescale(s, T, 0.25);
    sonic_predict(s);
}

// Corrector update with the RHS value in savf (functional iteration or chord/Newton).
SONIC_HD void sonic_corrector(SonicLane& s, const SonicTables* T, const SonicSink& sink,
                              double period) {
    const double el1 = SONIC_EL(s, T, 0);
    if (s.miter == 0) {
        double d[3];
        for (int i = 0; i < 3; i++) {
            s.savf[i] = s.h * s.savf[i] - s.yh[1][i];
            d[i] = s.savf[i] - s.acor[i];
        }
        s.del = sonic_mnorm(d, s.ewt);
        for (int i = 0; i < 3; i++) {
            s.y[i] = s.yh[0][i] + el1 * s.savf[i];
            s.acor[i] = s.savf[i];
        }
    } else {
        double d[3];
        for (int i = 0; i < 3; i++) d[i] = s.h * s.savf[i] - (s.yh[1][i] + s.acor[i]);
        sonic_lusolve3(s.wm, s.ipvt, d);
        s.del = sonic_mnorm(d, s.ewt);
        for (int i = 0; i < 3; i++) {
            s.acor[i] += d[i];
            s.y[i] = s.yh[0][i] + el1 * s.acor[i];
        }
    }
    // convergence test
    bool conv = false;
    if (s.del <= 100.0 * s.pnorm * SONIC_UROUND) {
        conv = true;
    } else if (!(s.m == 0 && s.meth == 1)) {
        if (s.m != 0) {
            double rm = 1024.0;
            if (s.del <= 1024.0 * s.delp) rm = s.del / s.delp;
            s.rate = fmax(s.rate, rm);
            s.crate = fmax(0.2 * s.crate, rm);
        }
        const double dcon = s.del * fmin(1.0, 1.5 * s.crate) / (SONIC_TESCO(s, T, 1) * s.conit);
        if (dcon <= 1.0) {
            s.pdest = fmax(s.pdest, s.rate / fabs(s.h * el1));
            if (s.pdest != 0.0) s.pdlast = s.pdest;
            conv = true;
        }
    }
    if (conv) {
        sonic_converged(s, T, sink, period);
        return;
    }
    s.m++;
    if (s.m == 3 || (s.m >= 2 && s.del > 2.0 * s.delp)) {
        sonic_corrector_failed(s, T);
        return;
    }
    s.delp = s.del;
    s.phase = PH_CORR_ITER;   // next RHS at (tn, y)
}

// Begin the finite-difference Jacobian: perturb component jcol of y.
SONIC_HD void sonic_jac_perturb(SonicLane& s, double r0) {
    const int j = s.jcol;
    const double yj = s.y[j];
    const double r = fmax(1.4901161193847656e-08 * fabs(yj), r0 / s.ewt[j]);
    s.yj_save = yj;
    s.y[j] = yj + r;
}

// One tick: consume the RHS value `f` evaluated at (sonic_eval_time, s.y) and advance the
// lane to its next evaluation point.
SONIC_HD void sonic_tick(SonicLane& s, const SonicTables* T, const SonicSink& sink,
                         double period, const double f[3]) {
    s.nfe++;
    switch (s.phase) {
    case PH_INIT: {
        // fresh problem (one per cycle): initial step size, history, first step
        s.yh[0][0] = s.y[0]; s.yh[0][1] = s.y[1]; s.yh[0][2] = s.y[2];
        s.nst = 0; s.nslast = 0; s.hu = 0.0; s.nqu = 0; s.mused = 0; s.miter = 0;
        s.meth = 1; s.jstart = 0; s.nq = 1;
        sonic_ewset(s);
        const double tdist = fabs(s.tout - s.tn);
        const double w0 = fmax(fabs(s.tn), fabs(s.tout));
        double tol = SONIC_RTOL;
        tol = fmax(tol, 100.0 * SONIC_UROUND);
        tol = fmin(tol, 0.001);
        double sum = sonic_mnorm(f, s.ewt);
        sum = 1.0 / (tol * w0 * w0) + tol * sum * sum;
        double h0 = 1.0 / sqrt(sum);
        h0 = fmin(h0, tdist);
        s.h = h0;   // tout > t always
        s.yh[1][0] = h0 * f[0]; s.yh[1][1] = h0 * f[1]; s.yh[1][2] = h0 * f[2];
        sonic_step_begin(s, T, true);
        break;
    }
    case PH_CORR_FIRST: {
        s.savf[0] = f[0]; s.savf[1] = f[1]; s.savf[2] = f[2];
        if (s.ipup > 0) {
            // P = I - h el0 J must be re-evaluated: finite-difference Jacobian, 3 more ticks
            s.nje++;
            s.ierpj = 0;
            s.jcur = 1;
            const double fac = sonic_mnorm(s.savf, s.ewt);
            double r0 = 1000.0 * fabs(s.h) * SONIC_UROUND * 3.0 * fac;
            if (r0 == 0.0) r0 = 1.0;
            s.jac_r0 = r0;
            s.jcol = 0;
            sonic_jac_perturb(s, r0);
            s.phase = PH_JAC;
            break;
        }
        s.acor[0] = s.acor[1] = s.acor[2] = 0.0;
        sonic_corrector(s, T, sink, period);
        break;
    }
    case PH_CORR_ITER: {
        s.savf[0] = f[0]; s.savf[1] = f[1]; s.savf[2] = f[2];
        sonic_corrector(s, T, sink, period);
        break;
    }
    case PH_JAC: {
        const int j = s.jcol;
        const double r0 = s.jac_r0;
        const double rr = fmax(1.4901161193847656e-08 * fabs(s.yj_save), r0 / s.ewt[j]);
        const double hl0 = s.h * s.el0;
        const double fac = -hl0 / rr;
        s.wm[0 + 3 * j] = (f[0] - s.savf[0]) * fac;
        s.wm[1 + 3 * j] = (f[1] - s.savf[1]) * fac;
        s.wm[2 + 3 * j] = (f[2] - s.savf[2]) * fac;
        s.y[j] = s.yj_save;
        if (j < 2) {
            s.jcol = j + 1;
            sonic_jac_perturb(s, r0);
            break;
        }
        // norm of the Jacobian (weighted max-norm consistent matrix norm)
        double an = 0.0;
        for (int i = 0; i < 3; i++) {
            double sm = 0.0;
            for (int jj = 0; jj < 3; jj++) sm += fabs(s.wm[i + 3 * jj]) / s.ewt[jj];
            an = fmax(an, sm * s.ewt[i]);
        }
        s.pdnorm = an / fabs(hl0);
        s.wm[0] += 1.0; s.wm[4] += 1.0; s.wm[8] += 1.0;
        const int info = sonic_lu3(s.wm, s.ipvt);
        s.ipup = 0;
        s.rc = 1.0;
        s.nslp = s.nst;
        s.crate = 0.7;
        if (info != 0) {
            s.ierpj = 1;
            // singular iteration matrix: treated as a corrector failure with current Jacobian
            sonic_corrector_failed(s, T);
            break;
        }
        s.acor[0] = s.acor[1] = s.acor[2] = 0.0;
        sonic_corrector(s, T, sink, period);
        break;
    }
    case PH_RESET: {
        s.yh[1][0] = s.h * f[0]; s.yh[1][1] = s.h * f[1]; s.yh[1][2] = s.h * f[2];
        s.ipup = s.miter;
        s.ialth = 5;
        if (s.nq != 1) {
            s.nq = 1;
            s.l = 2;
            sonic_set_order(s, T);
        }
        sonic_predict(s);
        break;
    }
    default:
        break;
    }
}

// Start a lane on its grid point with a precomputed initial deflection.
// y0 = (0, Z0, ng0) (bls.py:737-747).
SONIC_HD void sonic_lane_start(SonicLane& s, const SonicPoint& p, double f, double z0,
                               const SonicSink& sink) {
    s.status = SONIC_ST_OK;
    s.nfe = s.nje = s.nsteps = 0;
    s.cyc = 0;
    const double y0[3] = {0.0, z0, p.ng0};
    sink.zbuf[0] = z0;
    sink.ngbuf[0] = p.ng0;
    sonic_cycle_begin(s, 0.0, 1.0 / f, y0);
}

// Same, computing the initial deflection in place.
SONIC_HD void sonic_lane_init(SonicLane& s, const SonicPoint& p, double f, const SonicSink& sink) {
    double z0;
    if (!sonic_z0(p, f, &z0)) {
        s.status = SONIC_ST_Z0FAIL;
        s.nfe = s.nje = s.nsteps = 0;
        s.cyc = 0;
        s.phase = PH_DONE;
        return;
    }
    sonic_lane_start(s, p, f, z0, sink);
}

// ---------------------------------------------------------------------------------------
// Host-side construction of the method coefficient tables.
// ---------------------------------------------------------------------------------------
static void sonic_fill_tables(SonicTables* T) {
    double pc[13];
    // --- implicit Adams, orders 1..12 ---
    double (*el)[13] = T->elco[0];
    double (*te)[3] = T->tesco[0];
    for (int q = 0; q < 12; q++) {
        for (int i = 0; i < 13; i++) el[q][i] = 0.0;
        for (int i = 0; i < 3; i++) te[q][i] = 0.0;
    }
    el[0][0] = 1.0;
    el[0][1] = 1.0;
    te[0][0] = 0.0;
    te[0][1] = 2.0;
    te[1][0] = 1.0;
    te[11][2] = 0.0;
    pc[0] = 1.0;
    double rqfac = 1.0;
    for (int nq = 2; nq <= 12; nq++) {
        // pc holds the coefficients of p(x) = (x+1)(x+2)...(x+nq-1)
        const double rq1fac = rqfac;
        rqfac = rqfac / nq;
        const int nqm1 = nq - 1;
        const double fnqm1 = nqm1;
        const int nqp1 = nq + 1;
        pc[nq - 1] = 0.0;
        for (int ib = 1; ib <= nqm1; ib++) {
            const int i = nqp1 - ib;
            pc[i - 1] = pc[i - 2] + fnqm1 * pc[i - 1];
        }
        pc[0] = fnqm1 * pc[0];
        // integrals of p(x) and x p(x) over (-1, 0)
        double pint = pc[0];
        double xpin = pc[0] / 2.0;
        double tsign = 1.0;
        for (int i = 2; i <= nq; i++) {
            tsign = -tsign;
            pint += tsign * pc[i - 1] / i;
            xpin += tsign * pc[i - 1] / (i + 1);
        }
        el[nq - 1][0] = pint * rq1fac;
        el[nq - 1][1] = 1.0;
        for (int i = 2; i <= nq; i++) el[nq - 1][i] = rq1fac * pc[i - 1] / i;
        const double agamq = rqfac * xpin;
        const double ragq = 1.0 / agamq;
        te[nq - 1][1] = ragq;
        if (nq < 12) te[nqp1 - 1][0] = ragq * rqfac / nqp1;
        te[nqm1 - 1][2] = ragq;
    }
    // --- BDF, orders 1..5 ---
    el = T->elco[1];
    te = T->tesco[1];
    for (int q = 0; q < 12; q++) {
        for (int i = 0; i < 13; i++) el[q][i] = 0.0;
        for (int i = 0; i < 3; i++) te[q][i] = 0.0;
    }
    pc[0] = 1.0;
    double rq1fac = 1.0;
    for (int nq = 1; nq <= 5; nq++) {
        // pc holds the coefficients of p(x) = (x+1)(x+2)...(x+nq)
        const double fnq = nq;
        const int nqp1 = nq + 1;
        pc[nqp1 - 1] = 0.0;
        for (int ib = 1; ib <= nq; ib++) {
            const int i = nq + 2 - ib;
            pc[i - 1] = pc[i - 2] + fnq * pc[i - 1];
        }
        pc[0] = fnq * pc[0];
        for (int i = 1; i <= nqp1; i++) el[nq - 1][i - 1] = pc[i - 1] / pc[1];
        el[nq - 1][1] = 1.0;
        te[nq - 1][0] = rq1fac;
        te[nq - 1][1] = nqp1 / el[nq - 1][0];
        te[nq - 1][2] = (nq + 2) / el[nq - 1][0];
        rq1fac = rq1fac / fnq;
    }
    for (int i = 0; i < 5; i++) T->cm2[i] = T->tesco[1][i][1] * T->elco[1][i][i + 1];
    for (int i = 0; i < 12; i++) T->cm1[i] = T->tesco[0][i][1] * T->elco[0][i][i + 1];
    static const double sm1[12] = {0.5, 0.575, 0.55, 0.45, 0.35, 0.25, 0.2, 0.15, 0.1, 0.075, 0.05,
                                   0.025};
    for (int i = 0; i < 12; i++) T->sm1[i] = sm1[i];
}

