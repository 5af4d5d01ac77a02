// sonic_quad.h -- average intermolecular pressure of a bilayer sonophore (bls.py:359-408) with the
// quadrature the reference gets from scipy.integrate.quad.
//
// scipy.integrate.quad(f, 0, a) with its default tolerances (epsabs = epsrel = 1.49e-8) is QUADPACK's QAGS:
// a 21-point Gauss-Kronrod rule, adaptive bisection of the sub-interval with the largest error estimate,
// Wynn's epsilon algorithm on the sequence of partial sums, at most 50 sub-intervals.  The integrand here
// is a FORCE of order 1e-12 ... 1e-7 N, so the absolute tolerance 1.49e-8 is met after zero to a few
// bisections and the value the reference works with is the (unconverged, up to 15 % off for deflections
// around a / 2) result of exactly that sequence of rules.  The Lennard-Jones fit of computePMparams
// (bls.py:410-470) is made on those values, so the sequence is part of the result and is restated here
// from QUADPACK's published algorithm (Piessens, de Doncker-Kapenga, Ueberhuber, Kahaner 1983: QK21, QAGSE,
// QELG, QPSRT).  scipy's QUADPACK is third-party and not under /root/reference; the restatement is pinned
// bit for bit against scipy.integrate.quad in tests/test_hostsim.py (CPU build of this header).
//
// One deflection value = one sequential QAGS run (one thread on the device).  Products that feed sums are
// rounded separately (SONIC_QMUL), as the Fortran / NumPy originals do: no fused multiply-add.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define SONIC_QHD __host__ __device__ __forceinline__
#define SONIC_QHDM __host__ __device__ __forceinline__
#else
#define SONIC_QHD static inline
#define SONIC_QHDM inline
#endif
#if defined(__CUDA_ARCH__)
#define SONIC_QMUL(a, b) __dmul_rn((a), (b))
#define SONIC_QADD(a, b) __dadd_rn((a), (b))
#else
#define SONIC_QMUL(a, b) ((a) * (b))
#define SONIC_QADD(a, b) ((a) + (b))
#endif

#define SONIC_Q_LIMIT 50
#define SONIC_Q_EPMACH 2.220446049250313e-16
#define SONIC_Q_UFLOW 2.2250738585072014e-308
#define SONIC_Q_OFLOW 1.7976931348623157e+308

// Integrand of PMavg: 2 pi r PMlocal(r, Z, R), with the operation order of bls.py:359-404.
struct SonicPmIntegrand {
    double a, Delta, Z, R, absR, absZ, sgn;
    SONIC_QHDM void init(double a_, double Delta_gap, double Z_) {
        a = a_; Delta = Delta_gap; Z = Z_;
        R = (SONIC_QMUL(a_, a_) + SONIC_QMUL(Z_, Z_)) / (2.0 * Z_);        // curvrad, bls.py:296 (inf at Z = 0)
        absR = fabs(R); absZ = fabs(Z_);
        sgn = Z_ > 0.0 ? 1.0 : (Z_ < 0.0 ? -1.0 : 0.0);
    }
    SONIC_QHDM double operator()(double r) const {
        const double pDelta = 1.0e5, Delta_ = 1.4e-9, m = 5.0, n = 3.3;        // bls.py:92-97
        double z = 0.0;                                                        // localDeflection, bls.py:368-371
        if (absZ != 0.0) z = sgn * SONIC_QADD(sqrt(SONIC_QMUL(R, R) - SONIC_QMUL(r, r)) - absR, absZ);
        const double relgap = SONIC_QADD(2.0 * z, Delta) / Delta_;            // bls.py:385-386
        const double u = 1.0 / relgap;
        const double pm = SONIC_QMUL(pDelta, pow(u, m) - pow(u, n));
        return SONIC_QMUL(SONIC_QMUL(2.0 * 3.141592653589793, r), pm);       // bls.py:402
    }
};

// 21-point Gauss-Kronrod rule on [a, b] with QUADPACK's error heuristics (QK21).
template <class F>
SONIC_QHD void sonic_qk21(const F& f, double a, double b, double* result, double* abserr, double* resabs,
                          double* resasc) {
    const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                            0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                            0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                            0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                            0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
    const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                            0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                            0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                            0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                            0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                            0.149445554002916905664936468389821};
    const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                          0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                          0.295524224714752870173815619188769};
    const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    const double fc = f(centr);
    double resg = 0.0, resk = SONIC_QMUL(wgk[10], fc);
    double rabs = fabs(resk);
    double fv1[10], fv2[10];
    for (int j = 0; j < 5; j++) {
        const int jtw = 2 * j + 1;
        const double absc = SONIC_QMUL(hlgth, xgk[jtw]);
        const double f1 = f(centr - absc), f2 = f(centr + absc);
        fv1[jtw] = f1; fv2[jtw] = f2;
        const double fsum = f1 + f2;
        resg = SONIC_QADD(resg, SONIC_QMUL(wg[j], fsum));
        resk = SONIC_QADD(resk, SONIC_QMUL(wgk[jtw], fsum));
        rabs = SONIC_QADD(rabs, SONIC_QMUL(wgk[jtw], fabs(f1) + fabs(f2)));
    }
    for (int j = 0; j < 5; j++) {
        const int jt = 2 * j;
        const double absc = SONIC_QMUL(hlgth, xgk[jt]);
        const double f1 = f(centr - absc), f2 = f(centr + absc);
        fv1[jt] = f1; fv2[jt] = f2;
        const double fsum = f1 + f2;
        resk = SONIC_QADD(resk, SONIC_QMUL(wgk[jt], fsum));
        rabs = SONIC_QADD(rabs, SONIC_QMUL(wgk[jt], fabs(f1) + fabs(f2)));
    }
    const double reskh = resk * 0.5;
    double rasc = SONIC_QMUL(wgk[10], fabs(fc - reskh));
    for (int j = 0; j < 10; j++) rasc = SONIC_QADD(rasc, SONIC_QMUL(wgk[j], fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh)));
    *result = SONIC_QMUL(resk, hlgth);
    rabs = SONIC_QMUL(rabs, dhlgth);
    rasc = SONIC_QMUL(rasc, dhlgth);
    double err = fabs(SONIC_QMUL(resk - resg, hlgth));
    if (rasc != 0.0 && err != 0.0) {
        const double t = pow(200.0 * err / rasc, 1.5);
        err = SONIC_QMUL(rasc, t < 1.0 ? t : 1.0);
    }
    if (rabs > SONIC_Q_UFLOW / (50.0 * SONIC_Q_EPMACH)) {
        const double floor_ = SONIC_QMUL(SONIC_Q_EPMACH * 50.0, rabs);
        err = floor_ > err ? floor_ : err;
    }
    *abserr = err; *resabs = rabs; *resasc = rasc;
}

// Wynn's epsilon algorithm on the table of partial sums (QELG).  Arrays are addressed from 1 as in
// the original; epstab needs 52 + 3 entries, res3la 4.
SONIC_QHD void sonic_qelg(int* n_io, double* epstab, double* result, double* abserr, double* res3la, int* nres) {
    int n = *n_io;
    *nres += 1;
    *abserr = SONIC_Q_OFLOW;
    *result = epstab[n];
    if (n >= 3) {
        const int limexp = 50;
        epstab[n + 2] = epstab[n];
        const int newelm = (n - 1) / 2;
        epstab[n] = SONIC_Q_OFLOW;
        const int num = n;
        int k1 = n;
        for (int i = 1; i <= newelm; i++) {
            const int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            const double e1abs = fabs(e1);
            const double delta2 = e2 - e1, err2 = fabs(delta2);
            const double tol2 = SONIC_QMUL(fmax(fabs(e2), e1abs), SONIC_Q_EPMACH);
            const double delta3 = e1 - e0, err3 = fabs(delta3);
            const double tol3 = SONIC_QMUL(fmax(e1abs, fabs(e0)), SONIC_Q_EPMACH);
            if (!(err2 > tol2 || err3 > tol3)) {
                // e0, e1 and e2 equal to within machine accuracy: convergence
                *result = res;
                *abserr = fmax(err2 + err3, SONIC_QMUL(5.0 * SONIC_Q_EPMACH, fabs(res)));
                *n_io = n;
                return;
            }
            const double e3 = epstab[k1];
            epstab[k1] = e1;
            const double delta1 = e1 - e3, err1 = fabs(delta1);
            const double tol1 = SONIC_QMUL(fmax(e1abs, fabs(e3)), SONIC_Q_EPMACH);
            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
                n = i + i - 1;
                break;
            }
            const double ss = SONIC_QADD(1.0 / delta1, 1.0 / delta2) - 1.0 / delta3;
            const double epsinf = fabs(SONIC_QMUL(ss, e1));
            if (!(epsinf > 1e-4)) {
                n = i + i - 1;
                break;
            }
            res = SONIC_QADD(e1, 1.0 / ss);
            epstab[k1] = res;
            k1 -= 2;
            const double error = SONIC_QADD(SONIC_QADD(err2, fabs(res - e2)), err3);
            if (error > *abserr) continue;
            *abserr = error;
            *result = res;
        }
        if (n == limexp) n = 2 * (limexp / 2) - 1;
        int ib = ((num / 2) * 2 == num) ? 2 : 1;
        const int ie = newelm + 1;
        for (int i = 1; i <= ie; i++) {
            const int ib2 = ib + 2;
            epstab[ib] = epstab[ib2];
            ib = ib2;
        }
        if (num != n) {
            int indx = num - n + 1;
            for (int i = 1; i <= n; i++) epstab[i] = epstab[indx++];
        }
        if (*nres >= 4) {
            *abserr = SONIC_QADD(SONIC_QADD(fabs(*result - res3la[3]), fabs(*result - res3la[2])), fabs(*result - res3la[1]));
            res3la[1] = res3la[2];
            res3la[2] = res3la[3];
            res3la[3] = *result;
        } else {
            res3la[*nres] = *result;
            *abserr = SONIC_Q_OFLOW;
        }
    }
    *abserr = fmax(*abserr, SONIC_QMUL(5.0 * SONIC_Q_EPMACH, fabs(*result)));
    *n_io = n;
}

// Keep the list of error estimates in descending order and pick the interval to bisect next (QPSRT).
SONIC_QHD void sonic_qpsrt(int limit, int last, int* maxerr, double* ermax, const double* elist, int* iord, int* nrmax) {
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
    } else {
        const double errmax = elist[*maxerr];
        if (*nrmax != 1) {
            const int ido = *nrmax - 1;
            for (int i = 1; i <= ido; i++) {
                const int isucc = iord[*nrmax - 1];
                if (errmax <= elist[isucc]) break;
                iord[*nrmax] = isucc;
                *nrmax -= 1;
            }
        }
        int jupbn = last;
        if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
        const double errmin = elist[last];
        const int jbnd = jupbn - 1, ibeg = *nrmax + 1;
        bool done = false;
        for (int i = ibeg; i <= jbnd; i++) {
            int isucc = iord[i];
            if (errmax >= elist[isucc]) {
                // insert errmin by traversing the list bottom-up
                iord[i - 1] = *maxerr;
                int k = jbnd;
                bool placed = false;
                for (int j = i; j <= jbnd; j++) {
                    isucc = iord[k];
                    if (errmin < elist[isucc]) {
                        iord[k + 1] = last;
                        placed = true;
                        break;
                    }
                    iord[k + 1] = isucc;
                    k--;
                }
                if (!placed) iord[i] = last;
                done = true;
                break;
            }
            iord[i - 1] = isucc;
        }
        if (!done) {
            iord[jbnd] = *maxerr;
            iord[jupbn] = last;
        }
    }
    *maxerr = iord[*nrmax];
    *ermax = elist[*maxerr];
}

// QAGSE with scipy.integrate.quad's defaults.  Returns the integral; *last_out = number of sub-intervals.
template <class F>
SONIC_QHD double sonic_qags(const F& f, double a, double b, int* last_out) {
    const double epsabs = 1.49e-8, epsrel = 1.49e-8;
    const int limit = SONIC_Q_LIMIT;
    double alist[SONIC_Q_LIMIT + 2], blist[SONIC_Q_LIMIT + 2], rlist[SONIC_Q_LIMIT + 2], elist[SONIC_Q_LIMIT + 2];
    int iord[SONIC_Q_LIMIT + 2];
    double rlist2[56], res3la[4] = {0.0, 0.0, 0.0, 0.0};
    int ier = 0, ierro = 0;
    double result, abserr, defabs, resabs;
    alist[1] = a; blist[1] = b;
    sonic_qk21(f, a, b, &result, &abserr, &defabs, &resabs);
    const double dres = fabs(result);
    double errbnd = fmax(epsabs, SONIC_QMUL(epsrel, dres));
    int last = 1;
    rlist[1] = result; elist[1] = abserr; iord[1] = 1;
    if (abserr <= SONIC_QMUL(100.0 * SONIC_Q_EPMACH, defabs) && abserr > errbnd) ier = 2;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
        *last_out = last;
        return result;
    }
    rlist2[1] = result;
    double errmax = abserr, area = result, errsum = abserr;
    int maxerr = 1, nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
    bool extrap = false, noext = false;
    abserr = SONIC_Q_OFLOW;
    double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0;
    int exit_to = 0;       // 100 = final tests, 115 = sum of the list
    for (last = 2; last <= limit; last++) {
        const double a1 = alist[maxerr], b1 = 0.5 * (alist[maxerr] + blist[maxerr]), a2 = b1, b2 = blist[maxerr];
        const double erlast = errmax;
        double area1, error1, area2, error2, defab1, defab2, dummy;
        sonic_qk21(f, a1, b1, &area1, &error1, &dummy, &defab1);
        sonic_qk21(f, a2, b2, &area2, &error2, &dummy, &defab2);
        const double area12 = area1 + area2, erro12 = error1 + error2;
        errsum = SONIC_QADD(errsum, erro12) - errmax;
        area = SONIC_QADD(area, area12) - rlist[maxerr];
        if (!(defab1 == error1 || defab2 == error2)) {
            if (!(fabs(rlist[maxerr] - area12) > SONIC_QMUL(1e-5, fabs(area12)) || erro12 < SONIC_QMUL(0.99, errmax))) {
                if (extrap) iroff2++;
                else iroff1++;
            }
            if (last > 10 && erro12 > errmax) iroff3++;
        }
        rlist[maxerr] = area1;
        rlist[last] = area2;
        errbnd = fmax(epsabs, SONIC_QMUL(epsrel, fabs(area)));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= SONIC_QMUL(1.0 + 100.0 * SONIC_Q_EPMACH, fabs(a2) + 1000.0 * SONIC_Q_UFLOW)) ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
            rlist[maxerr] = area2; rlist[last] = area1;
            elist[maxerr] = error2; elist[last] = error1;
        } else {
            alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
            elist[maxerr] = error1; elist[last] = error2;
        }
        sonic_qpsrt(limit, last, &maxerr, &errmax, elist, iord, &nrmax);
        if (errsum <= errbnd) { exit_to = 115; break; }
        if (ier != 0) { exit_to = 100; break; }
        if (last == 2) {
            small = SONIC_QMUL(fabs(b - a), 0.375);
            erlarg = errsum;
            ertest = errbnd;
            rlist2[2] = area;
            continue;
        }
        if (noext) continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small) erlarg += erro12;
        if (!extrap) {
            // is the interval to be bisected next the smallest one?
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = true;
            nrmax = 2;
        }
        if (!(ierro == 3 || erlarg <= ertest)) {
            // the smallest interval has the largest error: before bisecting, decrease the sum of the
            // errors over the larger intervals and extrapolate
            const int id = nrmax;
            int jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            bool next = false;
            for (int k = id; k <= jupbnd; k++) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { next = true; break; }
                nrmax++;
            }
            if (next) continue;
        }
        numrl2++;
        rlist2[numrl2] = area;
        double reseps, abseps;
        sonic_qelg(&numrl2, rlist2, &reseps, &abseps, res3la, &nres);
        ktmin++;
        if (ktmin > 5 && abserr < SONIC_QMUL(1e-3, errsum)) ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, SONIC_QMUL(epsrel, fabs(reseps)));
            if (abserr <= ertest) { exit_to = 100; break; }
        }
        if (numrl2 == 1) noext = true;
        if (ier == 5) { exit_to = 100; break; }
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = false;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit;
    *last_out = last;
    if (exit_to != 115) {
        bool sum_list = false;
        if (abserr == SONIC_Q_OFLOW) {
            sum_list = true;
        } else if (ier + ierro != 0) {
            if (ierro == 3) abserr += correc;
            if (result != 0.0 && area != 0.0) {
                if (abserr / fabs(result) > errsum / fabs(area)) sum_list = true;
            } else if (abserr > errsum) {
                sum_list = true;
            } else if (area == 0.0) {
                return result;
            }
        }
        if (!sum_list) return result;       // (the divergence tests that follow only set the error flag)
    }
    result = 0.0;
    for (int k = 1; k <= last; k++) result = SONIC_QADD(result, rlist[k]);
    return result;
}

// PMavg(Z) for a sonophore of radius a and gap Delta (bls.py:390-404): total force / stretched surface.
SONIC_QHD double sonic_pmavg_point(double a, double Delta, double Z, int* last) {
    SonicPmIntegrand f;
    f.init(a, Delta, Z);
    const double ftotal = sonic_qags(f, 0.0, a, last);
    const double S = SONIC_QMUL(3.141592653589793, SONIC_QMUL(a, a) + SONIC_QMUL(Z, Z));     // surface, bls.py:309
    return ftotal / S;
}
