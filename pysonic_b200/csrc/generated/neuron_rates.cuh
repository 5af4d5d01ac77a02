// GENERATED FILE -- do not edit.  Produced by pysonic_b200/codegen.py from pysonic_b200/neurons.py.
// Voltage-dependent rate constants (s^-1) of every supported point neuron, as device functions.
#pragma once

#define SONIC_N_NEURONS 16
#define SONIC_MAX_RATES 18

// Arithmetic of the rate expressions.  The reference's formulas are compiled as they are written (neurons.py), on a
// wrapper type: a division is a multiplication by a refined hardware reciprocal (<= 2 ulp; 1 / 0 = inf, 1 / inf = 0),
// the exponential is the branch-free one of the integrator
// (sonic_exp, 1-2 ulp) inside its range and the library's outside.  The averaging kernel evaluates every rate of the
// neuron at 1000 samples x every coverage fraction of every point: on the STN coverage sweep (19 tables) this is the
// second largest kernel of the run.  (sonic_core.h is included before this header.)
struct rd {
    double v;
    __device__ __forceinline__ rd(double x) : v(x) {}
};
static __device__ __forceinline__ rd operator+(rd a, rd b) { return rd(a.v + b.v); }
static __device__ __forceinline__ rd operator-(rd a, rd b) { return rd(a.v - b.v); }
static __device__ __forceinline__ rd operator*(rd a, rd b) { return rd(a.v * b.v); }
// (sonic_rcp is for the integrator's tame arguments: its refinement turns 1 / 0 and 1 / inf into NaN, and the rates do
// reach exp() = inf and x / 0 at their singular points: this one keeps the hardware seed there, without a branch)
static __device__ __forceinline__ double rate_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    const double r2 = fma(r, e, r);
    return r2 == r2 ? r2 : r;          // 1 / 0 = inf and 1 / inf = 0 as the hardware seed gives them
}
static __device__ __forceinline__ rd operator/(rd a, rd b) { return rd(a.v * rate_rcp(b.v)); }
static __device__ __forceinline__ rd operator-(rd a) { return rd(-a.v); }
static __device__ __forceinline__ rd operator+(rd a) { return a; }
static __device__ __forceinline__ bool operator<(rd a, rd b) { return a.v < b.v; }
static __device__ __forceinline__ bool operator>(rd a, rd b) { return a.v > b.v; }
static __device__ __forceinline__ bool operator<=(rd a, rd b) { return a.v <= b.v; }
static __device__ __forceinline__ bool operator>=(rd a, rd b) { return a.v >= b.v; }
static __device__ __forceinline__ rd exp(rd x) { return rd(fabs(x.v) < 690.0 ? sonic_exp(x.v) : ::exp(x.v)); }
// x / (exp(x / y) - 1): naive form of the reference (pneuron.py:351-354), 0/0 at x = 0 kept.
static __device__ __forceinline__ rd vtrap(rd x, rd y) { return x / (exp(x / y) - 1); }

template <int ID> struct SonicRates;

// Net membrane current (mA/m2) of the neurons whose SONIC simulation is supported (NS = number of gating
// states, state k <-> rates 2k and 2k + 1 of SonicRates<ID>); NS = 0: not supported.
template <int ID> struct SonicSim {
    static constexpr int NS = 0;
    static __device__ __forceinline__ double inet(double, const double*) { return 0.0; }
};

// ---- RS ----
template <> struct SonicRates<0> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-56.2);
        const rd TauMax(0.608);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_p = 1.0 / (1 + exp(-(Vm + 35) / 10));
        const rd tau_p = TauMax / (3.3 * exp((Vm + 35) / 20) + exp(-(Vm + 35) / 20));
        r[6] = rd(inf_p / tau_p).v;
        r[7] = rd((1 - inf_p) / tau_p).v;
        (void)VT;
        (void)TauMax;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<0> {
    static constexpr int NS = 4;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 560.0;
        const double ENa = 50.0;
        const double gKdbar = 60.0;
        const double EK = -90.0;
        const double gMbar = 0.75;
        const double gLeak = 0.205;
        const double ELeak = -70.3;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double p = x[3];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- FS ----
template <> struct SonicRates<1> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-57.9);
        const rd TauMax(0.502);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_p = 1.0 / (1 + exp(-(Vm + 35) / 10));
        const rd tau_p = TauMax / (3.3 * exp((Vm + 35) / 20) + exp(-(Vm + 35) / 20));
        r[6] = rd(inf_p / tau_p).v;
        r[7] = rd((1 - inf_p) / tau_p).v;
        (void)VT;
        (void)TauMax;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<1> {
    static constexpr int NS = 4;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 580.0;
        const double ENa = 50.0;
        const double gKdbar = 39.0;
        const double EK = -90.0;
        const double gMbar = 0.787;
        const double gLeak = 0.38;
        const double ELeak = -70.4;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double p = x[3];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- LTS ----
template <> struct SonicRates<2> {
    static constexpr int N = 12;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-50.0);
        const rd TauMax(4.0);
        const rd Vx(-7.0);
        const rd xs = exp(-(Vm + Vx + 132.0) / 16.7) + exp((Vm + Vx + 16.8) / 18.2);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_p = 1.0 / (1 + exp(-(Vm + 35) / 10));
        const rd tau_p = TauMax / (3.3 * exp((Vm + 35) / 20) + exp(-(Vm + 35) / 20));
        r[6] = rd(inf_p / tau_p).v;
        r[7] = rd((1 - inf_p) / tau_p).v;
        const rd inf_s = 1.0 / (1.0 + exp(-(Vm + Vx + 57.0) / 6.2));
        const rd tau_s = 1.0 / 3.7 * (0.612 + 1.0 / xs) * 1e-3;
        r[8] = rd(inf_s / tau_s).v;
        r[9] = rd((1 - inf_s) / tau_s).v;
        const rd inf_u = 1.0 / (1.0 + exp((Vm + Vx + 81.0) / 4.0));
        const rd tau_u = ((Vm + Vx < -80.0) ? 1.0 / 3.7 * exp((Vm + Vx + 467.0) / 66.6) * 1e-3 : 1.0 / 3.7 * (exp(-(Vm + Vx + 22) / 10.5) + 28.0) * 1e-3);
        r[10] = rd(inf_u / tau_u).v;
        r[11] = rd((1 - inf_u) / tau_u).v;
        (void)VT;
        (void)TauMax;
        (void)Vx;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<2> {
    static constexpr int NS = 6;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 500.0;
        const double ENa = 50.0;
        const double gKdbar = 40.0;
        const double EK = -90.0;
        const double gMbar = 0.28;
        const double gLeak = 0.19;
        const double ELeak = -50.0;
        const double gCaTbar = 4.0;
        const double ECa = 120.0;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double p = x[3];
        const double s = x[4];
        const double u = x[5];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak) + gCaTbar * s * s * u * (Vm - ECa);
    }
};

// ---- IB ----
template <> struct SonicRates<3> {
    static constexpr int N = 12;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-56.2);
        const rd TauMax(0.608);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_p = 1.0 / (1 + exp(-(Vm + 35) / 10));
        const rd tau_p = TauMax / (3.3 * exp((Vm + 35) / 20) + exp(-(Vm + 35) / 20));
        r[6] = rd(inf_p / tau_p).v;
        r[7] = rd((1 - inf_p) / tau_p).v;
        r[8] = rd(0.055 * vtrap(-(Vm + 27), 3.8) * 1e3).v;
        r[9] = rd(0.94 * exp(-(Vm + 75) / 17) * 1e3).v;
        r[10] = rd(0.000457 * exp(-(Vm + 13) / 50) * 1e3).v;
        r[11] = rd(0.0065 / (exp(-(Vm + 15) / 28) + 1) * 1e3).v;
        (void)VT;
        (void)TauMax;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<3> {
    static constexpr int NS = 6;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 500.0;
        const double ENa = 50.0;
        const double gKdbar = 50.0;
        const double EK = -90.0;
        const double gMbar = 0.3;
        const double gLeak = 0.1;
        const double ELeak = -70.0;
        const double gCaLbar = 1.0;
        const double ECa = 120.0;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double p = x[3];
        const double q = x[4];
        const double r = x[5];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak) + gCaLbar * q * q * r * (Vm - ECa);
    }
};

// ---- RE ----
template <> struct SonicRates<4> {
    static constexpr int N = 10;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-67.0);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_s = 1.0 / (1.0 + exp(-(Vm + 52.0) / 7.4));
        const rd tau_s = (1 + 0.33 / (exp((Vm + 27.0) / 10.0) + exp(-(Vm + 102.0) / 15.0))) * 1e-3;
        r[6] = rd(inf_s / tau_s).v;
        r[7] = rd((1 - inf_s) / tau_s).v;
        const rd inf_u = 1.0 / (1.0 + exp((Vm + 80.0) / 5.0));
        const rd tau_u = (28.3 + 0.33 / (exp((Vm + 48.0) / 4.0) + exp(-(Vm + 407.0) / 50.0))) * 1e-3;
        r[8] = rd(inf_u / tau_u).v;
        r[9] = rd((1 - inf_u) / tau_u).v;
        (void)VT;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<4> {
    static constexpr int NS = 5;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 2000.0;
        const double ENa = 50.0;
        const double gKdbar = 200.0;
        const double EK = -90.0;
        const double gCaTbar = 30.0;
        const double ECa = 120.0;
        const double gLeak = 0.5;
        const double ELeak = -90.0;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double s = x[3];
        const double u = x[4];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gCaTbar * s * s * u * (Vm - ECa) + gLeak * (Vm - ELeak);
    }
};

// ---- TC ----
template <> struct SonicRates<5> {
    static constexpr int N = 12;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-52.0);
        const rd Vx(0.0);
        const rd xs = exp(-(Vm + Vx + 132.0) / 16.7) + exp((Vm + Vx + 16.8) / 18.2);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        const rd inf_s = 1.0 / (1.0 + exp(-(Vm + Vx + 57.0) / 6.2));
        const rd tau_s = 1.0 / 3.7 * (0.612 + 1.0 / xs) * 1e-3;
        r[6] = rd(inf_s / tau_s).v;
        r[7] = rd((1 - inf_s) / tau_s).v;
        const rd inf_u = 1.0 / (1.0 + exp((Vm + Vx + 81.0) / 4.0));
        const rd tau_u = ((Vm + Vx < -80.0) ? 1.0 / 3.7 * exp((Vm + Vx + 467.0) / 66.6) * 1e-3 : 1.0 / 3.7 * (exp(-(Vm + Vx + 22) / 10.5) + 28.0) * 1e-3);
        r[8] = rd(inf_u / tau_u).v;
        r[9] = rd((1 - inf_u) / tau_u).v;
        const rd inf_o = 1.0 / (1.0 + exp((Vm + 75.0) / 5.5));
        const rd tau_o = 1 / (exp(-14.59 - 0.086 * Vm) + exp(-1.87 + 0.0701 * Vm)) * 1e-3;
        r[10] = rd(inf_o / tau_o).v;
        r[11] = rd((1 - inf_o) / tau_o).v;
        (void)VT;
        (void)Vx;
        (void)Vm; (void)r;
    }
};

// ---- STN ----
template <> struct SonicRates<6> {
    static constexpr int N = 18;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd inf_a = 1 / (1 + exp((Vm - (-45)) / (-14.7)));
        const rd tau_a = 0.001 + 0.001 / (1 + exp(-(Vm - (-40)) / (-0.5)));
        r[0] = rd(inf_a / tau_a).v;
        r[1] = rd((1 - inf_a) / tau_a).v;
        const rd inf_b = 1 / (1 + exp((Vm - (-90)) / (7.5)));
        const rd tau_b = 0.0 + 0.2 / (exp(-(Vm - (-60)) / (-30)) + exp(-(Vm - (-40)) / (10)));
        r[2] = rd(inf_b / tau_b).v;
        r[3] = rd((1 - inf_b) / tau_b).v;
        const rd inf_c = 1 / (1 + exp((Vm - (-30.6)) / (-5)));
        const rd tau_c = 0.045 + 0.01 / (exp(-(Vm - (-27)) / (-20)) + exp(-(Vm - (-50)) / (15)));
        r[4] = rd(inf_c / tau_c).v;
        r[5] = rd((1 - inf_c) / tau_c).v;
        const rd inf_d1 = 1 / (1 + exp((Vm - (-60)) / (7.5)));
        const rd tau_d1 = 0.4 + 0.5 / (exp(-(Vm - (-40)) / (-15)) + exp(-(Vm - (-20)) / (20)));
        r[6] = rd(inf_d1 / tau_d1).v;
        r[7] = rd((1 - inf_d1) / tau_d1).v;
        const rd inf_m = 1 / (1 + exp((Vm - (-40)) / (-8)));
        const rd tau_m = 0.0002 + 0.003 / (1 + exp(-(Vm - (-53)) / (-0.7)));
        r[8] = rd(inf_m / tau_m).v;
        r[9] = rd((1 - inf_m) / tau_m).v;
        const rd inf_h = 1 / (1 + exp((Vm - (-45.5)) / (6.4)));
        const rd tau_h = 0.0 + 0.0245 / (exp(-(Vm - (-50)) / (-15)) + exp(-(Vm - (-50)) / (16)));
        r[10] = rd(inf_h / tau_h).v;
        r[11] = rd((1 - inf_h) / tau_h).v;
        const rd inf_n = 1 / (1 + exp((Vm - (-41)) / (-14)));
        const rd tau_n = 0.0 + 0.011 / (exp(-(Vm - (-40)) / (-40)) + exp(-(Vm - (-40)) / (50)));
        r[12] = rd(inf_n / tau_n).v;
        r[13] = rd((1 - inf_n) / tau_n).v;
        const rd inf_p = 1 / (1 + exp((Vm - (-56)) / (-6.7)));
        const rd tau_p = 0.005 + 0.00033 / (exp(-(Vm - (-27)) / (-10)) + exp(-(Vm - (-102)) / (15)));
        r[14] = rd(inf_p / tau_p).v;
        r[15] = rd((1 - inf_p) / tau_p).v;
        const rd inf_q = 1 / (1 + exp((Vm - (-85)) / (5.8)));
        const rd tau_q = 0.0 + 0.4 / (exp(-(Vm - (-50)) / (-15)) + exp(-(Vm - (-50)) / (16)));
        r[16] = rd(inf_q / tau_q).v;
        r[17] = rd((1 - inf_q) / tau_q).v;
        (void)Vm; (void)r;
    }
};

// ---- FHnode ----
template <> struct SonicRates<7> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd q10(5.799546134795289);
        const rd V0(-70.0);
        r[0] = rd(q10 * 0.36 * vtrap(22. - (Vm - V0), 3.) * 1e3).v;
        r[1] = rd(q10 * 0.4 * vtrap(Vm - V0 - 13., 20.) * 1e3).v;
        r[2] = rd(q10 * 0.1 * vtrap(Vm - V0 + 10.0, 6.) * 1e3).v;
        r[3] = rd(q10 * 4.5 / (exp((45. - (Vm - V0)) / 10.) + 1) * 1e3).v;
        r[4] = rd(q10 * 0.02 * vtrap(35. - (Vm - V0), 10.0) * 1e3).v;
        r[5] = rd(q10 * 0.05 * vtrap(Vm - V0 - 10., 10.) * 1e3).v;
        r[6] = rd(q10 * 0.006 * vtrap(40. - (Vm - V0), 10.0) * 1e3).v;
        r[7] = rd(q10 * 0.09 * vtrap(Vm - V0 + 25., 20.) * 1e3).v;
        (void)q10;
        (void)V0;
        (void)Vm; (void)r;
    }
};

// ---- SWnode ----
template <> struct SonicRates<8> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd am = (126 + 0.363 * Vm) / (1 + exp(-(Vm + 49) / 5.3)) * 1e3;
        const rd bh = 15.6 / (1 + exp(-(Vm + 56) / 10)) * 1e3;
        r[0] = rd(am).v;
        r[1] = rd(am / (exp((Vm + 56.2) / 4.17))).v;
        r[2] = rd(bh / exp((Vm + 74.5) / 5)).v;
        r[3] = rd(bh).v;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<8> {
    static constexpr int NS = 2;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 14450.0;
        const double ENa = 35.64;
        const double gLeak = 1280.0;
        const double ELeak = -80.01;
        const double m = x[0];
        const double h = x[1];
        return gNabar * m * m * h * (Vm - ENa) + gLeak * (Vm - ELeak);
    }
};

// ---- MRGnode ----
template <> struct SonicRates<9> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd q10_mp(3.530825783474764);
        const rd q10_h(5.493344008948558);
        const rd q10_s(1.0);
        const rd Vmh = Vm + 3.;
        const rd Vms = Vm - (-80.);
        r[0] = rd(q10_mp * 1.86 * vtrap(-(Vmh + 18.4), 10.3) * 1e3).v;
        r[1] = rd(q10_mp * 0.086 * vtrap(Vmh + 22.7, 9.16) * 1e3).v;
        r[2] = rd(q10_h * 0.062 * vtrap(Vmh + 111.0, 11.0) * 1e3).v;
        r[3] = rd(q10_h * 2.3 / (1 + exp(-(Vmh + 28.8) / 13.4)) * 1e3).v;
        r[4] = rd(q10_mp * 0.01 * vtrap(-(Vm + 27.), 10.2) * 1e3).v;
        r[5] = rd(q10_mp * 0.00025 * vtrap(Vm + 34., 10.) * 1e3).v;
        r[6] = rd(q10_s * 0.3 / (1 + exp(-(Vms - 27.) / 5.)) * 1e3).v;
        r[7] = rd(q10_s * 0.03 / (1 + exp(-(Vms + 10.) / 1.)) * 1e3).v;
        (void)q10_mp;
        (void)q10_h;
        (void)q10_s;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<9> {
    static constexpr int NS = 4;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNafbar = 30000.0;
        const double gNapbar = 100.0;
        const double ENa = 50.0;
        const double gKsbar = 800.0;
        const double EK = -90.0;
        const double gLeak = 70.0;
        const double ELeak = -90.0;
        const double m = x[0];
        const double h = x[1];
        const double p = x[2];
        const double s = x[3];
        return gNafbar * m * m * m * h * (Vm - ENa) + gNapbar * p * p * p * (Vm - ENa) + gKsbar * s * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- SUseg ----
template <> struct SonicRates<10> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd q10T(1.9331820449317627);
        const rd q10BG(1.9331820449317627);
        const rd FARADAY(96485.3);
        const rd RgT(2570.093793);
        const rd Vmm = (Vm - (-65.)) + (-6.0);
        const rd Vmhh = (Vm - (-65.)) + 6.0;
        const rd xn = (Vm - (-32.)) * FARADAY / RgT * 1e-3;
        const rd xl = (Vm - (-61.)) * FARADAY / RgT * 1e-3;
        r[0] = rd(q10T * 0.32 * vtrap((13.1 - Vmm), 4) * 1e3).v;
        r[1] = rd(q10T * 0.28 * vtrap((Vmm - 40.1), 5) * 1e3).v;
        r[2] = rd(q10T * 0.128 * exp((17.0 - Vmhh) / 18) * 1e3).v;
        r[3] = rd(q10T * 4 / (1 + exp((40.0 - Vmhh) / 5)) * 1e3).v;
        r[4] = rd(q10BG * (0.03 * exp(-(-5.) * 0.4 * xn)) * 1e3).v;
        r[5] = rd(q10BG * (0.03 * exp((-5.) * (1 - 0.4) * xn)) * 1e3).v;
        r[6] = rd(q10BG * (0.001 * exp(-(2.) * 1. * xl)) * 1e3).v;
        r[7] = rd(q10BG * (0.001 * exp((2.) * (1 - 1.) * xl)) * 1e3).v;
        (void)q10T;
        (void)q10BG;
        (void)FARADAY;
        (void)RgT;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<10> {
    static constexpr int NS = 4;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 400.0;
        const double ENa = 55.0;
        const double gKdbar = 400.0;
        const double EK = -90.0;
        const double gLeak = 1.0;
        const double ELeak = -60.069175300110516;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        const double l = x[3];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * l * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- HHseg ----
template <> struct SonicRates<11> {
    static constexpr int N = 6;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd q10(26.1246286895632);
        r[0] = rd(q10 * 0.1 * vtrap(-(Vm + 40), 10) * 1e3).v;
        r[1] = rd(q10 * 4 * exp(-(Vm + 65) / 18) * 1e3).v;
        r[2] = rd(q10 * 0.07 * exp(-(Vm + 65) / 20) * 1e3).v;
        r[3] = rd(q10 * 1.0 / (exp(-(Vm + 35) / 10) + 1) * 1e3).v;
        r[4] = rd(q10 * 0.01 * vtrap(-(Vm + 55), 10) * 1e3).v;
        r[5] = rd(q10 * 0.125 * exp(-(Vm + 65) / 80) * 1e3).v;
        (void)q10;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<11> {
    static constexpr int NS = 3;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 1200.0;
        const double ENa = 50.0;
        const double gKdbar = 360.0;
        const double EK = -77.0;
        const double gLeak = 3.0;
        const double ELeak = -54.3;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- LeechT ----
template <> struct SonicRates<12> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd hb = 1 + exp((Vm - (-50.0)) / 9.0);
        const rd inf_m = 1 / (1 + exp((Vm - (-35.0)) / (-5.0)));
        const rd tau_m = 0.1e-3;
        r[0] = rd(inf_m / tau_m).v;
        r[1] = rd((1 - inf_m) / tau_m).v;
        const rd inf_h = 1 / (hb * hb);
        const rd tau_h = (14.0e-3 - 0.2e-3) / (1 + exp((Vm - (-36.0)) / 3.5)) + 0.2e-3;
        r[2] = rd(inf_h / tau_h).v;
        r[3] = rd((1 - inf_h) / tau_h).v;
        const rd inf_n = 1 / (1 + exp((Vm - (-22.0)) / (-9.0)));
        const rd tau_n = (6.0e-3 - 1.0e-3) / (1 + exp((Vm - (-10.0)) / 10.0)) + 1.0e-3;
        r[4] = rd(inf_n / tau_n).v;
        r[5] = rd((1 - inf_n) / tau_n).v;
        const rd inf_s = 1 / (1 + exp((Vm - (-10.0)) / (-2.8)));
        const rd tau_s = 0.6e-3;
        r[6] = rd(inf_s / tau_s).v;
        r[7] = rd((1 - inf_s) / tau_s).v;
        (void)Vm; (void)r;
    }
};

// ---- LeechP ----
template <> struct SonicRates<13> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        r[0] = rd(-0.03 * (Vm + 28) / (exp(-(Vm + 28) / 15) - 1) * 1e3).v;
        r[1] = rd(2.7 * exp(-(Vm + 53) / 18) * 1e3).v;
        r[2] = rd(0.045 * exp(-(Vm + 58) / 18) * 1e3).v;
        r[3] = rd(0.72 / (exp(-(Vm + 23) / 14) + 1) * 1e3).v;
        r[4] = rd(-0.024 * (Vm - 17) / (exp(-(Vm - 17) / 8) - 1) * 1e3).v;
        r[5] = rd(0.2 * exp(-(Vm + 48) / 35) * 1e3).v;
        r[6] = rd(-1.5 * (Vm - 20) / (exp(-(Vm - 20) / 5) - 1) * 1e3).v;
        r[7] = rd(1.5 * exp(-(Vm + 25) / 10) * 1e3).v;
        (void)Vm; (void)r;
    }
};

// ---- template ----
template <> struct SonicRates<14> {
    static constexpr int N = 6;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        const rd VT(-56.2);
        r[0] = rd(0.32 * vtrap(13 - (Vm - VT), 4) * 1e3).v;
        r[1] = rd(0.28 * vtrap((Vm - VT) - 40, 5) * 1e3).v;
        r[2] = rd(0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3).v;
        r[3] = rd(4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3).v;
        r[4] = rd(0.032 * vtrap(15 - (Vm - VT), 5) * 1e3).v;
        r[5] = rd(0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3).v;
        (void)VT;
        (void)Vm; (void)r;
    }
};

template <> struct SonicSim<14> {
    static constexpr int NS = 3;
    static __device__ __forceinline__ double inet(const double Vm, const double* x) {
        const double gNabar = 560.0;
        const double ENa = 50.0;
        const double gKdbar = 60.0;
        const double EK = -90.0;
        const double gLeak = 0.205;
        const double ELeak = -70.3;
        const double m = x[0];
        const double h = x[1];
        const double n = x[2];
        return gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK) + gLeak * (Vm - ELeak);
    }
};

// ---- pas ----
template <> struct SonicRates<15> {
    static constexpr int N = 0;
    static __device__ __forceinline__ void eval(const double Vm_, double* r) {
        const rd Vm(Vm_);
        (void)Vm; (void)r;
    }
};

static const char* const SONIC_NEURON_NAMES[SONIC_N_NEURONS] = {"RS", "FS", "LTS", "IB", "RE", "TC", "STN", "FHnode", "SWnode", "MRGnode", "SUseg", "HHseg", "LeechT", "LeechP", "template", "pas"};
static const int SONIC_NEURON_NRATES[SONIC_N_NEURONS] = {8, 8, 12, 12, 10, 12, 18, 8, 4, 8, 8, 6, 8, 8, 6, 0};
static const double SONIC_NEURON_CM0[SONIC_N_NEURONS] = {0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.02, 0.025, 0.02, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01};
static const char* const SONIC_NEURON_RATE_NAMES[SONIC_N_NEURONS][SONIC_MAX_RATES] = {
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "alphas", "betas", "alphau", "betau", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "alphaq", "betaq", "alphar", "betar", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphas", "betas", "alphau", "betau", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphas", "betas", "alphau", "betau", "alphao", "betao", "", "", "", "", "", ""},
    {"alphaa", "betaa", "alphab", "betab", "alphac", "betac", "alphad1", "betad1", "alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "alphaq", "betaq"},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphap", "betap", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "", "", "", "", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphap", "betap", "alphas", "betas", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphal", "betal", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "", "", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphas", "betas", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "alphas", "betas", "", "", "", "", "", "", "", "", "", ""},
    {"alpham", "betam", "alphah", "betah", "alphan", "betan", "", "", "", "", "", "", "", "", "", "", "", ""},
    {"", "", "", "", "", "", "", "", "", "", "", "", "", "", "", "", "", ""}
};
#define SONIC_DISPATCH_NEURON(id, CALL) switch (id) { case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break; case 9: CALL(9); break; case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; default: break; }
