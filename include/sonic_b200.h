/* sonic_b200.h -- C ABI of libsonic_b200.so, the B200-native SONIC lookup-table engine.
 *
 * Plain C, caller-owned buffers, no exceptions: every function returns 0 on success or a
 * negative SONIC_E_* code, with a message retrievable through sonic_last_error().  There is
 * no CPU fallback: without a CUDA device every compute entry point fails with
 * SONIC_E_NODEVICE.
 *
 * The reference (tjjlemaire/PySONIC) is pure Python and has no FFI; these entry points are
 * what a binding for its lookup-generation path replaces (citations relative to the
 * reference tree):
 *
 *   sonic_lookup_run   <- scripts/run_lookups.py:98-172  (queue construction, one
 *                         Batch(nbls.computeEffVars, queue) per radius, reshape to
 *                         (na, nf, nA, nQ, nfs) tables) together with everything below it:
 *                         PySONIC/core/batches.py:135-153 (Batch.run),
 *                         PySONIC/core/nbls.py:153-222 (computeEffVars),
 *                         PySONIC/core/bls.py:749-789 (simCycles), :681-718 (derivatives),
 *                         :555-573 (balancedefQS), :334-349 (capacitance),
 *                         PySONIC/core/solvers.py:336-365 (PeriodicSolver.solve),
 *                         :317-330 (isPeriodicallyStable), :150-170 (odeint call),
 *                         PySONIC/core/pneuron.py:268-271 (getEffRates).
 *   sonic_points_run   <- the same for an explicit list of (radius, f, A, Q) points: one
 *                         NeuronalBilayerSonophore.computeEffVars(drive, fs, Qm) call per
 *                         point (nbls.py:153); used by multi-process sharding.
 *   sonic_lookup_run_multi / sonic_points_run_multi / sonic_plan_create_multi
 *                      <- the `for name in args['neuron']` loop of scripts/run_lookups.py:193-238 (one
 *                         computeAStimLookup per neuron): the grids of several neurons go through ONE
 *                         integrator launch, followed by one averaging launch per neuron.  Trajectories
 *                         depend on the sonophore constants and |Q| only, so neurons with the same
 *                         resting charge (same Delta and Lennard-Jones fit, bls.py:49-75) share them.
 *   sonic_plan_*       <- split form of sonic_points_run (upload / launch / fetch) so that a
 *                         caller can keep inputs resident on the device and time the kernels.
 *   sonic_plan_fetch_relcm <- BilayerSonophore.getRelCmCycle (bls.py:806-808) for every point of the
 *                         plan: what scripts/run_Cm_lookups.py:19-64 tabulates.
 *   sonic_pmavg        <- BilayerSonophore.PMavg / v_PMavg (bls.py:390-408): the quadrature behind the
 *                         Lennard-Jones fit of computePMparams (bls.py:410-470), batched over Z.
 *   sonic_simulate     <- NeuronalBilayerSonophore.__simSonic (nbls.py:389-437): effDerivatives (:280-315) with
 *                         Lookup.project / interpolate1D (lookups.py:234-333) integrated over the sample
 *                         times of EventDrivenSolver (solvers.py:445-478), batched over simulations.
 *   sonic_mean_rates   <- PointNeuron.getEffRates(Vm) (pneuron.py:268-271).
 *   sonic_eval_rates   <- the neuron's alphax/betax/xinf/taux methods evaluated elementwise
 *                         (PySONIC/neurons/*.py), for testing the generated device functions.
 *   sonic_neuron_*     <- PySONIC/neurons/__init__.py:24-44 (getPointNeuron) and
 *                         `pneuron.rates` (translators.py:411).
 */
#ifndef SONIC_B200_H
#define SONIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONIC_ABI_VERSION 2
#define SONIC_ABI_MAX_OVERTONES 4   /* charge overtones per point accepted by the *_ex entry points */

#define SONIC_OK 0
#define SONIC_E_NODEVICE (-1)   /* no CUDA device / bad device index */
#define SONIC_E_CUDA (-2)       /* CUDA runtime error */
#define SONIC_E_ARG (-3)        /* invalid argument */
#define SONIC_E_NEURON (-4)     /* unknown neuron id */
#define SONIC_E_ALLOC (-5)      /* allocation failure */

/* Per-point status bits (out_status). */
#define SONIC_STATUS_NOCONV 1u    /* periodic criterion not met at the 11-cycle cap */
#define SONIC_STATUS_ZCLAMP 2u    /* deflection clamped at Zmin in the RHS */
#define SONIC_STATUS_MXSTEP 4u    /* > 500 integrator steps within one output interval */
#define SONIC_STATUS_STEPFAIL 8u  /* repeated error-test / corrector failures */
#define SONIC_STATUS_Z0FAIL 16u   /* no quasi-static equilibrium deflection */
#define SONIC_STATUS_TOLSF 32u    /* tolerance below machine precision */

/* Constants of one bilayer sonophore (one per radius of the lookup).
 * Reference: BilayerSonophore.__init__ / computePMparams, bls.py:115-137,457-470. */
typedef struct {
    double a;      /* in-plane radius (m) */
    double Delta;  /* equilibrium inter-leaflet gap (m) */
    double x0;     /* Lennard-Jones fit of the average intermolecular pressure ... */
    double C;
    double nrep;
    double nattr;
    double Cm0;    /* resting membrane capacitance (F/m2) */
    double depth;  /* embedding tissue depth (m), 0 for a free sonophore */
} SonicBlsParams;

/* Aggregate statistics of one run (all devices summed unless noted). */
typedef struct {
    uint64_t n_points;      /* output points; the counters below cover the trajectories actually
                             * integrated (points that differ by the sign of Q share one) */
    uint64_t n_rhs;         /* right-hand-side evaluations counted as LSODA counts them (3 per
                             * Jacobian); the kernel evaluates n_rhs - 2 n_jac full right-hand sides */
    uint64_t n_jac;         /* finite-difference Jacobian evaluations */
    uint64_t n_steps;       /* accepted integrator steps */
    uint64_t n_cycles;      /* acoustic cycles simulated */
    uint64_t n_launches;    /* kernels launched */
    double ms_z0;           /* kernel time: initial deflection (max over devices) */
    double ms_integrate;    /* kernel time: batched integrator */
    double ms_average;      /* kernel time: fused cycle averaging */
    double ms_total;        /* upload + kernels + download (host wall clock) */
} SonicStats;

int sonic_version(void);
int sonic_device_count(void);
/* Copies the last error message of the calling thread into buf; returns its length. */
int sonic_last_error(char* buf, int len);

int sonic_neuron_count(void);
int sonic_neuron_id(const char* name);                 /* -1 if unknown */
int sonic_neuron_name(int id, char* buf, int len);
int sonic_neuron_nrates(int id);
int sonic_neuron_rate_name(int id, int i, char* buf, int len);

/* out[r * n + k] = rate r of neuron `id` at potential Vm[k] (mV). */
int sonic_eval_rates(int device, int id, const double* Vm, int64_t n, double* out);
/* out[r] = mean over k of rate r at Vm[k]. */
int sonic_mean_rates(int device, int id, const double* Vm, int64_t n, double* out);

/* Effective variables for an explicit list of points.
 *   ia[n], f[n], A[n], Q[n]  : radius index, frequency (Hz), amplitude (Pa), charge (C/m2)
 *   fs[nfs]                  : coverage fractions
 *   out_tables               : [1 + nrates][n][nfs]  ('V' first, then rates in neuron order)
 *   out_ncycles[n], out_status[n], out_tpoint[n] (seconds of device time), out_nrhs[n]:
 *                              any of these may be NULL
 */
int sonic_points_run(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                     const int32_t* ia, const double* f, const double* A, const double* Q,
                     const double* fs, int nfs, double* out_tables, int32_t* out_ncycles,
                     uint32_t* out_status, double* out_tpoint, uint32_t* out_nrhs,
                     SonicStats* stats);

/* Same with charge overtones (NeuronalBilayerSonophore.computeEffVars(..., Qm_overtones),
 * nbls.py:169-201; grid construction run_lookups.py:105-128): the imposed charge of point i is the
 * Fourier-series cycle Q0 + 2 sum_k A_k cos(2 pi j k / 1000 + phi_k), applied sample-and-hold
 * (bls.py:766-768), with overtones[i][k] = {A_k (C/m2), phi_k (rad)}, k < novertones <= 4.
 *   out_tables : [1 + 2 novertones + nrates][n][nfs]  ('V', then A_V1, phi_V1, ..., then rates) */
int sonic_points_run_ex(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                        const int32_t* ia, const double* f, const double* A, const double* Q,
                        int novertones, const double* overtones, const double* fs, int nfs,
                        double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                        double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats);

/* Full grid, reference queue order a > f > A > Q (> fs): out_tables is
 * [1 + nrates][na][nf][nA][nQ][nfs], the per-point outputs are [na][nf][nA][nQ].
 * device_mask: bit d set = use device d (0 = device 0 only). */
int sonic_lookup_run(const SonicBlsParams* radii, int na, const double* f, int nf,
                     const double* A, int nA, const double* Q, int nQ, const double* fs, int nfs,
                     int neuron_id, uint32_t device_mask, double* out_tables,
                     int32_t* out_ncycles, uint32_t* out_status, double* out_tpoint,
                     SonicStats* stats);

/* Several neurons in one batch.  radii holds n_neurons blocks of na entries (block k = the
 * sonophores of neuron_ids[k], which differ in Cm0 and, through the resting charge, in Delta and the
 * Lennard-Jones fit); neuron k has its own charge vector Qcat[Qoff[k] .. Qoff[k + 1]).
 *   out_tables : the neurons' blocks one after the other, block k = [1 + nrates_k][na][nf][nA][nQ_k][nfs]
 *   out_ncycles / out_status / out_tpoint : [sum_k na nf nA nQ_k], same order */
int sonic_lookup_run_multi(const SonicBlsParams* radii, int na, const double* f, int nf,
                           const double* A, int nA, const double* Qcat, const int32_t* Qoff,
                           const double* fs, int nfs, const int32_t* neuron_ids, int n_neurons,
                           uint32_t device_mask, double* out_tables, int32_t* out_ncycles,
                           uint32_t* out_status, double* out_tpoint, SonicStats* stats);

/* Explicit point list over several neurons: radius_neuron[na] gives the slot (index into
 * neuron_ids) each radius entry belongs to, a point's neuron is that of its radius entry.
 *   out_tables : block k = [1 + nrates_k][n_k][nfs] over the points of neuron slot k in input order;
 *   per-point outputs [n] in input order. */
int sonic_points_run_multi(int device, const SonicBlsParams* radii, const int32_t* radius_neuron, int na,
                           const int32_t* neuron_ids, int n_neurons, int64_t n, const int32_t* ia,
                           const double* f, const double* A, const double* Q, const double* fs, int nfs,
                           double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                           double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats);

/* Split form: inputs stay resident on the device between launches. */
typedef struct SonicPlan SonicPlan;
int sonic_plan_create(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                      const int32_t* ia, const double* f, const double* A, const double* Q,
                      const double* fs, int nfs, SonicPlan** plan);
/* Launch on a caller-owned CUDA stream (a cudaStream_t, e.g. torch's current stream) instead of
 * the plan's private one, so that the caller can order and time the launches with its own
 * events.  The stream must belong to the plan's device and outlive the plan. */
int sonic_plan_set_stream(SonicPlan* plan, void* stream);
int sonic_plan_create_ex(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                         const int32_t* ia, const double* f, const double* A, const double* Q,
                         int novertones, const double* overtones, const double* fs, int nfs,
                         SonicPlan** plan);
int sonic_plan_create_multi(int device, const SonicBlsParams* radii, const int32_t* radius_neuron, int na,
                            const int32_t* neuron_ids, int n_neurons, int64_t n, const int32_t* ia,
                            const double* f, const double* A, const double* Q, const double* fs, int nfs,
                            SonicPlan** plan);
int sonic_plan_launch(SonicPlan* plan);   /* asynchronous on the plan's stream */
int sonic_plan_sync(SonicPlan* plan);
int sonic_plan_fetch(SonicPlan* plan, double* out_tables, int32_t* out_ncycles,
                     uint32_t* out_status, double* out_tpoint, uint32_t* out_nrhs);
/* Last-cycle deflection profiles Z[n][1000] (m) of the last launch (bls.py:806-813). */
int sonic_plan_fetch_zprofiles(SonicPlan* plan, double* out_z);
/* Relative capacitance profiles Cm(Z(t)) / Cm0 [n][1000] of the last cycle of the last launch:
 * BilayerSonophore.getRelCmCycle (bls.py:806-808), the per-point output of
 * scripts/run_Cm_lookups.py:19-64. */
int sonic_plan_fetch_relcm(SonicPlan* plan, double* out_cm);
int sonic_plan_stats(SonicPlan* plan, SonicStats* stats);
int sonic_plan_destroy(SonicPlan* plan);

/* Average intermolecular pressure PMavg(Z) (Pa) of a sonophore of radius a and gap Delta for n
 * deflections Z (m): BilayerSonophore.v_PMavg (bls.py:390-408), i.e. scipy.integrate.quad (QUADPACK
 * QAGS, epsabs = epsrel = 1.49e-8, 50 sub-intervals at most) of the leaflet force 2 pi r PMlocal(r) over
 * [0, a], divided by the stretched surface.  out_last (optional): sub-intervals used per deflection. */
int sonic_pmavg(int device, double a, double Delta, int64_t n, const double* Z, double* out_pm, int32_t* out_last);

/* SONIC simulations on the tables, batched: simulation i integrates  dQm/dt = -iNet(V, x) 1e-3,
 * dx_k/dt = alpha_k (1 - x_k) - beta_k x_k  with V, alpha_k, beta_k interpolated linearly in Qm from
 * tab_on[i] while the stimulus is on and from tab_off otherwise ([1 + 2 NS][nQ] each: V, then the
 * (alpha, beta) pair of every gating state in the neuron's rate order; NS = sonic_sim_nstates).
 *   t[nt]        : sample times (non-decreasing; equal consecutive times repeat the state, as the
 *                  reference's solver does at every stimulus transition)
 *   stim_on[nt]  : stimulus state of the interval that ends at sample i
 *   y0[1 + NS]   : initial charge (C/m2) and states
 *   nsub         : integrator sub-steps between consecutive samples
 *   out          : [nsim][nt][1 + NS];  status[nsim]: 1 = the charge left the tabulated range (rows are
 *                  NaN from there on; the reference raises ValueError in that case) */
int sonic_simulate(int device, int neuron_id, int nsim, int nQ, const double* Qref, const double* tab_on,
                   const double* tab_off, int nt, const double* t, const uint8_t* stim_on, const double* y0, int nsub,
                   double* out, int32_t* status);
/* Number of gating states of the neuron's SONIC simulation (0: not supported). */
int sonic_sim_nstates(int neuron_id);

/* Releases the idle workspaces (one device allocation + one pinned host buffer + stream per plan,
 * kept per device between calls; up to 8 kB of device memory per point): call it when no further
 * lookup will be generated. */
int sonic_trim(void);

/* Sustained FP64 FMA throughput of the device (TFLOP/s), measured with a register-resident
 * DFMA kernel: the roofline denominator for the integrator. */
int sonic_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SONIC_B200_H */
