// sonic_b200.cu -- CUDA kernels (sm_100a) and C ABI of the SONIC lookup-table engine.
//
// Kernels
//   sonic_z0_kernel         initial quasi-static deflection of every point (one thread/point)
//   sonic_integrate_kernel  persistent batched integrator: one lane per grid point, lanes
//                           refill themselves from a cost-sorted work queue; wide warps: every tick =
//                           one warp-convergent RHS evaluation + staged per-lane bookkeeping; a lane
//                           alone in its warp: register-resident BDF runs (sonic_core.h); periodic-
//                           convergence test on the fly against the previous cycle
//   sonic_average_kernel    fused cycle averaging: Z(t) -> Cm -> per coverage fraction V(t)
//                           -> generated neuron rate functions -> warp-shuffle means -> coalesced stores
//   sonic_rates_kernel, sonic_mean_rates_kernel   elementwise / mean evaluation of the rate functions
//   sonic_relcm_kernel      relative capacitance profiles of the last cycle (run_Cm_lookups)
//   sonic_pmavg_kernel      intermolecular pressure by QUADPACK QAGS (sonic_quad.h), for the Lennard-Jones fit
//   sonic_simulate_kernel   SONIC simulations on the tables (effective system, one thread per simulation)
//   sonic_stats_kernel      device-side reduction of the per-point counters
//   sonic_dfma_kernel       FP64 FMA throughput microbenchmark (roofline denominator)
//
// Host side: plan objects (device buffers + stream + events), cost model for the work queue
// order, multi-device fan-out for sonic_lookup_run.  See include/sonic_b200.h for the ABI.

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define SONIC_BLOCK 128
#ifndef SONIC_BLOCKS_PER_SM
#define SONIC_BLOCKS_PER_SM 2      /* 3 (168 registers, 72 B of spills) measured slower: 2.71 s vs 2.37 s on C2 */
#endif
#define SONIC_HIST_STRIDE SONIC_BLOCK   /* per-lane indexed storage interleaved across the block */
/* warps with at most this many busy lanes run the nested tick (0 = never) */
#ifndef SONIC_LONE_MAXK
#define SONIC_LONE_MAXK 1
#endif
/* 1 = a warp whose busy set is a single lane runs the nested tick on any SM (measured: C2 2.02 s instead of 1.37 s --
   two instruction streams on one SM thrash its instruction cache); 0 = only in blocks the host marks (SMs that host
   nothing but one-point warps) */
#ifndef SONIC_LONE_DYNAMIC
#define SONIC_LONE_DYNAMIC 0
#endif
/* budgeted warps go to full width once their initial points are done (see the kernel) */
#ifndef SONIC_WIDEN
#define SONIC_WIDEN 1
#endif
#ifndef SONIC_LONE_PARK
#define SONIC_LONE_PARK 1
#endif
/* relative to the tick of a lone lane in the register-resident run with four such warps on its SM (1.106 us): the same with
   eight warps on the SM (1.44 us) */
#ifndef SONIC_SCHED_LONE8_RATIO
#define SONIC_SCHED_LONE8_RATIO 1.30
#endif
/* the points of the full-width queue are short: start-up (initial deflection read, first Adams steps of every cycle) and
   partly empty warps make a lane-tick there cost more than in the calibration runs on long chains (5.5-7.5 us per
   right-hand side measured on the bulk of C2 against 4.8 us) */
/* budgeted warps share their SM with full-width warps and tick slower than in the calibration runs, where every warp of the
   device holds the same number of lanes (scan over five workloads: profiles/README.md) */
#ifndef SONIC_SCHED_STAGED_SLOWDOWN
#define SONIC_SCHED_STAGED_SLOWDOWN 1.3
#endif
#ifndef SONIC_SCHED_QUEUE_OVERHEAD
#define SONIC_SCHED_QUEUE_OVERHEAD 1.6
#endif

#include "../../include/sonic_b200.h"
#include "generated/cost_table.h"
#include "sonic_core.h"
#include "generated/neuron_rates.cuh"
#include "sonic_quad.h"

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int set_err(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return set_err(SONIC_E_CUDA, "%s failed: %s (%s:%d)", #expr,                     \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    } while (0)

// ---------------------------------------------------------------------------------------
// device code
// ---------------------------------------------------------------------------------------
#define SONIC_AVG_WARPS 4
#ifndef SONIC_AVG_MIN_BLOCKS
#define SONIC_AVG_MIN_BLOCKS 1
#endif
#define SONIC_HIST_BYTES (SONIC_H_SIZE * SONIC_BLOCK * sizeof(double))

struct SonicJob {
    const SonicBls* radii;
    const int* order;          // work-queue order (cost-sorted point indices)
    const int* ia;
    const double* f;
    const double* A;
    const double* Q;
    const double* ov;          // [n][nov][2] charge overtones (amplitude, phase), or null
    int nov;
    double* z0;
    double* zbuf;              // [n][1000]
    double* ngbuf;             // [slots][1000]
    int* ncycles;
    unsigned* status;
    unsigned* nfe;
    unsigned* nje;
    unsigned* nsteps;
    double* tpoint;
    unsigned long long* counter;
    const int* warp_first;     // [warps]: first work-queue position of the warp's initial points
    const int* warp_cap;       // [warps]: lanes of the warp allowed to work while those run
    int* block_smid;           // [blocks]: SM each block ran on (placement probe / diagnostics)
    long long n;
    int probe;                 // 1 = record the block placement and return
    const int* block_nested;   // [blocks]: 1 = this block runs the nested tick (plans whose warps hold one point at a time)
    const int* block_group;    // [blocks]: SM group of a block whose SM hosts nothing but one-point warps, else -1
    int* group_left;           // [groups]: warps of the group still on their initial chain
    int widen;                 // 1 = budgeted warps go to full width once their initial points are done
    int lone_park;             // 1 = a one-point warp whose chain is done waits for its SM instead of taking queue points
};

__constant__ SonicTables c_tables;

static __device__ __forceinline__ unsigned long long sonic_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(128) sonic_z0_kernel(SonicJob job) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= job.n) return;
    SonicPoint p;
    const double f = job.f[i];
    sonic_point_init(p, job.radii[job.ia[i]], f, job.A[i], job.Q[i], job.nov,
                     job.nov ? job.ov + (size_t)i * 2 * job.nov : nullptr);
    if (p.nov) sonic_update_charge(p, 0.0);       // initial conditions use the first charge sample (bls.py:766)
    double z0;
    const bool ok = sonic_z0(p, f, &z0);
    job.z0[i] = ok ? z0 : nan("");
}

// OVT: points carry charge overtones (a separate instantiation keeps the Fourier-series charge code out of
// the instruction stream of the common case; the integrator's hot loop competes for the SM's instruction cache)
template <bool OVT>
__global__ void __launch_bounds__(SONIC_BLOCK, SONIC_BLOCKS_PER_SM) sonic_integrate_kernel(SonicJob job) {
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        job.block_smid[blockIdx.x] = (int)smid;
    }
    if (job.probe) {
        // keep every block resident until all have started, as in a real launch (bounded wait:
        // the grid is one resident wave, but never spin forever on that assumption)
        if (threadIdx.x == 0) {
            atomicAdd(job.counter, 1ULL);
            for (int spin = 0; spin < 200000; spin++) {
                if (atomicAdd(job.counter, 0ULL) >= (unsigned long long)gridDim.x) break;
                __nanosleep(200);
            }
        }
        return;
    }
    __shared__ SonicTables tab;
    {
        const double* src = reinterpret_cast<const double*>(&c_tables);
        double* dst = reinterpret_cast<double*>(&tab);
        for (int i = threadIdx.x; i < (int)(sizeof(SonicTables) / sizeof(double)); i += blockDim.x)
            dst[i] = src[i];
    }
    __syncthreads();

    extern __shared__ double hist_s[];   // [SONIC_H_SIZE][SONIC_BLOCK]
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    SonicHist H;
    H.base = hist_s + threadIdx.x;
    SonicSink sink;
    sink.ngbuf = job.ngbuf + slot * SONIC_NPC;
    sink.zbuf = nullptr;

    SonicLane s;
    SonicPoint p;
    long long pt = -1;
    double period = 0.0;
    unsigned long long t_start = 0;
    // Lane budget of this warp.  The expensive points are long serial chains (up to 6e5 ticks):
    // their wall time is chain length x tick latency, and the tick of a warp gets slower with
    // every extra lane in a different integrator phase.  The host therefore hands the most
    // expensive points out first, to warps that keep only `cap` lanes busy while those points
    // run; once they are done the warp works at full width on whatever is left in the queue.
    const int gwarp = (int)(slot >> 5);
#if SONIC_LONE_MAXK > 0
    const bool nested = job.block_nested[blockIdx.x] != 0;
#endif
    int cap = job.warp_cap[gwarp];
    long long q_init = (lane < cap) ? (long long)job.warp_first[gwarp] + lane : -1;
    if (q_init >= job.n) q_init = -1;
    bool exhausted = false;
    // A budgeted warp goes to full width once all of its initial points are done; on an SM that hosts nothing but
    // one-point warps (group >= 0) only when every warp of the SM has got there -- until then it keeps taking single
    // points from the queue, so that the SM runs one instruction stream at a time.
    bool is_init = false;                  // this lane works on one of the warp's initial points
    bool init_done_now = false;            // ... and has just finished it (or found it without equilibrium)
    int init_left = __popc(__ballot_sync(0xffffffffu, q_init >= 0));
    const int group = job.block_group ? job.block_group[blockIdx.x] : -1;
    bool released = init_left == 0;        // nothing (left) to wait for
    int park_spins = 0;
    if (released && group >= 0 && cap > 0 && cap < 32 && lane == 0) atomicSub(&job.group_left[group], 1);
    if (cap == 0) {
        // Parked warp: it shares its SM with the very longest chains of the grid, which then have a
        // scheduler (and the SM's instruction cache) almost to themselves.  It stays out of the way
        // until the shared queue is drained, then leaves.
        if (lane == 0)
            while (*reinterpret_cast<volatile unsigned long long*>(job.counter) < (unsigned long long)job.n) __nanosleep(20000);
        __syncwarp();
        cap = 32;
    }

    while (true) {
        if (!released) {
            init_left -= __popc(__ballot_sync(0xffffffffu, init_done_now));
            init_done_now = false;
            if (init_left <= 0) {
                released = true;
                if (group >= 0 && lane == 0) atomicSub(&job.group_left[group], 1);
            }
        }
        if (released && cap < 32 && job.widen) {
            int left = 0;
            if (group >= 0) {
                if (lane == 0) left = *reinterpret_cast<volatile int*>(job.group_left + group);
                left = __shfl_sync(0xffffffffu, left, 0);
            }
            if (left <= 0) cap = 32;
            else if (job.lone_park && park_spins < 2000000 && __all_sync(0xffffffffu, pt < 0)) {
                // the chains still running on this SM are among the longest of the grid: leave them the SM
                // (a lone lane ticks in 1.1 us with four warps on the SM, 1.44 us with eight).  Bounded wait
                // (~10 s): never spin forever on the assumption that the whole grid is resident.
                park_spins++;
                __nanosleep(4000);
                continue;
            }
        }
        if (pt < 0 && !exhausted && lane < cap) {
            // (re)fill this lane: its initial point first, then the shared work queue
            unsigned long long q;
            if (q_init >= 0) {
                q = (unsigned long long)q_init;
                q_init = -1;
                is_init = true;
            } else {
                q = atomicAdd(job.counter, 1ULL);
                is_init = false;
            }
            if (q < (unsigned long long)job.n) {
                pt = job.order[q];
                const double f = job.f[pt];
                sonic_point_init(p, job.radii[job.ia[pt]], f, job.A[pt], job.Q[pt], OVT ? job.nov : 0,
                                 OVT ? job.ov + (size_t)pt * 2 * job.nov : nullptr);
                period = 1.0 / f;
                sink.zbuf = job.zbuf + pt * SONIC_NPC;
                const double z0 = job.z0[pt];
                t_start = sonic_globaltimer();
                if (z0 != z0) {
                    // no quasi-static equilibrium: report and move on
                    job.ncycles[pt] = 0;
                    job.status[pt] = SONIC_ST_Z0FAIL;
                    job.nfe[pt] = 0; job.nje[pt] = 0; job.nsteps[pt] = 0;
                    job.tpoint[pt] = 0.0;
                    pt = -1;
                    init_done_now = is_init;
                    is_init = false;
                } else {
                    sonic_lane_start(s, H, p, f, z0, sink);
                }
            } else {
                exhausted = true;
            }
        }
        const bool active = pt >= 0;
        const unsigned wmask = __ballot_sync(0xffffffffu, active);
        if (wmask == 0u) {
            // nothing running in this warp: widen it, or stop when the queue is drained
            if (cap < 32) {
                cap = 32;
                continue;
            }
            if (__all_sync(0xffffffffu, exhausted)) break;
            continue;
        }
        // tick until a lane of this warp finishes its point (the set of busy lanes is fixed till then)
        bool fin = false;
#if SONIC_LONE_MAXK > 0
        if ((nested || SONIC_LONE_DYNAMIC) && __popc(wmask) == 1) {
            // a lane alone in its warp (the long chains the host hands out one per warp, on SMs that host
            // nothing else): nothing to share with other lanes, it follows its own path -- register-resident
            // runs at fixed order, generic nested ticks in between
            do {
                if (active) {
                    sonic_lone_advance<OVT>(s, H, &tab, p, sink, period);
                    fin = s.phase == PH_DONE;
                }
            } while (!__any_sync(0xffffffffu, fin));
        } else
#endif
        do {
            if (active) {
                double fv[3];
                if (OVT) sonic_update_charge(p, s.tn);
                if (sonic_rhs(p, s.tn, s.y, fv)) s.status |= SONIC_ST_ZCLAMP;
                sonic_tick(s, H, &tab, p, sink, period, fv, wmask);
                fin = s.phase == PH_DONE;
            }
        } while (!__any_sync(0xffffffffu, fin));
        if (fin) {
            job.ncycles[pt] = s.cyc;
            job.status[pt] = s.status;
            job.nfe[pt] = s.nfe;
            job.nje[pt] = s.nje;
            job.nsteps[pt] = s.nsteps;
            job.tpoint[pt] = (double)(sonic_globaltimer() - t_start) * 1e-9;
            pt = -1;
        }
        if (fin) {
            init_done_now = is_init;
            is_init = false;
        }
    }
}

// Fused cycle averaging.  A warp works through a contiguous range of (point, fs) entries -- a multiple of 32, sized
// on the host so that every SM gets a few dozen ranges whatever the shape of the lookup (many points x one coverage
// fraction, or few points x 100 fractions); the capacitance profile of the current point is staged in shared memory
// and reused for every coverage fraction of the range.  Every table of the output is laid out [point][fs], so the
// entries a warp produces one after the other are consecutive in memory: lane k keeps the results of the k-th entry
// of the current run of 32 in registers, and a full run is written with one coalesced 256-byte store per table.
struct SonicAvgArgs {
    const double* zbuf;        // [n_traj][1000] last-cycle deflections
    const int* ia_out;         // [n_out_all] radius entry of every output point (its own Cm0)
    const double* Q;           // [n_out_all] signed charge of every output point
    const SonicBls* radii;
    const unsigned* status;    // [n_traj]
    const double* fs;
    const double* ov;          // [n_traj][nov][2]
    const int* umap;           // [n_out_all] -> trajectory, or null (identity)
    const int* sel;            // [n] output points of this launch (one neuron of a multi-neuron plan), or null
    double* out;               // [nvar][n][nfs]
    long long n;
    int nfs, nov;
    long long entries_per_chunk;   // multiple of 32
};

template <int NID, bool OVT>
__global__ void __launch_bounds__(32 * SONIC_AVG_WARPS, SONIC_AVG_MIN_BLOCKS) sonic_average_kernel(SonicAvgArgs a) {
    constexpr int NR = SonicRates<NID>::N;
    constexpr int MOV = OVT ? SONIC_MAX_OVERTONES : 0;      // overtone slots (none in the common instantiation)
    constexpr int NV = 1 + 2 * MOV + NR;
    __shared__ double cm_s[SONIC_AVG_WARPS][SONIC_NPC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nov = OVT ? a.nov : 0, nfs = a.nfs;
    const int nvar = 1 + 2 * nov + NR;          // V, (A_Vk, phi_Vk) per overtone, rates
    double* cm = cm_s[warp];
    const long long n = a.n;
    const long long total = n * nfs, epc = a.entries_per_chunk;
    const long long nchunks = (total + epc - 1) / epc;
    const long long nwarps = (long long)gridDim.x * SONIC_AVG_WARPS;
    for (long long chunk = (long long)blockIdx.x * SONIC_AVG_WARPS + warp; chunk < nchunks; chunk += nwarps) {
        const long long e_first = chunk * epc;
        const long long e_last = (e_first + epc < total) ? e_first + epc : total;
        long long e_run = e_first;              // first entry (point * nfs + fs index) of the current run
        int fill = 0;                           // entries of the run computed so far
        double keep[NV];                        // this lane's entry of the run, one value per table
#pragma unroll
        for (int v = 0; v < NV; v++) keep[v] = 0.0;
        long long e = e_first;
        int j0 = (int)(e_first % nfs);
        for (long long pt = e_first / nfs; e < e_last; pt++, j0 = 0) {
            const long long g = a.sel ? a.sel[pt] : pt;         // among all output points of the plan
            const long long u = a.umap ? a.umap[g] : g;         // trajectory the point reads
            const SonicBls b = a.radii[a.ia_out[g]];
            const double a2 = b.a * b.a;
            const double q0 = a.Q[g];
            const double* ovp = nov ? a.ov + (size_t)u * 2 * nov : nullptr;
            const bool bad = (a.status[u] & (SONIC_ST_Z0FAIL | SONIC_ST_STEPFAIL | SONIC_ST_MXSTEP |
                                             SONIC_ST_TOLSF)) != 0;
            const double* z = a.zbuf + u * SONIC_NPC;
            __syncwarp();
            for (int k = lane; k < SONIC_NPC; k += 32) cm[k] = sonic_capacitance(a2, b.Delta, b.Cm0, z[k]);
            __syncwarp();
            const int j_end = (e_last - e < nfs - j0) ? j0 + (int)(e_last - e) : nfs;
            for (int j = j0; j < j_end; j++, e++) {
                const double x = a.fs[j];
                double acc[NV];
#pragma unroll
                for (int v = 0; v < NV; v++) acc[v] = 0.0;
                for (int k = lane; k < SONIC_NPC; k += 32) {
                    // imposed charge of sample k (constant without overtones, nbls.py:169-178)
                    const double q = nov ? sonic_charge_sample(q0, nov, ovp, k) : q0;
                    // spatial average of the capacitance, then membrane potential in mV
                    const double vm = q / (x * cm[k] + (1 - x) * b.Cm0) * 1e3;   // nbls.py:148-151,188
                    double r[NR > 0 ? NR : 1];
                    SonicRates<NID>::eval(vm, r);
                    acc[0] += vm;
                    // Fourier coefficients of the potential, rfft(Vm)[m] (nbls.py:194-201)
                    for (int m = 1; m <= nov; m++) {
                        double sn, cs;
                        sincospi((double)((2 * m * k) % (2 * SONIC_NPC)) * (1.0 / SONIC_NPC), &sn, &cs);
                        acc[2 * m - 1] += vm * cs;
                        acc[2 * m] -= vm * sn;
                    }
#pragma unroll
                    for (int v = 0; v < NR; v++) acc[1 + 2 * MOV + v] += r[v];
                }
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    double t = acc[v];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    acc[v] = t * (1.0 / (double)SONIC_NPC);
                }
                // amplitude-phase form of the overtone coefficients
                for (int m = 1; m <= nov; m++) {
                    const double re = acc[2 * m - 1], im = acc[2 * m];
                    acc[2 * m - 1] = hypot(re, im);
                    acc[2 * m] = atan2(im, re);
                }
                // every lane holds all sums: the lane of this entry keeps them
                if (lane == fill) {
#pragma unroll
                    for (int v = 0; v < NV; v++) keep[v] = bad ? nan("") : acc[v];
                }
                fill++;
                const bool last = e + 1 == e_last;
                if (fill == 32 || last) {
                    // one store per table: slot v of `keep` is table v (V, overtone pairs) or, past the
                    // overtone slots, rate v - 2 (MOV - nov)
                    if (lane < fill) {
#pragma unroll
                        for (int v = 0; v < NV; v++) {
                            const bool used = v <= 2 * nov || v > 2 * MOV;
                            const int tv = v <= 2 * nov ? v : v - 2 * MOV + 2 * nov;
                            if (used) a.out[(long long)tv * n * nfs + e_run + lane] = keep[v];
                        }
                    }
                    e_run += fill;
                    fill = 0;
                }
            }
        }
    }
}

// Sum of the per-trajectory work counters (device-side reduction for SonicStats).
__global__ void __launch_bounds__(256) sonic_stats_kernel(const unsigned* __restrict__ nfe, const unsigned* __restrict__ nje,
                                                          const unsigned* __restrict__ nsteps, const int* __restrict__ ncycles,
                                                          long long n, unsigned long long* __restrict__ out) {
    unsigned long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        s0 += nfe[i]; s1 += nje[i]; s2 += nsteps[i]; s3 += (unsigned long long)ncycles[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, s0); atomicAdd(out + 1, s1); atomicAdd(out + 2, s2); atomicAdd(out + 3, s3);
    }
}

template <int NID>
__global__ void sonic_rates_kernel(const double* __restrict__ vm, long long n, double* __restrict__ out) {
    constexpr int NR = SonicRates<NID>::N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        double r[NR > 0 ? NR : 1];
        SonicRates<NID>::eval(vm[k], r);
#pragma unroll
        for (int v = 0; v < NR; v++) out[(long long)v * n + k] = r[v];
    }
}

// mean of each rate over a potential vector: single block, deterministic tree reduction
template <int NID>
__global__ void __launch_bounds__(256) sonic_mean_rates_kernel(const double* __restrict__ vm, long long n,
                                                              double* __restrict__ out) {
    constexpr int NR = SonicRates<NID>::N;
    constexpr int NS = NR > 0 ? NR : 1;
    __shared__ double part[8][NS];
    double acc[NS];
#pragma unroll
    for (int v = 0; v < NR; v++) acc[v] = 0.0;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) {
        double r[NS];
        SonicRates<NID>::eval(vm[k], r);
#pragma unroll
        for (int v = 0; v < NR; v++) acc[v] += r[v];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < NR; v++) {
        double t = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) part[warp][v] = t;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += part[w][threadIdx.x];
        out[threadIdx.x] = t / (double)n;
    }
}

// Relative capacitance profiles Cm(Z(t)) / Cm0 of the last cycle (bls.py:806-808).
__global__ void __launch_bounds__(256) sonic_relcm_kernel(const double* __restrict__ zbuf, const int* __restrict__ ia,
                                                          const SonicBls* __restrict__ radii, long long n,
                                                          double* __restrict__ out) {
    const long long total = n * SONIC_NPC;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        const SonicBls b = radii[ia[k / SONIC_NPC]];
        out[k] = sonic_capacitance(b.a * b.a, b.Delta, b.Cm0, zbuf[k]) / b.Cm0;
    }
}

// Average intermolecular pressure PMavg(Z) (bls.py:390-408), one thread per deflection value: each runs the
// QAGS sequence of scipy.integrate.quad on the leaflet force integrand (sonic_quad.h).
__global__ void __launch_bounds__(64) sonic_pmavg_kernel(double a, double Delta, long long n, const double* __restrict__ Z,
                                                         double* __restrict__ out, int* __restrict__ last) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l = 0;
    out[i] = sonic_pmavg_point(a, Delta, Z[i], &l);
    if (last) last[i] = l;
}

// ---------------------------------------------------------------------------------------
// SONIC simulation on the tables (nbls.py:280-315,389-437): one thread per simulation integrates the
// effective system  dQm/dt = -iNet(V_eff(Qm), x) 1e-3,  dx_k/dt = alpha_k(Qm) (1 - x_k) - beta_k(Qm) x_k
// whose coefficients are interpolated linearly in the charge from the 1-D lookups of the stimulus-on and
// stimulus-off conditions (Lookup.project('A', .), interpolate1D: lookups.py:234-333), over the sample
// times the reference's EventDrivenSolver produces (solvers.py:445-478).  The effective rates reach
// 1e6 ... 1e16 s^-1 in parts of the tables, so the gates are advanced with the exact solution of their
// linear equation over each sub-step (rates frozen at the mid-point charge) and the charge with the
// mid-point rule: second order, unconditionally stable in the gates.
// ---------------------------------------------------------------------------------------
struct SonicSimArgs {
    const double* Qref;        // [nQ] ascending
    const double* tab_on;      // [nsim][1 + 2 NS][nQ]: V, alpha_k, beta_k at the drive amplitude of each simulation
    const double* tab_off;     // [1 + 2 NS][nQ] at zero amplitude (shared)
    const double* t;           // [nt] sample times
    const unsigned char* on;   // [nt] stimulus state of the interval that ends at sample i
    const double* y0;          // [1 + NS]
    double* out;               // [nsim][nt][1 + NS]
    int* status;               // [nsim]: 0 ok, 1 charge left the tabulated range (the reference raises there)
    int nsim, nQ, nt, nsub;
};

template <int NS>
static __device__ __forceinline__ bool sonic_sim_lookup(const double* __restrict__ tab, const double* __restrict__ Qref, int nQ,
                                                        double Q, double* V, double* al, double* be) {
    if (!(Q >= Qref[0] && Q <= Qref[nQ - 1])) return false;
    int lo = 0, hi = nQ - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (Qref[mid] <= Q) lo = mid;
        else hi = mid;
    }
    const double w = (Q - Qref[lo]) / (Qref[hi] - Qref[lo]);
    *V = tab[lo] + w * (tab[hi] - tab[lo]);
#pragma unroll
    for (int k = 0; k < NS; k++) {
        const double* a = tab + (size_t)(1 + 2 * k) * nQ;
        const double* b = a + nQ;
        al[k] = a[lo] + w * (a[hi] - a[lo]);
        be[k] = b[lo] + w * (b[hi] - b[lo]);
    }
    return true;
}

template <int NID>
__global__ void __launch_bounds__(64) sonic_simulate_kernel(SonicSimArgs a) {
    constexpr int NS = SonicSim<NID>::NS;
    constexpr int NSA = NS > 0 ? NS : 1;
    const int sim = blockIdx.x * blockDim.x + threadIdx.x;
    if (sim >= a.nsim) return;
    const int nQ = a.nQ, NV = 1 + 2 * NS;
    const double* on = a.tab_on + (size_t)sim * NV * nQ;
    double Q = a.y0[0], x[NSA];
#pragma unroll
    for (int k = 0; k < NS; k++) x[k] = a.y0[1 + k];
    double* out = a.out + (size_t)sim * a.nt * (1 + NS);
    out[0] = Q;
#pragma unroll
    for (int k = 0; k < NS; k++) out[1 + k] = x[k];
    int st = 0;
    for (int i = 1; i < a.nt; i++) {
        const double h = a.t[i] - a.t[i - 1];
        if (h > 0.0 && st == 0) {
            const double* tab = a.on[i] ? on : a.tab_off;
            const double hs = h / a.nsub;
            for (int s = 0; s < a.nsub; s++) {
                double V, al[NSA], be[NSA];
                if (!sonic_sim_lookup<NS>(tab, a.Qref, nQ, Q, &V, al, be)) { st = 1; break; }
                const double Qh = Q - 0.5 * hs * (SonicSim<NID>::inet(V, x) * 1e-3);
                if (!sonic_sim_lookup<NS>(tab, a.Qref, nQ, Qh, &V, al, be)) { st = 1; break; }
                double xm[NSA];
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const double r = al[k] + be[k];
                    if (r > 0.0) {
                        const double xinf = al[k] / r;
                        const double e = exp(-0.5 * hs * r);
                        xm[k] = xinf + (x[k] - xinf) * e;
                        x[k] = xinf + (xm[k] - xinf) * e;
                    } else {
                        xm[k] = x[k];
                    }
                }
                Q -= hs * (SonicSim<NID>::inet(V, xm) * 1e-3);
            }
        }
        double* o = out + (size_t)i * (1 + NS);
        o[0] = st ? nan("") : Q;
#pragma unroll
        for (int k = 0; k < NS; k++) o[1 + k] = st ? nan("") : x[k];
    }
    a.status[sim] = st;
}

template <int NID>
static cudaError_t launch_simulate(const SonicSimArgs& a) {
    sonic_simulate_kernel<NID><<<(a.nsim + 63) / 64, 64>>>(a);
    return cudaGetLastError();
}

template <int NID>
static int sim_nstates() { return SonicSim<NID>::NS; }

// FP64 FMA peak: 8 independent register chains per thread.
__global__ void __launch_bounds__(256) sonic_dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// Method coefficient tables: built once per process (thread-safe function-local static), read-only
// afterwards; every device gets the same bytes in its constant memory once.
static const SonicTables& host_tables() {
    static const SonicTables tables = [] {
        SonicTables t;
        memset(&t, 0, sizeof(t));
        sonic_fill_tables(&t);
        return t;
    }();
    return tables;
}

static int check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_err(SONIC_E_NODEVICE, "no CUDA device available (%s): libsonic_b200 has no CPU path",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return set_err(SONIC_E_NODEVICE, "device %d out of range (0..%d)", device, count - 1);
    return SONIC_OK;
}

// Per-device facts that never change: launch geometry of the persistent integrator, and whether
// the device already holds the coefficient tables.
struct DeviceInfo {
    bool ready = false;
    int sm_count = 0;
    int blocks_per_sm = 1;
};
static std::mutex g_dev_mutex;
static DeviceInfo g_dev[64];

static int device_info(int device, DeviceInfo* out) {
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    DeviceInfo& d = g_dev[device & 63];
    if (!d.ready) {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaMemcpyToSymbol(c_tables, &host_tables(), sizeof(SonicTables)));
        CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
        CUDA_TRY(cudaFuncSetAttribute(sonic_integrate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SONIC_HIST_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(sonic_integrate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SONIC_HIST_BYTES));
        int bps = 0, bps_ov = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sonic_integrate_kernel<false>, SONIC_BLOCK,
                                                               SONIC_HIST_BYTES));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_ov, sonic_integrate_kernel<true>, SONIC_BLOCK,
                                                               SONIC_HIST_BYTES));
        if (bps_ov < bps) bps = bps_ov;
        d.blocks_per_sm = bps < 1 ? 1 : bps;
        d.ready = true;
    }
    *out = d;
    return SONIC_OK;
}

// Predicted cost (log of right-hand-side evaluations) of a point, used only to order the work
// queue (longest first) and to size the lane budgets: a table measured on the RS 4-D grid
// (generated/cost_table.h; the charge enters the mechanics through Q^2 only), interpolated
// linearly in log(radius) and log(frequency) between its nodes (clamped outside them), so that
// grids off the table's nodes are ordered by a smooth estimate rather than by their nearest node.
static void bracket_log(const double* nodes, int n, double x, int* i0, int* i1, double* w) {
    if (!(x > nodes[0])) { *i0 = *i1 = 0; *w = 0.0; return; }
    if (!(x < nodes[n - 1])) { *i0 = *i1 = n - 1; *w = 0.0; return; }
    int i = 0;
    while (i + 2 < n && x >= nodes[i + 1]) i++;
    *i0 = i; *i1 = i + 1;
    *w = log(x / nodes[i]) / log(nodes[i + 1] / nodes[i]);
}

static int bin_of(const double* edges, int nbins, double x) {
    int j = 0;
    while (j + 1 < nbins && x >= edges[j + 1]) j++;
    return j;
}

static double predict_log_cost(double a, double f, double A, double Q) {
    int i0, i1, j0, j1;
    double wi, wj;
    bracket_log(SONIC_COST_A, SONIC_COST_NA, a, &i0, &i1, &wi);
    bracket_log(SONIC_COST_F, SONIC_COST_NF, f, &j0, &j1, &wj);
    const int k = bin_of(SONIC_COST_AMP_EDGES, SONIC_COST_NAMP, A);
    const int l = bin_of(SONIC_COST_Q_EDGES, SONIC_COST_NQ, fabs(Q) + 1e-14);
    auto at = [&](int i, int j) {
        return (double)SONIC_COST_LOG[((i * SONIC_COST_NF + j) * SONIC_COST_NAMP + k) * SONIC_COST_NQ + l];
    };
    const double c0 = at(i0, j0) + wj * (at(i0, j1) - at(i0, j0));
    const double c1 = at(i1, j0) + wj * (at(i1, j1) - at(i1, j0));
    return c0 + wi * (c1 - c0);
}

// ---------------------------------------------------------------------------------------
// Workspaces.  A plan lives in ONE device allocation (all its arrays, cycle profiles included)
// plus one pinned host staging buffer, a stream and its timing events.  Workspaces are kept per
// device between calls, so that a one-shot call (sonic_points_run / sonic_lookup_run) performs no
// cudaMalloc / cudaFree / cudaHostAlloc / stream or event creation in the steady state: inputs go
// up in one copy from pinned memory, results come back in one copy into pinned memory.
// sonic_trim() releases the idle ones.
// ---------------------------------------------------------------------------------------
struct Workspace {
    int device = 0;
    char* dptr = nullptr;
    size_t dbytes = 0;
    char* hptr = nullptr;      // pinned
    size_t hbytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};
static std::mutex g_pool_mutex;
static std::vector<Workspace*> g_pool[64];

static void workspace_free(Workspace* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->dptr) cudaFree(w->dptr);
    if (w->hptr) cudaFreeHost(w->hptr);
    for (auto& e : w->ev)
        if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

static cudaError_t workspace_take(int device, size_t dbytes, size_t hbytes, Workspace** out) {
    Workspace* w = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        auto& pool = g_pool[device & 63];
        // best fit among the idle workspaces; else recycle the largest idle one (grown below)
        int best = -1;
        for (int i = 0; i < (int)pool.size(); i++) {
            const bool fits = pool[i]->dbytes >= dbytes && pool[i]->hbytes >= hbytes;
            if (fits && (best < 0 || pool[i]->dbytes < pool[best]->dbytes)) best = i;
        }
        if (best < 0 && !pool.empty()) {
            best = 0;
            for (int i = 1; i < (int)pool.size(); i++)
                if (pool[i]->dbytes > pool[best]->dbytes) best = i;
        }
        if (best >= 0) {
            w = pool[best];
            pool.erase(pool.begin() + best);
        }
    }
    cudaError_t e = cudaSuccess;
    if (!w) {
        w = new Workspace();
        w->device = device;
        e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
        for (auto& ev : w->ev)
            if (e == cudaSuccess) e = cudaEventCreate(&ev);
    }
    if (e == cudaSuccess && w->dbytes < dbytes) {
        if (w->dptr) cudaFree(w->dptr);
        w->dptr = nullptr;
        w->dbytes = 0;
        e = cudaMalloc(reinterpret_cast<void**>(&w->dptr), dbytes);
        if (e == cudaSuccess) w->dbytes = dbytes;
    }
    if (e == cudaSuccess && w->hbytes < hbytes) {
        if (w->hptr) cudaFreeHost(w->hptr);
        w->hptr = nullptr;
        w->hbytes = 0;
        e = cudaHostAlloc(reinterpret_cast<void**>(&w->hptr), hbytes, cudaHostAllocDefault);
        if (e == cudaSuccess) w->hbytes = hbytes;
    }
    if (e != cudaSuccess) {
        workspace_free(w);
        return e;
    }
    *out = w;
    return cudaSuccess;
}

static void workspace_give(Workspace* w) {
    if (!w) return;
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    g_pool[w->device & 63].push_back(w);
}

// Block placement of the persistent grid (SM of every block), probed once per (device, grid size,
// size of the initial-deflection launch that precedes it) and reused: the hardware's block
// distribution for a given launch sequence is reproducible.
struct ProbeKey {
    int device, grid;
    long long n;
    bool operator<(const ProbeKey& o) const {
        return device != o.device ? device < o.device : grid != o.grid ? grid < o.grid : n < o.n;
    }
};
static std::mutex g_probe_mutex;
static std::map<ProbeKey, std::vector<int>> g_probe;

struct SonicPlan {
    int device = 0;
    int na = 0, nfs = 0;
    int nov = 0;                      // charge overtones per point
    std::vector<int> neurons;         // neuron ids of the plan (one for the classic entry points)
    std::vector<long long> n_out_k;   // output points per neuron
    std::vector<size_t> out_off_k;    // offset (doubles) of each neuron's [nvar][n_out_k][nfs] block
    size_t out_count = 0;             // doubles in all table blocks
    long long n = 0;                  // points integrated (unique trajectories)
    long long n_out = 0;              // output points (>= n: +Q / -Q pairs and neurons with the same
                                      // sonophore constants share a trajectory)
    std::vector<int> umap;            // [n_out] -> trajectory (empty = identity)
    long long slots = 0;
    int grid = 0, lanes_per_warp = 32, nwarps = 0;
    unsigned long long n_initial = 0; // work-queue positions handed out statically (counter start)
    std::vector<int> probe_smid;
    Workspace* ws = nullptr;
    cudaStream_t stream = nullptr;    // ws->stream unless the caller supplied one
    // device arrays (all inside ws->dptr)
    SonicBls* d_radii = nullptr;
    int *d_order = nullptr, *d_ia = nullptr, *d_ia_out = nullptr, *d_umap = nullptr, *d_sel = nullptr;
    int *d_warp_first = nullptr, *d_warp_cap = nullptr, *d_block_smid = nullptr, *d_ncycles = nullptr;
    int* d_block_nested = nullptr;
    int *d_block_group = nullptr, *d_group_left0 = nullptr, *d_group_left = nullptr;
    int n_groups = 0;
    double *d_f = nullptr, *d_A = nullptr, *d_Q = nullptr, *d_fs = nullptr, *d_ov = nullptr, *d_Qout = nullptr;
    double *d_z0 = nullptr, *d_zbuf = nullptr, *d_ngbuf = nullptr, *d_tpoint = nullptr, *d_out = nullptr;
    unsigned *d_status = nullptr, *d_nfe = nullptr, *d_nje = nullptr, *d_nsteps = nullptr;
    unsigned long long *d_counter = nullptr, *d_counter0 = nullptr, *d_stats = nullptr;
    // one contiguous device range holds everything sonic_plan_fetch brings back
    char* d_result = nullptr;
    size_t result_bytes = 0, r_out = 0, r_ncycles = 0, r_status = 0, r_tpoint = 0, r_nfe = 0, r_stats = 0;
    size_t in_bytes = 0;              // staged input bytes (pinned area [0, in_bytes))
    bool result_on_host = false;      // pinned copy of the result range is current
    uint64_t launches = 0;
    double ms_upload = 0.0;
    bool launched = false;
};

static int plan_free(SonicPlan* p) {
    if (!p) return SONIC_OK;
    if (p->ws) {
        cudaSetDevice(p->device);
        cudaStreamSynchronize(p->stream);
        if (p->stream != p->ws->stream) cudaStreamSynchronize(p->ws->stream);
        workspace_give(p->ws);
    }
    delete p;
    return SONIC_OK;
}

template <int NID>
static cudaError_t launch_average(SonicPlan* p, int k) {
    SonicAvgArgs a;
    a.zbuf = p->d_zbuf; a.ia_out = p->d_ia_out; a.Q = p->d_Qout; a.radii = p->d_radii; a.status = p->d_status;
    a.fs = p->d_fs; a.ov = p->d_ov; a.umap = p->umap.empty() ? nullptr : p->d_umap;
    long long first = 0;
    for (int j = 0; j < k; j++) first += p->n_out_k[j];
    a.sel = p->neurons.size() > 1 ? p->d_sel + first : nullptr;
    a.out = p->d_out + p->out_off_k[k];
    a.n = p->n_out_k[k];
    a.nfs = p->nfs; a.nov = p->nov;
    if (a.n == 0) return cudaSuccess;
    // entries per warp range: a multiple of 32 (full 256-byte runs), about 32 ranges per SM
    {
        const long long total = a.n * p->nfs;
        long long m = total / (32LL * 148 * 32);
        if (m < 1) m = 1;
        a.entries_per_chunk = 32 * m;
    }
    const long long nchunks = (a.n * p->nfs + a.entries_per_chunk - 1) / a.entries_per_chunk;
    long long blocks = (nchunks + SONIC_AVG_WARPS - 1) / SONIC_AVG_WARPS;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    if (p->nov) sonic_average_kernel<NID, true><<<(int)blocks, 32 * SONIC_AVG_WARPS, 0, p->stream>>>(a);
    else sonic_average_kernel<NID, false><<<(int)blocks, 32 * SONIC_AVG_WARPS, 0, p->stream>>>(a);
    return cudaGetLastError();
}

template <int NID>
static void launch_rates(const double* vm, long long n, double* out, bool mean) {
    if (mean)
        sonic_mean_rates_kernel<NID><<<1, 256>>>(vm, n, out);
    else
        sonic_rates_kernel<NID><<<(int)std::min<long long>((n + 255) / 256, 4096), 256>>>(vm, n, out);
}

// Per-trajectory rows [n][1000] on the device -> per-output-point rows on the host.
static int fetch_rows_expanded(SonicPlan* p, const double* d_src, double* out) {
    if (p->umap.empty()) {
        CUDA_TRY(cudaMemcpyAsync(out, d_src, (size_t)p->n * SONIC_NPC * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        return SONIC_OK;
    }
    std::vector<double> tmp((size_t)p->n * SONIC_NPC);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), d_src, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (long long i = 0; i < p->n_out; i++)
        memcpy(out + (size_t)i * SONIC_NPC, tmp.data() + (size_t)p->umap[i] * SONIC_NPC, SONIC_NPC * sizeof(double));
    return SONIC_OK;
}

// |Q| as the integrator sees it: rounded to 36 significant bits (1.5e-11 relative, four orders
// below the integrator's tolerance), so that grid values such as the -3 and +3 nC/cm2 of an
// np.arange, which differ in their last bits, select the same trajectory -- and so that a point
// gets the same trajectory whether it is computed alone, inside a grid, or with zero-amplitude
// charge overtones.  The averaging kernel uses the exact charges.
static inline uint64_t dbits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double qint(double q) {
    uint64_t u = dbits(fabs(q));
    u = (u + 0x8000ULL) & ~0xFFFFULL;
    memcpy(&q, &u, 8);
    return q;
}

// Build a plan.  `radius_neuron[na]` (index into `neurons`) says which neuron each radius entry
// belongs to; null = all entries belong to neurons[0].
static int plan_build(int device, const SonicBlsParams* radii, const int32_t* radius_neuron, int na,
                      const int32_t* neurons, int nn, int64_t n_in, const int32_t* ia_in, const double* f_in,
                      const double* A_in, const double* Q_in, int novertones, const double* overtones,
                      const double* fs, int nfs, SonicPlan** out_plan) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!radii || na <= 0 || n_in <= 0 || !ia_in || !f_in || !A_in || !Q_in || !fs || nfs <= 0 || !out_plan || !neurons || nn <= 0)
        return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    for (int k = 0; k < nn; k++)
        if (neurons[k] < 0 || neurons[k] >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", neurons[k]);
    if (n_in > 0x7fffffffLL) return set_err(SONIC_E_ARG, "too many points for one plan (%lld)", (long long)n_in);
    if (novertones < 0 || novertones > SONIC_MAX_OVERTONES || (novertones > 0 && !overtones))
        return set_err(SONIC_E_ARG, "invalid charge overtones (0..%d per point, array required)", SONIC_MAX_OVERTONES);
    if (novertones > 0 && nn > 1) return set_err(SONIC_E_ARG, "charge overtones are not supported in multi-neuron plans");
    for (int64_t i = 0; i < n_in; i++) {
        if (ia_in[i] < 0 || ia_in[i] >= na) return set_err(SONIC_E_ARG, "radius index out of range at point %lld", (long long)i);
        if (!(f_in[i] > 0.)) return set_err(SONIC_E_ARG, "frequency must be strictly positive (point %lld)", (long long)i);
        if (!(A_in[i] >= 0.)) return set_err(SONIC_E_ARG, "amplitude must be positive or null (point %lld)", (long long)i);
    }
    for (int i = 0; i < na; i++) {
        if (!(radii[i].a > 0.) || !(radii[i].Delta > 0.) || !(radii[i].Cm0 > 0.))
            return set_err(SONIC_E_ARG, "invalid sonophore constants for radius %d", i);
        if (radius_neuron && (radius_neuron[i] < 0 || radius_neuron[i] >= nn))
            return set_err(SONIC_E_ARG, "radius %d refers to neuron slot %d of %d", i, radius_neuron[i], nn);
    }
    DeviceInfo dev;
    rc = device_info(device, &dev);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    const auto t0 = std::chrono::steady_clock::now();

    // ---- trajectories: the mechanics see the sonophore constants (not Cm0, not the neuron) and the
    // charge through Q^2 only (bls.py:482-491), so output points that differ by the sign of Q, or by
    // the neuron when two neurons share the constants, share one trajectory bit for bit (the
    // reference integrates each of them).  Integrate each (constants, f, A, |Q|) once; the averaging
    // kernel then applies each output point's own signed charge, capacitance and rate functions.
    std::vector<int> mech(na);          // radius entry -> first entry with the same mechanics
    for (int i = 0; i < na; i++) {
        mech[i] = i;
        for (int j = 0; j < i; j++)
            if (radii[j].a == radii[i].a && radii[j].Delta == radii[i].Delta && radii[j].x0 == radii[i].x0 &&
                radii[j].C == radii[i].C && radii[j].nrep == radii[i].nrep && radii[j].nattr == radii[i].nattr &&
                radii[j].depth == radii[i].depth) { mech[i] = j; break; }
    }
    std::vector<int32_t> u_ia;
    std::vector<double> u_f, u_A, u_Q;
    std::vector<int> umap;
    int64_t n = n_in;
    if (novertones > 0) {
        u_ia.assign(ia_in, ia_in + n_in); u_f.assign(f_in, f_in + n_in); u_A.assign(A_in, A_in + n_in);
        u_Q.resize(n_in);
        for (int64_t i = 0; i < n_in; i++) u_Q[i] = copysign(qint(Q_in[i]), Q_in[i]);
    } else {
        // open-addressing table on (mechanics, f, A, |Q|)
        size_t cap = 16;
        while (cap < (size_t)n_in * 2) cap <<= 1;
        std::vector<int> slot(cap, -1);
        umap.resize(n_in);
        u_ia.reserve(n_in); u_f.reserve(n_in); u_A.reserve(n_in); u_Q.reserve(n_in);
        for (int64_t i = 0; i < n_in; i++) {
            const int m = mech[ia_in[i]];
            const double qi = qint(Q_in[i]);
            const uint64_t kf = dbits(f_in[i]), kA = dbits(A_in[i]), kq = dbits(qi);
            uint64_t h = 1469598103934665603ULL;
            for (uint64_t v : {(uint64_t)m, kf, kA, kq}) { h ^= v; h *= 1099511628211ULL; h ^= h >> 29; }
            size_t s = (size_t)h & (cap - 1);
            int found = -1;
            while (slot[s] >= 0) {
                const int t = slot[s];
                if (u_ia[t] == m && dbits(u_f[t]) == kf && dbits(u_A[t]) == kA && dbits(u_Q[t]) == kq) { found = t; break; }
                s = (s + 1) & (cap - 1);
            }
            if (found < 0) {
                found = (int)u_ia.size();
                slot[s] = found;
                u_ia.push_back(m); u_f.push_back(f_in[i]); u_A.push_back(A_in[i]); u_Q.push_back(qi);
            }
            umap[i] = found;
        }
        n = (int64_t)u_ia.size();
        bool identity = n == n_in;
        for (int i = 0; identity && i < na; i++) identity = mech[i] == i;
        if (identity) umap.clear();       // nothing to share: trajectories are the points, in order
    }
    const int32_t* ia = u_ia.data();
    const double *f = u_f.data(), *A = u_A.data(), *Q = u_Q.data();

    std::unique_ptr<SonicPlan, int (*)(SonicPlan*)> guard(new SonicPlan(), plan_free);
    SonicPlan* p = guard.get();
    p->device = device;
    p->na = na;
    p->nfs = nfs;
    p->n = n;
    p->n_out = n_in;
    p->nov = novertones;
    p->neurons.assign(neurons, neurons + nn);
    p->umap = umap;

    // output points per neuron, in input order within a neuron
    std::vector<int> sel;
    p->n_out_k.assign(nn, 0);
    if (nn > 1) {
        std::vector<std::vector<int>> by(nn);
        for (int64_t i = 0; i < n_in; i++) by[radius_neuron ? radius_neuron[ia_in[i]] : 0].push_back((int)i);
        for (int k = 0; k < nn; k++) {
            p->n_out_k[k] = (long long)by[k].size();
            sel.insert(sel.end(), by[k].begin(), by[k].end());
        }
    } else {
        p->n_out_k[0] = n_in;
    }
    p->out_off_k.assign(nn, 0);
    for (int k = 0; k < nn; k++) {
        p->out_off_k[k] = p->out_count;
        p->out_count += (size_t)(1 + 2 * novertones + SONIC_NEURON_NRATES[neurons[k]]) * p->n_out_k[k] * nfs;
    }

    // ---- launch geometry of the persistent integrator
    const long long max_blocks = (long long)dev.sm_count * dev.blocks_per_sm;
    const long long warps_total = max_blocks * (SONIC_BLOCK / 32);
    // few points: spread them over as many warps as possible (the chain of one point is
    // serial, so idle lanes cost nothing while extra warps shorten the critical path)
    long long lpw = (n + warps_total - 1) / warps_total;
    if (lpw > 32) lpw = 32;
    if (lpw < 1) lpw = 1;
    p->lanes_per_warp = (int)lpw;
    const long long need_warps = (n + lpw - 1) / lpw;
    long long blocks = (need_warps + (SONIC_BLOCK / 32) - 1) / (SONIC_BLOCK / 32);
    if (blocks > max_blocks) blocks = max_blocks;
    p->grid = (int)blocks;
    p->slots = blocks * SONIC_BLOCK;
    const int warps_per_block = SONIC_BLOCK / 32;
    const int nwarps = (int)blocks * warps_per_block;
    p->nwarps = nwarps;

    // ---- work-queue order: predicted cost, longest first (bucket sort on 1/64 log units, stable)
    std::vector<int> order(n);
    std::vector<double> cost(n);
    double cmax = -1e300, cmin = 1e300;
    for (int64_t i = 0; i < n; i++) {
        double c = predict_log_cost(radii[ia[i]].a, f[i], A[i], Q[i]);
        if (novertones > 0) {
            // a sample-and-hold charge restarts the integrator a thousand times per cycle: about
            // 8e4 right-hand sides per cycle whatever the drive, so the chain length is set by the
            // number of cycles (11 in the noise regime A < 8 kPa, 3 otherwise; measured on RS)
            c = log(8e4 * (A[i] < 8e3 ? 11.0 : 3.0)) + 0.01 * c;
        }
        cost[i] = c;
        cmax = c > cmax ? c : cmax;
        cmin = c < cmin ? c : cmin;
    }
    {
        const int nb = (int)((cmax - cmin) * 64.0) + 2;
        std::vector<int> start(nb + 1, 0);
        std::vector<int> bucket(n);
        for (int64_t i = 0; i < n; i++) {
            bucket[i] = (int)((cmax - cost[i]) * 64.0);
            start[bucket[i] + 1]++;
        }
        for (int b = 0; b < nb; b++) start[b + 1] += start[b];
        for (int64_t i = 0; i < n; i++) order[start[bucket[i]]++] = (int)i;
    }

    // ---- layout: [inputs | warp budgets] (one upload) [work] [results] (one download) [profiles]
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; const size_t o = off; off += bytes; return o; };
    const size_t o_radii = take(na * sizeof(SonicBls)), o_order = take(n * sizeof(int)), o_ia = take(n * sizeof(int));
    const size_t o_f = take(n * sizeof(double)), o_A = take(n * sizeof(double)), o_Q = take(n * sizeof(double));
    const size_t o_fs = take(nfs * sizeof(double));
    const size_t o_ov = take((size_t)n * 2 * std::max(novertones, 1) * sizeof(double));
    const size_t o_Qout = take(n_in * sizeof(double)), o_iaout = take(n_in * sizeof(int));
    const size_t o_umap = take((umap.empty() ? 1 : n_in) * sizeof(int)), o_sel = take((sel.empty() ? 1 : n_in) * sizeof(int));
    const size_t o_counter0 = take(sizeof(unsigned long long));
    const size_t o_wfirst = take(nwarps * sizeof(int)), o_wcap = take(nwarps * sizeof(int));
    const size_t o_nested = take(blocks * sizeof(int));
    const size_t o_bgroup = take(blocks * sizeof(int)), o_group0 = take(blocks * sizeof(int));
    const size_t in_bytes = (off + 255) & ~(size_t)255;
    off = in_bytes;
    const size_t o_z0 = take(n * sizeof(double)), o_nje = take(n * sizeof(unsigned)), o_nsteps = take(n * sizeof(unsigned));
    const size_t o_counter = take(sizeof(unsigned long long)), o_smid = take(blocks * sizeof(int));
    const size_t o_group = take(blocks * sizeof(int));
    const size_t o_result = take(0);
    const size_t r0 = off;
    const size_t o_out = take(p->out_count * sizeof(double)), o_ncyc = take(n * sizeof(int)), o_status = take(n * sizeof(unsigned));
    const size_t o_tpoint = take(n * sizeof(double)), o_nfe = take(n * sizeof(unsigned)), o_stats = take(8 * sizeof(unsigned long long));
    const size_t result_bytes = ((off + 255) & ~(size_t)255) - r0;
    off = r0 + result_bytes;
    const size_t o_zbuf = take((size_t)n * SONIC_NPC * sizeof(double));
    const size_t o_ngbuf = take((size_t)p->slots * SONIC_NPC * sizeof(double));
    const size_t dbytes = off;
    (void)o_result;

    cudaError_t e = workspace_take(device, dbytes, in_bytes + result_bytes, &p->ws);
    if (e != cudaSuccess)
        return set_err(e == cudaErrorMemoryAllocation ? SONIC_E_ALLOC : SONIC_E_CUDA, "plan creation failed: %s (%zu MB of device memory)",
                       cudaGetErrorString(e), dbytes >> 20);
    Workspace* ws = p->ws;
    p->stream = ws->stream;
    char* D = ws->dptr;
    char* H = ws->hptr;
    p->d_radii = (SonicBls*)(D + o_radii); p->d_order = (int*)(D + o_order); p->d_ia = (int*)(D + o_ia);
    p->d_f = (double*)(D + o_f); p->d_A = (double*)(D + o_A); p->d_Q = (double*)(D + o_Q); p->d_fs = (double*)(D + o_fs);
    p->d_ov = (double*)(D + o_ov); p->d_Qout = (double*)(D + o_Qout); p->d_ia_out = (int*)(D + o_iaout);
    p->d_umap = (int*)(D + o_umap); p->d_sel = (int*)(D + o_sel);
    p->d_warp_first = (int*)(D + o_wfirst); p->d_warp_cap = (int*)(D + o_wcap);
    p->d_block_nested = (int*)(D + o_nested);
    p->d_block_group = (int*)(D + o_bgroup); p->d_group_left0 = (int*)(D + o_group0); p->d_group_left = (int*)(D + o_group);
    p->n_groups = (int)blocks;
    p->d_z0 = (double*)(D + o_z0); p->d_nje = (unsigned*)(D + o_nje); p->d_nsteps = (unsigned*)(D + o_nsteps);
    p->d_counter0 = (unsigned long long*)(D + o_counter0);
    p->d_counter = (unsigned long long*)(D + o_counter); p->d_block_smid = (int*)(D + o_smid);
    p->d_out = (double*)(D + o_out); p->d_ncycles = (int*)(D + o_ncyc); p->d_status = (unsigned*)(D + o_status);
    p->d_tpoint = (double*)(D + o_tpoint); p->d_nfe = (unsigned*)(D + o_nfe); p->d_stats = (unsigned long long*)(D + o_stats);
    p->d_zbuf = (double*)(D + o_zbuf); p->d_ngbuf = (double*)(D + o_ngbuf);
    p->d_result = D + r0;
    p->result_bytes = result_bytes;
    p->r_out = o_out - r0; p->r_ncycles = o_ncyc - r0; p->r_status = o_status - r0; p->r_tpoint = o_tpoint - r0;
    p->r_nfe = o_nfe - r0; p->r_stats = o_stats - r0;
    p->in_bytes = in_bytes;

    // ---- stage the inputs in pinned memory
    {
        SonicBls* hb = (SonicBls*)(H + o_radii);
        for (int i = 0; i < na; i++) {
            hb[i].a = radii[i].a; hb[i].Delta = radii[i].Delta; hb[i].x0 = radii[i].x0; hb[i].C = radii[i].C;
            hb[i].nrep = radii[i].nrep; hb[i].nattr = radii[i].nattr; hb[i].Cm0 = radii[i].Cm0;
            hb[i].depth = radii[i].depth;
        }
        memcpy(H + o_order, order.data(), n * sizeof(int));
        memcpy(H + o_ia, ia, n * sizeof(int));
        memcpy(H + o_f, f, n * sizeof(double));
        memcpy(H + o_A, A, n * sizeof(double));
        memcpy(H + o_Q, Q, n * sizeof(double));
        memcpy(H + o_fs, fs, nfs * sizeof(double));
        if (novertones > 0) memcpy(H + o_ov, overtones, (size_t)n * 2 * novertones * sizeof(double));
        memcpy(H + o_Qout, Q_in, n_in * sizeof(double));
        memcpy(H + o_iaout, ia_in, n_in * sizeof(int));
        if (!umap.empty()) memcpy(H + o_umap, umap.data(), n_in * sizeof(int));
        if (!sel.empty()) memcpy(H + o_sel, sel.data(), n_in * sizeof(int));
    }

    // ---- block placement: cached per (device, grid, n), probed once otherwise.  The probe is the
    // same kernel in the same launch configuration, preceded by the initial-deflection kernel as in
    // a real launch (where its last blocks retire decides which SMs take the first integrator
    // blocks), and returns after recording the SM of every block.  If a real launch ever lands
    // differently only the schedule quality suffers, never the results.
    std::vector<int> smid;
    {
        std::lock_guard<std::mutex> lk(g_probe_mutex);
        auto it = g_probe.find(ProbeKey{device, p->grid, (long long)n});
        if (it != g_probe.end()) smid = it->second;
    }
    const size_t budget_off = o_counter0;       // [counter0 | warp_first | warp_cap] go up after the budgets are known
    if (smid.empty()) {
        smid.assign(blocks, 0);
        CUDA_TRY(cudaMemcpyAsync(D, H, budget_off, cudaMemcpyHostToDevice, p->stream));
        SonicJob probe;
        memset(&probe, 0, sizeof(probe));
        probe.block_smid = p->d_block_smid;
        probe.counter = p->d_counter;
        probe.probe = 1;
        CUDA_TRY(cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned long long), p->stream));
        SonicJob zj;
        memset(&zj, 0, sizeof(zj));
        zj.radii = p->d_radii; zj.ia = p->d_ia; zj.f = p->d_f; zj.A = p->d_A; zj.Q = p->d_Q;
        zj.z0 = p->d_z0; zj.n = p->n; zj.ov = p->d_ov; zj.nov = p->nov;
        sonic_z0_kernel<<<(unsigned)((p->n + 127) / 128), 128, 0, p->stream>>>(zj);
        if (p->nov) sonic_integrate_kernel<true><<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(probe);
        else sonic_integrate_kernel<false><<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(probe);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(smid.data(), p->d_block_smid, blocks * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        std::lock_guard<std::mutex> lk(g_probe_mutex);
        g_probe[ProbeKey{device, p->grid, (long long)n}] = smid;
    }

    // ---- lane budgets (see the kernel).  Warps are walked SM by SM (all warps of all blocks of one
    // SM, then the next SM), each taking the next `cap` points of the sorted list, so that the
    // most expensive points end up alone in their warp on SMs that host nothing but such warps.
    // A warp with k busy lanes ticks in t(k) (measured table below, 1.9 us for k = 1 to 4.8 us
    // for k = 32); every chain c should finish within the deadline T, so a warp whose most
    // expensive point has predicted chain length c may keep k lanes busy with c t(k) <= T.
    int* wfirst = (int*)(H + o_wfirst);
    int* wcap = (int*)(H + o_wcap);
    for (int w = 0; w < nwarps; w++) { wfirst[w] = (int)n; wcap[w] = (int)lpw; }
    {
        std::vector<int> border(blocks);
        for (int b = 0; b < (int)blocks; b++) border[b] = b;
        std::stable_sort(border.begin(), border.end(), [&](int x, int y) { return smid[x] < smid[y]; });
        // Tick times, measured (tools/gpu_tk.py, tools/gpu_lone_rates.py: every warp of the device with k busy lanes on
        // chains of the same kind), in units of t_f = 1.106 us, the tick of a lone lane in the register-resident run on an
        // SM that hosts four such warps, one per scheduler: R8 = the same with eight warps on the SM (1.44 us), tk[k] = a
        // staged warp with k busy lanes (1.65 us for k = 1, 2.25, 2.70, 3.05, ... 4.81 us for k = 32).
        static const double tk_us[33] = {0, 1.645, 2.250, 2.701, 3.051, 3.32, 3.549, 3.73, 3.879, 3.99, 4.09, 4.18, 4.263,
                                         4.33, 4.39, 4.45, 4.496, 4.52, 4.55, 4.57, 4.60, 4.62, 4.64, 4.66, 4.677, 4.70, 4.71,
                                         4.73, 4.75, 4.76, 4.78, 4.79, 4.809};
        const double tf_us = 1.106;
        double R8 = SONIC_SCHED_LONE8_RATIO, SLOW = SONIC_SCHED_STAGED_SLOWDOWN;
        if (const char* e = getenv("SONIC_SCHED_LONE8_RATIO")) R8 = atof(e);    // (tuning runs)
        if (const char* e = getenv("SONIC_SCHED_STAGED_SLOWDOWN")) SLOW = atof(e);
        double tk[33];
        // (a sparsely populated staged warp next to full-width warps ticks slower than in the calibration runs, where
        // every warp of the device holds the same number of lanes: 2.2-2.5 us measured for k = 1 against 1.65, 2.8-3.2 us
        // for k = 2-3 against 2.25-2.70, no difference at full width.  A factor that decays with k, SDEC ~ 3, describes
        // that, but the schedules it leads to -- more SMs of lone warps -- starve the queue on C2, FHnode and TC:
        // 1.49 / 3.31 / 1.31 s against 1.26 / 2.67 / 1.07 s with the uniform factor, which is the default)
        double SDEC = 1e9;
        if (const char* e = getenv("SONIC_SCHED_SLOWDOWN_DECAY")) SDEC = atof(e);
        for (int k = 1; k <= 32; k++) tk[k] = tk_us[k] / tf_us * (1.0 + (SLOW - 1.0) * exp(-(k - 1) / SDEC));
        double MARGIN = 1.0;
        if (const char* e = getenv("SONIC_SCHED_TIER_MARGIN")) MARGIN = atof(e);
        double QOV = SONIC_SCHED_QUEUE_OVERHEAD;
        if (const char* e = getenv("SONIC_SCHED_QUEUE_OVERHEAD")) QOV = atof(e);
        const double RS = tk[1], T32 = tk[32] * QOV;
        std::vector<double> chain(n), tail(n + 1, 0.0);               // predicted ticks, suffix sums
        for (long long i = 0; i < n; i++) chain[i] = exp(cost[order[i]]);
        for (long long i = n - 1; i >= 0; i--) tail[i] = tail[i + 1] + chain[i];
        // lanes a staged warp may keep busy if its longest chain c has to finish within T
        auto cap_for = [&](double c, double T) {
            int k = 1;
            while (k < 32 && c * tk[k + 1] <= T) k++;
            return k;
        };
        // One decision, the deadline T.  It fixes everything else: chains too long for eight lone warps per SM
        // (c R8 > T) go to tier-1 SMs, which keep one warp per scheduler for them and park their second block;
        // chains too long for a lone lane of a staged warp (c RS > T) go to tier-2 SMs, eight per SM; both kinds of
        // SM host nothing but one-point warps in the register-resident run (one instruction stream in the SM's
        // cache).  The other SMs run staged warps: budgets while the longer chains last, full width on the queue
        // afterwards -- as do the lone warps once their chain is done.  Keep the T with the smallest predicted makespan.
        const bool can_excl = lpw > 1;
        const int wps = warps_per_block * 2;                                     // warps of one SM
        int force1 = -1, force2 = -1, force_cap = 0;
        if (const char* e = getenv("SONIC_SCHED_FORCE_CAP")) force_cap = atoi(e);
        if (const char* e = getenv("SONIC_SCHED_TIER1_SMS")) force1 = atoi(e);
        if (const char* e = getenv("SONIC_SCHED_TIER2_SMS")) force2 = atoi(e);
        auto count_above = [&](double c) {                                      // chains longer than c (sorted descending)
            return (long long)(std::lower_bound(chain.begin(), chain.end(), c, [](double a, double b) { return a > b; }) - chain.begin());
        };
        double best_T = chain[0] * RS, best_span = 1e300;
        int best_E1 = 0, best_E2 = 0;
        for (double T = chain[0] * 1.02; T < chain[0] * 60.0; T *= 1.03) {
            int E1 = 0, E2 = 0;
            if (can_excl) {
                // (the predicted order of the longest chains is not exact: chains down to MARGIN of a tier's
                // threshold are taken into the tier as well)
                const long long n1 = count_above(MARGIN * T / R8), n12 = count_above(MARGIN * T / RS);
                E1 = (int)((n1 + wps / 2 - 1) / (wps / 2));
                if (force1 >= 0) E1 = force1;
                const long long in1 = std::min<long long>((long long)E1 * (wps / 2), n);
                E2 = (int)((std::max(n12 - in1, 0LL) + wps - 1) / wps);
                if (force2 >= 0) E2 = force2;
                if ((E1 + E2) * wps > nwarps * 3 / 4) continue;                  // (a later deadline needs fewer)
            }
            const long long nl1 = std::min<long long>((long long)E1 * (wps / 2), n);
            const long long nl = std::min<long long>(nl1 + (long long)E2 * wps, n);
            if ((nl1 > 0 && chain[0] > T) || (nl > nl1 && chain[nl1] * R8 > T)) continue;
            const int staged_warps = nwarps - (E1 + E2) * wps;
            long long pos = nl;
            int used = 0;
            double have = 0.0;               // warp-time left for the full-width queue before T
            while (used < staged_warps && pos < n) {
                const int k = cap_for(chain[pos], T);
                if (k >= (int)lpw) break;
                // (a budgeted warp goes to full width when its own chains are done)
                have += std::max(0.0, T - chain[pos] * tk[k]);
                pos += k;
                used++;
            }
            // warp-time the full-width queue needs, and what the warps can deliver by T
            const double need = tail[pos] * T32 / 32.0;
            have += (double)(staged_warps - used) * T;
            // (the lone warps of an SM go to full width together, when the SM's longest chain is done)
            for (long long i = 0; i < nl1; i += wps / 2) have += (wps / 2) * std::max(0.0, T - chain[i]);
            for (long long i = nl1; i < nl; i += wps) have += wps * std::max(0.0, T - chain[i] * R8);
            const double queue = have > 0.0 ? T * need / have : (pos < n ? 1e300 : 0.0);
            // (when the budgets run out of warps the queue starts with chains that are too long for a full warp)
            const double head = pos < n ? chain[pos] * tk[32] : 0.0;
            const double span = std::max(std::max(T, queue), head);
            if (span < best_span) { best_span = span; best_T = T; best_E1 = E1; best_E2 = E2; }
            if (queue <= T) break;        // larger T only makes the deadline later
        }
        if (getenv("SONIC_DEBUG"))
            fprintf(stderr, "[sonic] schedule: %d + %d SMs of lone warps, deadline %.3g ticks (longest chain %.3g), predicted span %.3g\n",
                    best_E1, best_E2, best_T, chain[0], best_span);
        long long pos = 0;
        std::vector<int> excl_block(blocks, 0);
        int* bgroup = (int*)(H + o_bgroup);
        int* group0 = (int*)(H + o_group0);
        for (int b = 0; b < (int)blocks; b++) { bgroup[b] = -1; group0[b] = 0; }
        int sm_rank = -1, last_sm = -1, blk_on_sm = 0;
        for (int r = 0; r < nwarps; r++) {
            const int b = border[r / warps_per_block];
            const int gw = b * warps_per_block + r % warps_per_block;
            if (r % warps_per_block == 0) {
                if (smid[b] != last_sm) { last_sm = smid[b]; sm_rank++; blk_on_sm = 0; }
                else blk_on_sm++;
            }
            const bool lone_sm = sm_rank < best_E1 + best_E2;
            if (lone_sm) { excl_block[b] = 1; bgroup[b] = sm_rank; }
            if (sm_rank < best_E1 && blk_on_sm >= 1) {
                wfirst[gw] = (int)n;
                wcap[gw] = 0;                  // parked
                continue;
            }
            if (lone_sm) group0[sm_rank]++;    // every working warp of the group reports once
            if (pos >= n) { if (lone_sm) wcap[gw] = 1; continue; }
            int cap = lone_sm ? 1 : cap_for(chain[pos], best_T);
            if (force_cap > 0) cap = force_cap;                                  // (calibration runs)
            if (cap > (int)lpw) cap = (int)lpw;
            wfirst[gw] = (int)pos;
            wcap[gw] = cap;
            pos += cap;
        }
        p->n_initial = (unsigned long long)(pos < n ? pos : n);
        p->probe_smid = smid;
        // The nested tick is used when no warp of the plan ever holds more than one point (small grids,
        // single points: C1 -12 %).  On a grid that mixes sparse and full warps it is a loss even on SMs
        // that host nothing but one-point warps (C2: 1.36 s instead of 1.26 s, profiles/README.md).
        int* nested = (int*)(H + o_nested);
        int mode = lpw == 1 ? 1 : 2;
        if (const char* e = getenv("SONIC_NESTED")) mode = atoi(e);      // (experiments)
        for (int b = 0; b < (int)blocks; b++) nested[b] = (mode == 2) ? excl_block[b] : mode;
    }
    *(unsigned long long*)(H + o_counter0) = p->n_initial;
    // one upload of everything (or of the budgets alone when the probe already sent the rest)
    CUDA_TRY(cudaMemcpyAsync(D, H, in_bytes, cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->ms_upload = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out_plan = guard.release();
    return SONIC_OK;
}

template <typename T>
static void expand_to(const SonicPlan* p, const T* src, T* out) {
    if (p->umap.empty()) {
        memcpy(out, src, (size_t)p->n * sizeof(T));
        return;
    }
    const int* um = p->umap.data();
    for (long long i = 0; i < p->n_out; i++) out[i] = src[um[i]];
}

extern "C" {

int sonic_version(void) { return SONIC_ABI_VERSION; }

int sonic_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

int sonic_last_error(char* buf, int len) {
    if (buf && len > 0) {
        strncpy(buf, g_err.c_str(), len - 1);
        buf[len - 1] = 0;
    }
    return (int)g_err.size();
}

int sonic_neuron_count(void) { return SONIC_N_NEURONS; }

int sonic_neuron_id(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < SONIC_N_NEURONS; i++)
        if (strcmp(name, SONIC_NEURON_NAMES[i]) == 0) return i;
    return -1;
}

int sonic_neuron_name(int id, char* buf, int len) {
    if (id < 0 || id >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", id);
    if (buf && len > 0) {
        strncpy(buf, SONIC_NEURON_NAMES[id], len - 1);
        buf[len - 1] = 0;
    }
    return SONIC_OK;
}

int sonic_neuron_nrates(int id) {
    if (id < 0 || id >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", id);
    return SONIC_NEURON_NRATES[id];
}

int sonic_neuron_rate_name(int id, int i, char* buf, int len) {
    if (id < 0 || id >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", id);
    if (i < 0 || i >= SONIC_NEURON_NRATES[id]) return set_err(SONIC_E_ARG, "rate index %d out of range", i);
    if (buf && len > 0) {
        strncpy(buf, SONIC_NEURON_RATE_NAMES[id][i], len - 1);
        buf[len - 1] = 0;
    }
    return SONIC_OK;
}

static int rates_common(int device, int id, const double* Vm, int64_t n, double* out, bool mean) {
    int rc = check_device(device);
    if (rc) return rc;
    if (id < 0 || id >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", id);
    if (!Vm || !out || n <= 0) return set_err(SONIC_E_ARG, "invalid Vm/out/n");
    CUDA_TRY(cudaSetDevice(device));
    const int nr = SONIC_NEURON_NRATES[id];
    if (nr == 0) return SONIC_OK;          // passive membrane: no rate constant
    double *d_vm = nullptr, *d_out = nullptr;
    const size_t nout = mean ? (size_t)nr : (size_t)nr * n;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_vm), n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_out), nout * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_vm, Vm, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
#define CALL(ID) launch_rates<ID>(d_vm, n, d_out, mean)
        SONIC_DISPATCH_NEURON(id, CALL)
#undef CALL
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_vm);
    cudaFree(d_out);
    if (e != cudaSuccess) return set_err(SONIC_E_CUDA, "rate evaluation failed: %s", cudaGetErrorString(e));
    return SONIC_OK;
}

int sonic_eval_rates(int device, int id, const double* Vm, int64_t n, double* out) {
    return rates_common(device, id, Vm, n, out, false);
}

int sonic_mean_rates(int device, int id, const double* Vm, int64_t n, double* out) {
    return rates_common(device, id, Vm, n, out, true);
}

int sonic_plan_create(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                      const int32_t* ia, const double* f, const double* A, const double* Q,
                      const double* fs, int nfs, SonicPlan** out_plan) {
    return sonic_plan_create_ex(device, radii, na, neuron_id, n, ia, f, A, Q, 0, nullptr, fs, nfs, out_plan);
}

int sonic_plan_create_ex(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                         const int32_t* ia, const double* f, const double* A, const double* Q,
                         int novertones, const double* overtones, const double* fs, int nfs,
                         SonicPlan** out_plan) {
    const int32_t nid = neuron_id;
    return plan_build(device, radii, nullptr, na, &nid, 1, n, ia, f, A, Q, novertones, overtones, fs, nfs, out_plan);
}

int sonic_plan_create_multi(int device, const SonicBlsParams* radii, const int32_t* radius_neuron, int na,
                            const int32_t* neuron_ids, int n_neurons, int64_t n, const int32_t* ia,
                            const double* f, const double* A, const double* Q, const double* fs, int nfs,
                            SonicPlan** out_plan) {
    if (!radius_neuron) return set_err(SONIC_E_ARG, "radius_neuron is required");
    return plan_build(device, radii, radius_neuron, na, neuron_ids, n_neurons, n, ia, f, A, Q, 0, nullptr, fs, nfs, out_plan);
}

int sonic_plan_launch(SonicPlan* p) {
    if (!p) return set_err(SONIC_E_ARG, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    SonicJob job;
    job.radii = p->d_radii; job.order = p->d_order; job.ia = p->d_ia; job.f = p->d_f; job.A = p->d_A;
    job.Q = p->d_Q; job.ov = p->d_ov; job.nov = p->nov; job.z0 = p->d_z0; job.zbuf = p->d_zbuf; job.ngbuf = p->d_ngbuf;
    job.ncycles = p->d_ncycles; job.status = p->d_status; job.nfe = p->d_nfe; job.nje = p->d_nje;
    job.nsteps = p->d_nsteps; job.tpoint = p->d_tpoint; job.counter = p->d_counter; job.n = p->n;
    job.warp_first = p->d_warp_first; job.warp_cap = p->d_warp_cap;
    job.block_smid = p->d_block_smid; job.probe = 0;
    job.block_nested = p->d_block_nested;
    job.block_group = p->d_block_group; job.group_left = p->d_group_left;
    job.widen = SONIC_WIDEN;
    if (const char* e = getenv("SONIC_WIDEN")) job.widen = atoi(e);      // (tuning runs)
    job.lone_park = SONIC_LONE_PARK;
    if (const char* e = getenv("SONIC_LONE_PARK")) job.lone_park = atoi(e);
    cudaEvent_t* ev = p->ws->ev;
    p->result_on_host = false;
    CUDA_TRY(cudaMemcpyAsync(p->d_counter, p->d_counter0, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_group_left, p->d_group_left0, p->n_groups * sizeof(int), cudaMemcpyDeviceToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_stats, 0, 8 * sizeof(unsigned long long), p->stream));
    CUDA_TRY(cudaEventRecord(ev[0], p->stream));
    sonic_z0_kernel<<<(unsigned)((p->n + 127) / 128), 128, 0, p->stream>>>(job);
    CUDA_TRY(cudaEventRecord(ev[1], p->stream));
    if (p->nov) sonic_integrate_kernel<true><<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(job);
    else sonic_integrate_kernel<false><<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(job);
    CUDA_TRY(cudaEventRecord(ev[2], p->stream));
    for (int k = 0; k < (int)p->neurons.size(); k++) {
        cudaError_t e = cudaSuccess;
#define CALL(ID) e = launch_average<ID>(p, k)
        SONIC_DISPATCH_NEURON(p->neurons[k], CALL)
#undef CALL
        CUDA_TRY(e);
    }
    CUDA_TRY(cudaEventRecord(ev[3], p->stream));
    {
        long long blocks = (p->n + 255) / 256;
        if (blocks > 148) blocks = 148;
        sonic_stats_kernel<<<(int)blocks, 256, 0, p->stream>>>(p->d_nfe, p->d_nje, p->d_nsteps, p->d_ncycles, p->n, p->d_stats);
    }
    CUDA_TRY(cudaEventRecord(ev[4], p->stream));
    CUDA_TRY(cudaGetLastError());
    p->launches += 3 + p->neurons.size();
    p->launched = true;
    return SONIC_OK;
}

int sonic_plan_set_stream(SonicPlan* p, void* stream) {
    if (!p) return set_err(SONIC_E_ARG, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->stream = stream ? static_cast<cudaStream_t>(stream) : p->ws->stream;
    return SONIC_OK;
}

int sonic_plan_sync(SonicPlan* p) {
    if (!p) return set_err(SONIC_E_ARG, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SONIC_OK;
}

// Bring the result range of the last launch into the plan's pinned staging area (one copy).
static int plan_download(SonicPlan* p) {
    if (p->result_on_host) return SONIC_OK;
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaMemcpyAsync(p->ws->hptr + p->in_bytes, p->d_result, p->result_bytes, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->result_on_host = true;
    return SONIC_OK;
}

int sonic_plan_fetch(SonicPlan* p, double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                     double* out_tpoint, uint32_t* out_nrhs) {
    if (!p || !p->launched) return set_err(SONIC_E_ARG, "plan not launched");
    int rc = plan_download(p);
    if (rc) return rc;
    const char* R = p->ws->hptr + p->in_bytes;
    if (out_tables) memcpy(out_tables, R + p->r_out, p->out_count * sizeof(double));
    if (out_ncycles) expand_to(p, (const int32_t*)(R + p->r_ncycles), out_ncycles);
    if (out_status) expand_to(p, (const uint32_t*)(R + p->r_status), out_status);
    if (out_tpoint) expand_to(p, (const double*)(R + p->r_tpoint), out_tpoint);
    if (out_nrhs) expand_to(p, (const uint32_t*)(R + p->r_nfe), out_nrhs);
    return SONIC_OK;
}

int sonic_plan_fetch_zprofiles(SonicPlan* p, double* out_z) {
    if (!p || !p->launched || !out_z) return set_err(SONIC_E_ARG, "plan not launched or null buffer");
    CUDA_TRY(cudaSetDevice(p->device));
    return fetch_rows_expanded(p, p->d_zbuf, out_z);
}

int sonic_plan_fetch_relcm(SonicPlan* p, double* out_cm) {
    if (!p || !p->launched || !out_cm) return set_err(SONIC_E_ARG, "plan not launched or null buffer");
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t total = (size_t)p->n * SONIC_NPC;
    double* d_cm = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_cm), total * sizeof(double)));
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    sonic_relcm_kernel<<<blocks, 256, 0, p->stream>>>(p->d_zbuf, p->d_ia, p->d_radii, p->n, d_cm);
    int rc = SONIC_OK;
    if (cudaGetLastError() != cudaSuccess) rc = set_err(SONIC_E_CUDA, "relative capacitance kernel failed");
    if (!rc) rc = fetch_rows_expanded(p, d_cm, out_cm);
    cudaFree(d_cm);
    if (rc) return rc;
    p->launches += 1;
    return SONIC_OK;
}

int sonic_plan_stats(SonicPlan* p, SonicStats* st) {
    if (!p || !p->launched || !st) return set_err(SONIC_E_ARG, "plan not launched or null stats");
    int rc = plan_download(p);
    if (rc) return rc;
    memset(st, 0, sizeof(*st));
    float ms = 0.f;
    cudaEvent_t* ev = p->ws->ev;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev[0], ev[1])); st->ms_z0 = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev[1], ev[2])); st->ms_integrate = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev[2], ev[3])); st->ms_average = ms;
    const unsigned long long* s = (const unsigned long long*)(p->ws->hptr + p->in_bytes + p->r_stats);
    st->n_rhs = s[0]; st->n_jac = s[1]; st->n_steps = s[2]; st->n_cycles = s[3];
    if (getenv("SONIC_DEBUG")) {
        std::vector<int> now(p->grid);
        CUDA_TRY(cudaMemcpy(now.data(), p->d_block_smid, p->grid * sizeof(int), cudaMemcpyDeviceToHost));
        int diff = 0;
        for (int b = 0; b < p->grid; b++) diff += now[b] != p->probe_smid[b];
        fprintf(stderr, "[sonic] block placement: %d of %d blocks differ from the probe\n", diff, p->grid);
    }
    st->n_points = (uint64_t)p->n_out;
    st->n_launches = p->launches;
    st->ms_total = p->ms_upload;
    return SONIC_OK;
}

int sonic_plan_destroy(SonicPlan* p) { return plan_free(p); }

// one-shot: create, launch, fetch, release (the workspace goes back to the per-device pool)
static int run_plan(SonicPlan* p, std::chrono::steady_clock::time_point t0, double* out_tables, int32_t* out_ncycles,
                    uint32_t* out_status, double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats) {
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    const double ms_create = ms_since(t0);
    int rc = sonic_plan_launch(p);
    if (!rc) rc = sonic_plan_fetch(p, out_tables, out_ncycles, out_status, out_tpoint, out_nrhs);
    const double ms_run = ms_since(t0) - ms_create;
    if (!rc && stats) {
        rc = sonic_plan_stats(p, stats);
        stats->ms_total = ms_since(t0);
    }
    const long long n = p->n_out;
    plan_free(p);
    if (getenv("SONIC_DEBUG"))
        fprintf(stderr, "[sonic] points_run n=%lld: create %.1f ms, launch+fetch %.1f ms, stats+release %.1f ms\n",
                n, ms_create, ms_run, ms_since(t0) - ms_create - ms_run);
    return rc;
}

int sonic_points_run(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                     const int32_t* ia, const double* f, const double* A, const double* Q,
                     const double* fs, int nfs, double* out_tables, int32_t* out_ncycles,
                     uint32_t* out_status, double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats) {
    return sonic_points_run_ex(device, radii, na, neuron_id, n, ia, f, A, Q, 0, nullptr, fs, nfs, out_tables,
                               out_ncycles, out_status, out_tpoint, out_nrhs, stats);
}

int sonic_points_run_ex(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n,
                        const int32_t* ia, const double* f, const double* A, const double* Q,
                        int novertones, const double* overtones, const double* fs, int nfs,
                        double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                        double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    SonicPlan* p = nullptr;
    int rc = sonic_plan_create_ex(device, radii, na, neuron_id, n, ia, f, A, Q, novertones, overtones, fs, nfs, &p);
    if (rc) return rc;
    return run_plan(p, t0, out_tables, out_ncycles, out_status, out_tpoint, out_nrhs, stats);
}

int sonic_points_run_multi(int device, const SonicBlsParams* radii, const int32_t* radius_neuron, int na,
                           const int32_t* neuron_ids, int n_neurons, int64_t n, const int32_t* ia,
                           const double* f, const double* A, const double* Q, const double* fs, int nfs,
                           double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                           double* out_tpoint, uint32_t* out_nrhs, SonicStats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    SonicPlan* p = nullptr;
    int rc = sonic_plan_create_multi(device, radii, radius_neuron, na, neuron_ids, n_neurons, n, ia, f, A, Q, fs, nfs, &p);
    if (rc) return rc;
    return run_plan(p, t0, out_tables, out_ncycles, out_status, out_tpoint, out_nrhs, stats);
}

// Grids of several neurons in one batch: neuron k has its own radii block radii[k * na .. k * na + na)
// and its own charge vector Qcat[Qoff[k] .. Qoff[k + 1]); f, A and fs are shared.  Points are
// flattened neuron > a > f > A > Q; several devices split the trajectories as described below.
static int lookup_common(const SonicBlsParams* radii, int na, const double* f, int nf, const double* A, int nA,
                         const double* Qcat, const int32_t* Qoff, const double* fs, int nfs,
                         const int32_t* neuron_ids, int nn, uint32_t device_mask, double* out_tables,
                         int32_t* out_ncycles, uint32_t* out_status, double* out_tpoint, SonicStats* stats) {
    if (!radii || !f || !A || !Qcat || !Qoff || !fs || na <= 0 || nf <= 0 || nA <= 0 || nfs <= 0 || !out_tables || !neuron_ids || nn <= 0)
        return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    for (int k = 0; k < nn; k++) {
        if (neuron_ids[k] < 0 || neuron_ids[k] >= SONIC_N_NEURONS)
            return set_err(SONIC_E_NEURON, "unknown neuron id %d", neuron_ids[k]);
        if (Qoff[k + 1] <= Qoff[k]) return set_err(SONIC_E_ARG, "empty charge vector for neuron slot %d", k);
    }
    const int ndev_avail = sonic_device_count();
    if (ndev_avail == 0) return check_device(0);
    std::vector<int> devs;
    if (device_mask == 0) device_mask = 1;
    for (int d = 0; d < 32; d++)
        if (device_mask & (1u << d)) {
            if (d >= ndev_avail) return set_err(SONIC_E_NODEVICE, "device %d in mask but only %d present", d, ndev_avail);
            devs.push_back(d);
        }
    const auto t0 = std::chrono::steady_clock::now();
    int64_t n = 0;
    for (int k = 0; k < nn; k++) n += (int64_t)na * nf * nA * (Qoff[k + 1] - Qoff[k]);
    // flatten the grids in the reference's queue order: (neuron >) a > f > A > Q
    std::vector<int32_t> pia(n), rneu((size_t)nn * na);
    std::vector<double> pf(n), pA(n), pQ(n);
    int64_t i = 0;
    for (int k = 0; k < nn; k++) {
        const int nQ = Qoff[k + 1] - Qoff[k];
        const double* Q = Qcat + Qoff[k];
        for (int x = 0; x < na; x++) {
            rneu[(size_t)k * na + x] = k;
            for (int y = 0; y < nf; y++)
                for (int z = 0; z < nA; z++)
                    for (int w = 0; w < nQ; w++, i++) {
                        pia[i] = k * na + x; pf[i] = f[y]; pA[i] = A[z]; pQ[i] = Q[w];
                    }
        }
    }
    const int nd = (int)devs.size();
    if (nd == 1) {
        SonicPlan* p = nullptr;
        int rc = plan_build(devs[0], radii, rneu.data(), nn * na, neuron_ids, nn, n, pia.data(), pf.data(), pA.data(),
                            pQ.data(), 0, nullptr, fs, nfs, &p);
        if (rc) return rc;
        return run_plan(p, t0, out_tables, out_ncycles, out_status, out_tpoint, nullptr, stats);
    }
    // Several devices: the unit of work is the trajectory group (the output points that share one
    // integration: both signs of a charge, neurons with the same sonophore constants), so that no
    // trajectory is integrated twice.  Groups are sorted by predicted cost and dealt round-robin:
    // every device gets the same cost profile.  One host thread per device, host-side scatter.
    std::vector<int> group(n);
    int ngroups = 0;
    {
        std::vector<int> mech(nn * na);
        for (int a1 = 0; a1 < nn * na; a1++) {
            mech[a1] = a1;
            for (int a0 = 0; a0 < a1; a0++)
                if (radii[a0].a == radii[a1].a && radii[a0].Delta == radii[a1].Delta && radii[a0].x0 == radii[a1].x0 &&
                    radii[a0].C == radii[a1].C && radii[a0].nrep == radii[a1].nrep && radii[a0].nattr == radii[a1].nattr &&
                    radii[a0].depth == radii[a1].depth) { mech[a1] = a0; break; }
        }
        struct Key {
            int m; uint64_t f, A, q;
            bool operator==(const Key& o) const { return m == o.m && f == o.f && A == o.A && q == o.q; }
        };
        struct KeyHash {
            size_t operator()(const Key& k) const {
                uint64_t h = 1469598103934665603ULL;
                for (uint64_t v : {(uint64_t)k.m, k.f, k.A, k.q}) { h ^= v; h *= 1099511628211ULL; h ^= h >> 29; }
                return (size_t)h;
            }
        };
        std::unordered_map<Key, int, KeyHash> seen;
        seen.reserve((size_t)n * 2);
        for (int64_t k = 0; k < n; k++) {
            const Key key{mech[pia[k]], dbits(pf[k]), dbits(pA[k]), dbits(qint(pQ[k]))};
            auto it = seen.find(key);
            if (it == seen.end()) it = seen.emplace(key, ngroups++).first;
            group[k] = it->second;
        }
    }
    std::vector<double> gcost(ngroups, 0.0);
    for (int64_t k = 0; k < n; k++) gcost[group[k]] = predict_log_cost(radii[pia[k]].a, pf[k], pA[k], pQ[k]);
    std::vector<int> gorder(ngroups);
    for (int g = 0; g < ngroups; g++) gorder[g] = g;
    std::stable_sort(gorder.begin(), gorder.end(), [&](int x, int y) { return gcost[x] > gcost[y]; });
    std::vector<int> owner(ngroups);
    for (int r = 0; r < ngroups; r++) owner[gorder[r]] = r % nd;
    struct Shard {
        std::vector<int> idx;
        std::vector<int32_t> ia, ncyc;
        std::vector<double> f, A, Q, out, tp;
        std::vector<uint32_t> st;
        std::vector<int64_t> nk;      // output points per neuron in this shard
        SonicStats stats;
        int rc = 0;
        std::string err;
    };
    std::vector<Shard> sh(nd);
    for (int64_t k = 0; k < n; k++) sh[owner[group[k]]].idx.push_back((int)k);
    std::vector<int> nvar(nn);
    for (int k = 0; k < nn; k++) nvar[k] = 1 + SONIC_NEURON_NRATES[neuron_ids[k]];
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        Shard& s = sh[d];
        const size_t m = s.idx.size();
        s.ia.resize(m); s.f.resize(m); s.A.resize(m); s.Q.resize(m);
        s.ncyc.resize(m); s.st.resize(m); s.tp.resize(m);
        s.nk.assign(nn, 0);
        size_t outc = 0;
        for (size_t k = 0; k < m; k++) {
            const int g = s.idx[k];
            s.ia[k] = pia[g]; s.f[k] = pf[g]; s.A[k] = pA[g]; s.Q[k] = pQ[g];
            s.nk[rneu[pia[g]]]++;
        }
        for (int k = 0; k < nn; k++) outc += (size_t)nvar[k] * s.nk[k] * nfs;
        s.out.resize(outc);
        th.emplace_back([&, d]() {
            Shard& s2 = sh[d];
            if (s2.idx.empty()) return;
            s2.rc = sonic_points_run_multi(devs[d], radii, rneu.data(), nn * na, neuron_ids, nn, (int64_t)s2.idx.size(),
                                           s2.ia.data(), s2.f.data(), s2.A.data(), s2.Q.data(), fs, nfs, s2.out.data(),
                                           s2.ncyc.data(), s2.st.data(), s2.tp.data(), nullptr, &s2.stats);
            if (s2.rc) s2.err = g_err;
        });
    }
    for (auto& t : th) t.join();
    // global table block of neuron k: [nvar_k][n_k][nfs], its points in flattened order
    std::vector<int64_t> nk_all(nn), first_pt(nn), out_off(nn);
    {
        int64_t pt = 0, off = 0;
        for (int k = 0; k < nn; k++) {
            nk_all[k] = (int64_t)na * nf * nA * (Qoff[k + 1] - Qoff[k]);
            first_pt[k] = pt; out_off[k] = off;
            pt += nk_all[k]; off += (int64_t)nvar[k] * nk_all[k] * nfs;
        }
    }
    SonicStats tot;
    memset(&tot, 0, sizeof(tot));
    for (int d = 0; d < nd; d++) {
        Shard& s = sh[d];
        if (s.rc) return set_err(s.rc, "device %d: %s", devs[d], s.err.c_str());
        const size_t m = s.idx.size();
        std::vector<int64_t> seen_k(nn, 0), soff(nn, 0);
        {
            int64_t off = 0;
            for (int k = 0; k < nn; k++) { soff[k] = off; off += (int64_t)nvar[k] * s.nk[k] * nfs; }
        }
        for (size_t k = 0; k < m; k++) {
            const int64_t g = s.idx[k];
            const int nk = rneu[pia[g]];
            const int64_t local = seen_k[nk]++;          // position among this shard's points of the neuron
            const int64_t gl = g - first_pt[nk];         // position among all points of the neuron
            for (int v = 0; v < nvar[nk]; v++)
                for (int j = 0; j < nfs; j++)
                    out_tables[out_off[nk] + ((int64_t)v * nk_all[nk] + gl) * nfs + j] =
                        s.out[soff[nk] + ((int64_t)v * s.nk[nk] + local) * nfs + j];
            if (out_ncycles) out_ncycles[g] = s.ncyc[k];
            if (out_status) out_status[g] = s.st[k];
            if (out_tpoint) out_tpoint[g] = s.tp[k];
        }
        if (m) {
            tot.n_points += s.stats.n_points; tot.n_rhs += s.stats.n_rhs; tot.n_jac += s.stats.n_jac;
            tot.n_steps += s.stats.n_steps; tot.n_cycles += s.stats.n_cycles; tot.n_launches += s.stats.n_launches;
            tot.ms_z0 = std::max(tot.ms_z0, s.stats.ms_z0);
            tot.ms_integrate = std::max(tot.ms_integrate, s.stats.ms_integrate);
            tot.ms_average = std::max(tot.ms_average, s.stats.ms_average);
        }
    }
    tot.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = tot;
    return SONIC_OK;
}

int sonic_lookup_run(const SonicBlsParams* radii, int na, const double* f, int nf, const double* A,
                     int nA, const double* Q, int nQ, const double* fs, int nfs, int neuron_id,
                     uint32_t device_mask, double* out_tables, int32_t* out_ncycles,
                     uint32_t* out_status, double* out_tpoint, SonicStats* stats) {
    if (nQ <= 0) return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    const int32_t qoff[2] = {0, nQ};
    const int32_t nid = neuron_id;
    return lookup_common(radii, na, f, nf, A, nA, Q, qoff, fs, nfs, &nid, 1, device_mask, out_tables, out_ncycles,
                         out_status, out_tpoint, stats);
}

int sonic_lookup_run_multi(const SonicBlsParams* radii, int na, const double* f, int nf, const double* A, int nA,
                           const double* Qcat, const int32_t* Qoff, const double* fs, int nfs,
                           const int32_t* neuron_ids, int n_neurons, uint32_t device_mask, double* out_tables,
                           int32_t* out_ncycles, uint32_t* out_status, double* out_tpoint, SonicStats* stats) {
    return lookup_common(radii, na, f, nf, A, nA, Qcat, Qoff, fs, nfs, neuron_ids, n_neurons, device_mask, out_tables,
                         out_ncycles, out_status, out_tpoint, stats);
}

int sonic_pmavg(int device, double a, double Delta, int64_t n, const double* Z, double* out_pm, int32_t* out_last) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!(a > 0.) || !(Delta > 0.) || n <= 0 || !Z || !out_pm)
        return set_err(SONIC_E_ARG, "invalid argument (a, Delta > 0; n > 0; Z and out_pm required)");
    CUDA_TRY(cudaSetDevice(device));
    double* d_Z = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_Z), (size_t)n * (2 * sizeof(double) + sizeof(int)));
    double* d_out = d_Z + n;
    int* d_last = reinterpret_cast<int*>(d_out + n);
    if (e == cudaSuccess) e = cudaMemcpy(d_Z, Z, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        sonic_pmavg_kernel<<<(unsigned)((n + 63) / 64), 64>>>(a, Delta, n, d_Z, d_out, d_last);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_pm, d_out, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_last) e = cudaMemcpy(out_last, d_last, n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_Z);
    if (e != cudaSuccess) return set_err(SONIC_E_CUDA, "intermolecular pressure quadrature failed: %s", cudaGetErrorString(e));
    return SONIC_OK;
}

int sonic_sim_nstates(int neuron_id) {
    if (neuron_id < 0 || neuron_id >= SONIC_N_NEURONS) return set_err(SONIC_E_NEURON, "unknown neuron id %d", neuron_id);
    int ns = 0;
#define CALL(ID) ns = sim_nstates<ID>()
    SONIC_DISPATCH_NEURON(neuron_id, CALL)
#undef CALL
    return ns;
}

int sonic_simulate(int device, int neuron_id, int nsim, int nQ, const double* Qref, const double* tab_on,
                   const double* tab_off, int nt, const double* t, const uint8_t* stim_on, const double* y0, int nsub,
                   double* out, int32_t* status) {
    int rc = check_device(device);
    if (rc) return rc;
    const int ns = sonic_sim_nstates(neuron_id);
    if (ns < 0) return ns;
    if (ns == 0) return set_err(SONIC_E_NEURON, "the SONIC simulation of neuron %d is not supported (states that are not gates)", neuron_id);
    if (nsim <= 0 || nQ < 2 || nt < 2 || nsub < 1 || !Qref || !tab_on || !tab_off || !t || !stim_on || !y0 || !out || !status)
        return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    for (int i = 1; i < nQ; i++)
        if (!(Qref[i] > Qref[i - 1])) return set_err(SONIC_E_ARG, "charge vector must be strictly ascending");
    for (int i = 1; i < nt; i++)
        if (!(t[i] >= t[i - 1])) return set_err(SONIC_E_ARG, "sample times must not decrease");
    CUDA_TRY(cudaSetDevice(device));
    const int nv = 1 + 2 * ns;
    const size_t b_Q = (size_t)nQ * 8, b_on = (size_t)nsim * nv * nQ * 8, b_off = (size_t)nv * nQ * 8, b_t = (size_t)nt * 8;
    const size_t b_y0 = (size_t)(1 + ns) * 8, b_out = (size_t)nsim * nt * (1 + ns) * 8, b_st = (size_t)nsim * 4;
    const size_t b_s = ((size_t)nt + 7) & ~(size_t)7;
    char* d = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d), b_Q + b_on + b_off + b_t + b_y0 + b_out + b_st + b_s);
    SonicSimArgs a;
    if (e == cudaSuccess) {
        char* q = d;
        a.Qref = (double*)q; q += b_Q;
        a.tab_on = (double*)q; q += b_on;
        a.tab_off = (double*)q; q += b_off;
        a.t = (double*)q; q += b_t;
        a.y0 = (double*)q; q += b_y0;
        a.out = (double*)q; q += b_out;
        a.status = (int*)q; q += b_st;
        a.on = (unsigned char*)q;
        a.nsim = nsim; a.nQ = nQ; a.nt = nt; a.nsub = nsub;
        e = cudaMemcpy((void*)a.Qref, Qref, b_Q, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemcpy((void*)a.tab_on, tab_on, b_on, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)a.tab_off, tab_off, b_off, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)a.t, t, b_t, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)a.y0, y0, b_y0, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((void*)a.on, stim_on, nt, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
#define CALL(ID) e = launch_simulate<ID>(a)
        SONIC_DISPATCH_NEURON(neuron_id, CALL)
#undef CALL
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, a.out, b_out, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(status, a.status, b_st, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(SONIC_E_CUDA, "SONIC simulation failed: %s", cudaGetErrorString(e));
    return SONIC_OK;
}

int sonic_trim(void) {
    std::vector<Workspace*> idle;
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        for (auto& pool : g_pool) {
            idle.insert(idle.end(), pool.begin(), pool.end());
            pool.clear();
        }
    }
    for (Workspace* w : idle) workspace_free(w);
    return SONIC_OK;
}

int sonic_fp64_peak(int device, double* tflops) {
    int rc = check_device(device);
    if (rc) return rc;
    if (!tflops) return set_err(SONIC_E_ARG, "null output");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_out), (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        sonic_dfma_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return SONIC_OK;
}

}  // extern "C"
