// sonic_b200.cu -- CUDA kernels (sm_100a) and C ABI of the SONIC lookup-table engine.
//
// Kernels
//   sonic_z0_kernel         initial quasi-static deflection of every point (one thread/point)
//   sonic_integrate_kernel  persistent batched integrator: one lane per grid point, lanes
//                           refill themselves from a cost-sorted work queue; every tick = one
//                           warp-convergent RHS evaluation + per-lane controller bookkeeping;
//                           periodic-convergence test on the fly against the previous cycle
//   sonic_average_kernel    fused cycle averaging: Z(t) -> Cm -> per coverage fraction V(t)
//                           -> generated neuron rate functions -> warp-shuffle means
//   sonic_rates_kernel      elementwise / mean evaluation of the generated rate functions
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

#include "../../include/sonic_b200.h"
#include "generated/cost_table.h"
#include "generated/neuron_rates.cuh"
#include "sonic_core.h"

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
    int cap = job.warp_cap[gwarp];
    long long q_init = (lane < cap) ? (long long)job.warp_first[gwarp] + lane : -1;
    bool exhausted = false;

    while (true) {
        if (pt < 0 && !exhausted && lane < cap) {
            // (re)fill this lane: its initial point first, then the shared work queue
            unsigned long long q;
            if (q_init >= 0) {
                q = (unsigned long long)q_init;
                q_init = -1;
            } else {
                q = atomicAdd(job.counter, 1ULL);
            }
            if (q < (unsigned long long)job.n) {
                pt = job.order[q];
                const double f = job.f[pt];
                sonic_point_init(p, job.radii[job.ia[pt]], f, job.A[pt], job.Q[pt], job.nov,
                                 job.nov ? job.ov + (size_t)pt * 2 * job.nov : nullptr);
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
        do {
            if (active) {
                double fv[3];
                if (p.nov) sonic_update_charge(p, s.tn);
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
    }
}

// Fused cycle averaging.  One warp per ODE point; the point's capacitance profile is staged in
// shared memory and reused for every coverage fraction.
template <int NID>
__global__ void __launch_bounds__(32 * SONIC_AVG_WARPS)
sonic_average_kernel(const double* __restrict__ zbuf, const int* __restrict__ ia,
                     const double* __restrict__ Q, const SonicBls* __restrict__ radii,
                     const unsigned* __restrict__ status, long long n,
                     const double* __restrict__ fs, int nfs, int nov, const double* __restrict__ ov,
                     const int* __restrict__ umap, double* __restrict__ out) {
    constexpr int NR = SonicRates<NID>::N;
    constexpr int NV = 1 + 2 * SONIC_MAX_OVERTONES + NR;
    __shared__ double cm_s[SONIC_AVG_WARPS][SONIC_NPC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * SONIC_AVG_WARPS;
    const int nvar = 1 + 2 * nov + NR;          // V, (A_Vk, phi_Vk) per overtone, rates
    // n = output points, Q = their signed charges; u = trajectory the point reads
    for (long long pt = (long long)blockIdx.x * SONIC_AVG_WARPS + warp; pt < n; pt += nwarps) {
        const long long u = umap ? umap[pt] : pt;
        const SonicBls b = radii[ia[u]];
        const double a2 = b.a * b.a;
        const double q0 = Q[pt];
        const double* ovp = nov ? ov + (size_t)u * 2 * nov : nullptr;
        const bool bad = (status[u] & (SONIC_ST_Z0FAIL | SONIC_ST_STEPFAIL | SONIC_ST_MXSTEP |
                                       SONIC_ST_TOLSF)) != 0;
        const double* z = zbuf + u * SONIC_NPC;
        double* cm = cm_s[warp];
        for (int k = lane; k < SONIC_NPC; k += 32) cm[k] = sonic_capacitance(a2, b.Delta, b.Cm0, z[k]);
        __syncwarp();
        for (int j = 0; j < nfs; j++) {
            const double x = fs[j];
            double acc[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] = 0.0;
            for (int k = lane; k < SONIC_NPC; k += 32) {
                // imposed charge of sample k (constant without overtones, nbls.py:169-178)
                const double q = nov ? sonic_charge_sample(q0, nov, ovp, k) : q0;
                // spatial average of the capacitance, then membrane potential in mV
                const double vm = q / (x * cm[k] + (1 - x) * b.Cm0) * 1e3;   // nbls.py:148-151,188
                double r[NR];
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
                for (int v = 0; v < NR; v++) acc[1 + 2 * SONIC_MAX_OVERTONES + v] += r[v];
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
            // lane v stores table v
            double mine = 0.0;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int tv = v <= 2 * nov ? v : v - 2 * SONIC_MAX_OVERTONES + 2 * nov;   // table of slot v
                const bool used = v <= 2 * nov || v > 2 * SONIC_MAX_OVERTONES;
                if (used && lane == tv) mine = acc[v];
            }
            if (lane < nvar) out[((long long)lane * n + pt) * nfs + j] = bad ? nan("") : mine;
        }
        __syncwarp();
    }
}

template <int NID>
__global__ void sonic_rates_kernel(const double* __restrict__ vm, long long n, double* __restrict__ out) {
    constexpr int NR = SonicRates<NID>::N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        double r[NR];
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
    __shared__ double part[8][NR];
    double acc[NR];
#pragma unroll
    for (int v = 0; v < NR; v++) acc[v] = 0.0;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) {
        double r[NR];
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
static SonicTables g_host_tables;
static bool g_tables_ready = false;

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

// Predicted cost (log of right-hand-side evaluations) of a point, used only to order the work
// queue (longest first) and to size the lane budgets: nearest node of a table measured on the
// RS 4-D grid (generated/cost_table.h; the charge enters the mechanics through Q^2 only).
static int nearest_log(const double* nodes, int n, double x) {
    // nearest node on a logarithmic axis = first node whose geometric midpoint with the next
    // one lies above x (nodes ascending)
    int i = 0;
    while (i + 1 < n && x * x > nodes[i] * nodes[i + 1]) i++;
    return i;
}

static int bin_of(const double* edges, int nbins, double x) {
    int j = 0;
    while (j + 1 < nbins && x >= edges[j + 1]) j++;
    return j;
}

static double predict_log_cost(double a, double f, double A, double Q) {
    const int i = nearest_log(SONIC_COST_A, SONIC_COST_NA, a);
    const int j = nearest_log(SONIC_COST_F, SONIC_COST_NF, f);
    const int k = bin_of(SONIC_COST_AMP_EDGES, SONIC_COST_NAMP, A);
    const int l = bin_of(SONIC_COST_Q_EDGES, SONIC_COST_NQ, fabs(Q) + 1e-14);
    return SONIC_COST_LOG[((i * SONIC_COST_NF + j) * SONIC_COST_NAMP + k) * SONIC_COST_NQ + l];
}

// ---------------------------------------------------------------------------------------
// Workspace pool: the two large per-plan buffers (cycle profiles) are kept per device between
// calls, so that repeated one-shot calls (sonic_points_run / sonic_lookup_run) do not pay a
// multi-GB cudaMalloc / cudaFree every time.  sonic_trim() releases them.
// ---------------------------------------------------------------------------------------
struct PoolSlot {
    double* ptr = nullptr;
    size_t count = 0;
};
static std::mutex g_pool_mutex;
static PoolSlot g_pool[64][2];   // [device][0 = zbuf, 1 = ngbuf]

static cudaError_t pool_take(int device, int which, size_t count, double** out) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        PoolSlot& sl = g_pool[device & 63][which];
        if (sl.ptr && sl.count >= count) {
            *out = sl.ptr;
            sl.ptr = nullptr;
            sl.count = 0;
            return cudaSuccess;
        }
    }
    return cudaMalloc(reinterpret_cast<void**>(out), std::max<size_t>(count, 1) * sizeof(double));
}

static void pool_give(int device, int which, double* ptr, size_t count) {
    if (!ptr) return;
    double* drop = ptr;
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        PoolSlot& sl = g_pool[device & 63][which];
        if (!sl.ptr || sl.count < count) {
            drop = sl.ptr;
            sl.ptr = ptr;
            sl.count = count;
        }
    }
    if (drop) cudaFree(drop);
}

struct SonicPlan {
    int device = 0;
    int neuron_id = 0;
    int nrates = 0;
    int na = 0, nfs = 0;
    int nov = 0;               // charge overtones per point
    int nvar = 0;              // tables per point: 1 + 2 nov + nrates
    double* d_ov = nullptr;
    long long n = 0;           // points integrated (unique trajectories)
    long long n_out = 0;       // points of the output tables (>= n: +Q / -Q pairs share a trajectory)
    int* d_umap = nullptr;     // [n_out] -> trajectory index, or null (identity)
    double* d_Qout = nullptr;  // [n_out] signed charges of the output points
    std::vector<int> umap;
    long long slots = 0;
    int grid = 0, lanes_per_warp = 32;
    size_t zbuf_count = 0, ngbuf_count = 0;
    unsigned long long n_initial = 0;   // work-queue positions handed out statically (counter start)
    int *d_warp_first = nullptr, *d_warp_cap = nullptr, *d_block_smid = nullptr;
    std::vector<int> probe_smid, last_smid;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // device buffers
    SonicBls* d_radii = nullptr;
    int *d_order = nullptr, *d_ia = nullptr, *d_ncycles = nullptr;
    double *d_f = nullptr, *d_A = nullptr, *d_Q = nullptr, *d_fs = nullptr, *d_z0 = nullptr;
    double *d_zbuf = nullptr, *d_ngbuf = nullptr, *d_tpoint = nullptr, *d_out = nullptr;
    unsigned *d_status = nullptr, *d_nfe = nullptr, *d_nje = nullptr, *d_nsteps = nullptr;
    unsigned long long* d_counter = nullptr;
    uint64_t launches = 0;
    double ms_upload = 0.0;
    bool launched = false;
};

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
}

static int plan_free(SonicPlan* p) {
    if (!p) return SONIC_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_radii); cudaFree(p->d_order); cudaFree(p->d_ia); cudaFree(p->d_ncycles);
    cudaFree(p->d_f); cudaFree(p->d_A); cudaFree(p->d_Q); cudaFree(p->d_fs); cudaFree(p->d_z0);
    pool_give(p->device, 0, p->d_zbuf, p->zbuf_count); pool_give(p->device, 1, p->d_ngbuf, p->ngbuf_count);
    cudaFree(p->d_tpoint); cudaFree(p->d_out); cudaFree(p->d_ov); cudaFree(p->d_umap); cudaFree(p->d_Qout);
    cudaFree(p->d_status); cudaFree(p->d_nfe); cudaFree(p->d_nje); cudaFree(p->d_nsteps);
    cudaFree(p->d_counter); cudaFree(p->d_warp_first); cudaFree(p->d_warp_cap); cudaFree(p->d_block_smid);
    for (auto& e : p->ev)
        if (e) cudaEventDestroy(e);
    if (p->stream && p->own_stream) cudaStreamDestroy(p->stream);
    delete p;
    return SONIC_OK;
}

template <int NID>
static void launch_average(SonicPlan* p, int blocks) {
    sonic_average_kernel<NID><<<blocks, 32 * SONIC_AVG_WARPS, 0, p->stream>>>(
        p->d_zbuf, p->d_ia, p->d_Qout, p->d_radii, p->d_status, p->n_out, p->d_fs, p->nfs, p->nov, p->d_ov, p->d_umap,
        p->d_out);
}

template <int NID>
static void launch_rates(const double* vm, long long n, double* out, bool mean) {
    if (mean)
        sonic_mean_rates_kernel<NID><<<1, 256>>>(vm, n, out);
    else
        sonic_rates_kernel<NID><<<(int)std::min<long long>((n + 255) / 256, 4096), 256>>>(vm, n, out);
}

// Per-trajectory array on the device -> per-output-point array on the host.
template <typename T>
static int fetch_expanded(SonicPlan* p, const T* d_src, T* out) {
    if (p->umap.empty()) {
        CUDA_TRY(cudaMemcpyAsync(out, d_src, (size_t)p->n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        return SONIC_OK;
    }
    std::vector<T> tmp((size_t)p->n);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), d_src, (size_t)p->n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (long long i = 0; i < p->n_out; i++) out[i] = tmp[p->umap[i]];
    return SONIC_OK;
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
    double *d_vm = nullptr, *d_out = nullptr;
    const size_t nout = mean ? (size_t)nr : (size_t)nr * n;
    CUDA_TRY(dalloc(&d_vm, n));
    CUDA_TRY(dalloc(&d_out, nout));
    CUDA_TRY(cudaMemcpy(d_vm, Vm, n * sizeof(double), cudaMemcpyHostToDevice));
#define CALL(ID) launch_rates<ID>(d_vm, n, d_out, mean)
    SONIC_DISPATCH_NEURON(id, CALL)
#undef CALL
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, d_out, nout * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_vm);
    cudaFree(d_out);
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

int sonic_plan_create_ex(int device, const SonicBlsParams* radii, int na, int neuron_id, int64_t n_in,
                         const int32_t* ia_in, const double* f_in, const double* A_in, const double* Q_in,
                         int novertones, const double* overtones, const double* fs, int nfs,
                         SonicPlan** out_plan) {
    int64_t n = n_in;
    const int32_t* ia = ia_in;
    const double *f = f_in, *A = A_in, *Q = Q_in;
    int rc = check_device(device);
    if (rc) return rc;
    if (neuron_id < 0 || neuron_id >= SONIC_N_NEURONS)
        return set_err(SONIC_E_NEURON, "unknown neuron id %d", neuron_id);
    if (!radii || na <= 0 || n <= 0 || !ia || !f || !A || !Q || !fs || nfs <= 0 || !out_plan)
        return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    if (n > 0x7fffffffLL) return set_err(SONIC_E_ARG, "too many points for one plan (%lld)", (long long)n);
    if (novertones < 0 || novertones > SONIC_MAX_OVERTONES || (novertones > 0 && !overtones))
        return set_err(SONIC_E_ARG, "invalid charge overtones (0..%d per point, array required)", SONIC_MAX_OVERTONES);
    for (int64_t i = 0; i < n; i++) {
        if (ia[i] < 0 || ia[i] >= na) return set_err(SONIC_E_ARG, "radius index out of range at point %lld", (long long)i);
        if (!(f[i] > 0.)) return set_err(SONIC_E_ARG, "frequency must be strictly positive (point %lld)", (long long)i);
        if (!(A[i] >= 0.)) return set_err(SONIC_E_ARG, "amplitude must be positive or null (point %lld)", (long long)i);
    }
    for (int i = 0; i < na; i++)
        if (!(radii[i].a > 0.) || !(radii[i].Delta > 0.) || !(radii[i].Cm0 > 0.))
            return set_err(SONIC_E_ARG, "invalid sonophore constants for radius %d", i);
    CUDA_TRY(cudaSetDevice(device));
    if (!g_tables_ready) {
        sonic_fill_tables(&g_host_tables);
        g_tables_ready = true;
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_tables, &g_host_tables, sizeof(SonicTables)));

    const auto t0 = std::chrono::steady_clock::now();
    // The mechanics see the charge through Q^2 only (bls.py:482-491): points that differ by the
    // sign of Q share one trajectory bit for bit (the reference integrates both).  Integrate each
    // (radius, f, A, |Q|) once; the averaging kernel then applies each point's own signed charge.
    std::vector<int32_t> u_ia;
    std::vector<double> u_f, u_A, u_Q;
    std::vector<int> umap;
    // |Q| as the integrator sees it: rounded to 36 significant bits (1.5e-11 relative, four orders
    // below the integrator's tolerance), so that grid values such as the -3 and +3 nC/cm2 of an
    // np.arange, which differ in their last bits, select the same trajectory -- and so that a point
    // gets the same trajectory whether it is computed alone, inside a grid, or with zero-amplitude
    // charge overtones.  The averaging kernel uses the exact charges.
    auto qint = [](double q) {
        uint64_t u;
        q = fabs(q);
        memcpy(&u, &q, 8);
        u = (u + 0x8000ULL) & ~0xFFFFULL;
        memcpy(&q, &u, 8);
        return q;
    };
    if (novertones > 0) {
        u_Q.resize(n_in);
        for (int64_t i = 0; i < n_in; i++) u_Q[i] = copysign(qint(Q_in[i]), Q_in[i]);
        Q = u_Q.data();
    } else {
        struct Key {
            int32_t ia; uint64_t f, A, q;
            bool operator==(const Key& o) const { return ia == o.ia && f == o.f && A == o.A && q == o.q; }
        };
        struct KeyHash {
            size_t operator()(const Key& k) const {
                uint64_t h = 1469598103934665603ULL;
                for (uint64_t v : {(uint64_t)k.ia, k.f, k.A, k.q}) { h ^= v; h *= 1099511628211ULL; h ^= h >> 29; }
                return (size_t)h;
            }
        };
        auto bits = [](double x) { uint64_t u; memcpy(&u, &x, 8); return u; };
        std::unordered_map<Key, int, KeyHash> seen;
        seen.reserve((size_t)n_in * 2);
        umap.resize(n_in);
        for (int64_t i = 0; i < n_in; i++) {
            const double qi = qint(Q_in[i]);
            const Key k{ia_in[i], bits(f_in[i]), bits(A_in[i]), bits(qi)};
            auto it = seen.find(k);
            if (it == seen.end()) {
                it = seen.emplace(k, (int)u_ia.size()).first;
                u_ia.push_back(ia_in[i]); u_f.push_back(f_in[i]); u_A.push_back(A_in[i]); u_Q.push_back(qi);
            }
            umap[i] = it->second;
        }
        n = (int64_t)u_ia.size();
        ia = u_ia.data(); f = u_f.data(); A = u_A.data(); Q = u_Q.data();
        if (n == n_in) umap.clear();      // nothing to share: identity (the order is unchanged)
    }
    SonicPlan* p = new SonicPlan();
    p->n_out = n_in;
    p->device = device;
    p->neuron_id = neuron_id;
    p->nrates = SONIC_NEURON_NRATES[neuron_id];
    p->na = na;
    p->nfs = nfs;
    p->n = n;
    p->nov = novertones;
    p->nvar = 1 + 2 * novertones + p->nrates;

    // launch geometry of the persistent integrator
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    int blocks_per_sm = 0;
    CUDA_TRY(cudaFuncSetAttribute(sonic_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SONIC_HIST_BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sonic_integrate_kernel,
                                                           SONIC_BLOCK, SONIC_HIST_BYTES));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const long long max_blocks = (long long)prop.multiProcessorCount * blocks_per_sm;
    const long long warps_total = max_blocks * (SONIC_BLOCK / 32);
    // few points: spread them over as many warps as possible (the chain of one point is
    // serial, so idle lanes cost nothing while extra warps shorten the critical path)
    long long lpw = (n + warps_total - 1) / warps_total;
    if (lpw > 32) lpw = 32;
    if (lpw < 1) lpw = 1;
    p->lanes_per_warp = (int)lpw;
    long long need_warps = (n + lpw - 1) / lpw;
    long long blocks = (need_warps + (SONIC_BLOCK / 32) - 1) / (SONIC_BLOCK / 32);
    if (blocks > max_blocks) blocks = max_blocks;
    p->grid = (int)blocks;
    p->slots = blocks * SONIC_BLOCK;

    // work-queue order: predicted cost, longest first
    std::vector<int> order(n);
    std::vector<double> cost(n);
    for (int64_t i = 0; i < n; i++) {
        order[i] = (int)i;
        cost[i] = predict_log_cost(radii[ia[i]].a, f[i], A[i], Q[i]);
        if (novertones > 0) {
            // a sample-and-hold charge restarts the integrator a thousand times per cycle: about
            // 8e4 right-hand sides per cycle whatever the drive, so the chain length is set by the
            // number of cycles (11 in the noise regime A < 8 kPa, 3 otherwise; measured on RS)
            cost[i] = log(8e4 * (A[i] < 8e3 ? 11.0 : 3.0)) + 0.01 * cost[i];
        }
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });

    std::vector<SonicBls> hb(na);
    for (int i = 0; i < na; i++) {
        hb[i].a = radii[i].a; hb[i].Delta = radii[i].Delta; hb[i].x0 = radii[i].x0; hb[i].C = radii[i].C;
        hb[i].nrep = radii[i].nrep; hb[i].nattr = radii[i].nattr; hb[i].Cm0 = radii[i].Cm0;
        hb[i].depth = radii[i].depth;
    }
    cudaError_t e = cudaSuccess;
#define TRYA(x) if (e == cudaSuccess) e = (x)
    TRYA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (auto& ev : p->ev) TRYA(cudaEventCreate(&ev));
    TRYA(dalloc(&p->d_radii, na));
    TRYA(dalloc(&p->d_order, n)); TRYA(dalloc(&p->d_ia, n)); TRYA(dalloc(&p->d_ncycles, n));
    TRYA(dalloc(&p->d_f, n)); TRYA(dalloc(&p->d_A, n)); TRYA(dalloc(&p->d_Q, n));
    TRYA(dalloc(&p->d_fs, nfs)); TRYA(dalloc(&p->d_z0, n));
    p->zbuf_count = (size_t)n * SONIC_NPC;
    p->ngbuf_count = (size_t)p->slots * SONIC_NPC;
    TRYA(pool_take(device, 0, p->zbuf_count, &p->d_zbuf));
    TRYA(pool_take(device, 1, p->ngbuf_count, &p->d_ngbuf));
    TRYA(dalloc(&p->d_tpoint, n));
    TRYA(dalloc(&p->d_out, (size_t)p->nvar * n_in * nfs));
    TRYA(dalloc(&p->d_Qout, n_in));
    TRYA(cudaMemcpyAsync(p->d_Qout, Q_in, n_in * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (!umap.empty()) {
        TRYA(dalloc(&p->d_umap, n_in));
        TRYA(cudaMemcpyAsync(p->d_umap, umap.data(), n_in * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        p->umap = umap;
    }
    TRYA(dalloc(&p->d_ov, (size_t)n * 2 * std::max(novertones, 1)));
    if (novertones > 0)
        TRYA(cudaMemcpyAsync(p->d_ov, overtones, (size_t)n * 2 * novertones * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    TRYA(dalloc(&p->d_status, n)); TRYA(dalloc(&p->d_nfe, n)); TRYA(dalloc(&p->d_nje, n));
    TRYA(dalloc(&p->d_nsteps, n)); TRYA(dalloc(&p->d_counter, 1));
    const int warps_per_block = SONIC_BLOCK / 32;
    const int nwarps = (int)blocks * warps_per_block;
    TRYA(dalloc(&p->d_warp_first, nwarps)); TRYA(dalloc(&p->d_warp_cap, nwarps));
    TRYA(dalloc(&p->d_block_smid, blocks));
    TRYA(cudaMemcpyAsync(p->d_radii, hb.data(), na * sizeof(SonicBls), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_order, order.data(), n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_ia, ia, n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_f, f, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_A, A, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_Q, Q, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_fs, fs, nfs * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    // Placement probe: the same kernel, same launch configuration, returns after recording the
    // SM of every block.  The persistent grid is exactly one resident wave, so the real launches
    // land the same way; if they ever do not, only the schedule quality suffers, never the
    // results (every point is still owned by exactly one warp).
    std::vector<int> smid(blocks, 0);
    if (e == cudaSuccess) {
        SonicJob probe;
        memset(&probe, 0, sizeof(probe));
        probe.block_smid = p->d_block_smid;
        probe.counter = p->d_counter;
        probe.probe = 1;
        TRYA(cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned long long), p->stream));
        // same launch sequence as sonic_plan_launch (the initial-deflection kernel right before):
        // where its last blocks retire decides which SMs take the first integrator blocks
        {
            SonicJob zj;
            memset(&zj, 0, sizeof(zj));
            zj.radii = p->d_radii; zj.ia = p->d_ia; zj.f = p->d_f; zj.A = p->d_A; zj.Q = p->d_Q;
            zj.z0 = p->d_z0; zj.n = p->n; zj.ov = p->d_ov; zj.nov = p->nov;
            sonic_z0_kernel<<<(unsigned)((p->n + 127) / 128), 128, 0, p->stream>>>(zj);
        }
        sonic_integrate_kernel<<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(probe);
        TRYA(cudaGetLastError());
        TRYA(cudaMemcpyAsync(smid.data(), p->d_block_smid, blocks * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        TRYA(cudaStreamSynchronize(p->stream));
    }
    // Lane budgets (see the kernel).  Warps are walked SM by SM (all warps of all blocks of one
    // SM, then the next SM), each taking the next `cap` points of the sorted list, so that the
    // most expensive points end up alone in their warp on SMs that host nothing but such warps:
    // a lone lane ticks in t1 = 2.3 us there, but in 3-4 us next to full, phase-diverged warps
    // (instruction-cache and issue contention), and a warp with k busy lanes in about
    //   t(k) = t1 (1 + GAIN (1 - exp(-(k - 1) / KDEC))),   t(32) = 2.55 t1.
    // Every chain c should finish within (1 + MARGIN) c_max t1, so a warp whose most expensive
    // point has predicted chain length c may run at t <= (1 + MARGIN) (c_max / c) t1.
    std::vector<int> wfirst(nwarps, (int)n), wcap(nwarps, (int)lpw);
    {
        std::vector<int> border(blocks);
        for (int b = 0; b < (int)blocks; b++) border[b] = b;
        std::stable_sort(border.begin(), border.end(), [&](int x, int y) { return smid[x] < smid[y]; });
        const double GAIN = 1.55, KDEC = 2.5, T32 = 1.0 + GAIN;       // t(k) / t1, t(32) / t1
        std::vector<double> chain(n), tail(n + 1, 0.0);               // predicted ticks, suffix sums
        for (long long i = 0; i < n; i++) chain[i] = exp(cost[order[i]]);
        for (long long i = n - 1; i >= 0; i--) tail[i] = tail[i + 1] + chain[i];
        // lanes a warp may keep busy if its longest chain c has to finish within T (units of t1)
        auto cap_for = [&](double c, double T) {
            const double x = (T / c - 1.0) / GAIN;
            if (x >= 0.98) return 32;
            if (x <= 0.0) return 1;
            const int k = (int)(1.0 - KDEC * log(1.0 - x));
            return k < 1 ? 1 : (k > 32 ? 32 : k);
        };
        // Pick the deadline T: the budgeted warps finish by T by construction; whatever is not
        // handed out statically runs on the remaining warps at full width.  A small T isolates
        // the long chains but leaves few full-width warps; scan T upwards from the longest chain
        // and keep the T with the smallest predicted makespan max(T, queue time).
        double best_T = chain[0], best_span = 1e300;
        for (double T = chain[0] * 1.02; T < chain[0] * 40.0; T *= 1.04) {
            long long pos = 0;
            int used = 0;
            while (used < nwarps && pos < n) {
                const int k = cap_for(chain[pos], T);
                if (k >= (int)lpw) break;
                pos += k;
                used++;
            }
            const int full = nwarps - used;
            const double queue = full > 0 ? tail[pos] / (32.0 * full) * T32 : (pos < n ? 1e300 : 0.0);
            const double span = T > queue ? T : queue;
            if (span < best_span) { best_span = span; best_T = T; }
            if (queue <= T) break;        // larger T only makes the deadline later
        }
        long long pos = 0;
        for (int r = 0; r < nwarps; r++) {
            const int gw = border[r / warps_per_block] * warps_per_block + r % warps_per_block;
            if (pos >= n) continue;
            int cap = cap_for(chain[pos], best_T);
            if (cap > (int)lpw) cap = (int)lpw;
            wfirst[gw] = (int)pos;
            wcap[gw] = cap;
            pos += cap;
        }
        p->n_initial = (unsigned long long)(pos < n ? pos : n);
        p->probe_smid = smid;
        if (getenv("SONIC_DEBUG"))
            fprintf(stderr, "[sonic] schedule: longest chain %.3g ticks, deadline %.2f x, predicted makespan %.2f x\n",
                    chain[0], best_T / chain[0], best_span / chain[0]);
    }
    TRYA(cudaMemcpyAsync(p->d_warp_first, wfirst.data(), nwarps * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaMemcpyAsync(p->d_warp_cap, wcap.data(), nwarps * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    TRYA(cudaStreamSynchronize(p->stream));
#undef TRYA
    if (e != cudaSuccess) {
        plan_free(p);
        return set_err(e == cudaErrorMemoryAllocation ? SONIC_E_ALLOC : SONIC_E_CUDA,
                       "plan creation failed: %s", cudaGetErrorString(e));
    }
    p->ms_upload = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out_plan = p;
    return SONIC_OK;
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
    if (getenv("SONIC_DEBUG") && p->launched) {
        // placement of the previous launch against the probe (and the launch before)
        std::vector<int> now(p->grid);
        CUDA_TRY(cudaMemcpy(now.data(), p->d_block_smid, p->grid * sizeof(int), cudaMemcpyDeviceToHost));
        int d_probe = 0, d_prev = 0, same_sm_set = 0;
        for (int b = 0; b < p->grid; b++) {
            d_probe += now[b] != p->probe_smid[b];
            if (!p->last_smid.empty()) d_prev += now[b] != p->last_smid[b];
        }
        fprintf(stderr, "[sonic] placement of last launch: %d of %d blocks differ from the probe, %d from the launch before\n",
                d_probe, p->grid, d_prev);
        for (int b = 0; b < p->grid && same_sm_set < 12; b++)
            if (now[b] != p->probe_smid[b]) {
                fprintf(stderr, "    block %d: probe SM %d, real SM %d\n", b, p->probe_smid[b], now[b]);
                same_sm_set++;
            }
        p->last_smid = now;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_counter, &p->n_initial, sizeof(p->n_initial), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaEventRecord(p->ev[0], p->stream));
    sonic_z0_kernel<<<(unsigned)((p->n + 127) / 128), 128, 0, p->stream>>>(job);
    CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));
    sonic_integrate_kernel<<<p->grid, SONIC_BLOCK, SONIC_HIST_BYTES, p->stream>>>(job);
    CUDA_TRY(cudaEventRecord(p->ev[2], p->stream));
    {
        long long blocks = (p->n_out + SONIC_AVG_WARPS - 1) / SONIC_AVG_WARPS;
        if (blocks > 148LL * 64) blocks = 148LL * 64;
#define CALL(ID) launch_average<ID>(p, (int)blocks)
        SONIC_DISPATCH_NEURON(p->neuron_id, CALL)
#undef CALL
    }
    CUDA_TRY(cudaEventRecord(p->ev[3], p->stream));
    CUDA_TRY(cudaGetLastError());
    p->launches += 3;
    p->launched = true;
    return SONIC_OK;
}

int sonic_plan_set_stream(SonicPlan* p, void* stream) {
    if (!p) return set_err(SONIC_E_ARG, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (p->stream && p->own_stream) cudaStreamDestroy(p->stream);
    p->stream = static_cast<cudaStream_t>(stream);
    p->own_stream = false;
    return SONIC_OK;
}

int sonic_plan_sync(SonicPlan* p) {
    if (!p) return set_err(SONIC_E_ARG, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SONIC_OK;
}

int sonic_plan_fetch(SonicPlan* p, double* out_tables, int32_t* out_ncycles, uint32_t* out_status,
                     double* out_tpoint, uint32_t* out_nrhs) {
    if (!p || !p->launched) return set_err(SONIC_E_ARG, "plan not launched");
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = SONIC_OK;
    if (out_tables)
        CUDA_TRY(cudaMemcpyAsync(out_tables, p->d_out, (size_t)p->nvar * p->n_out * p->nfs * sizeof(double),
                                 cudaMemcpyDeviceToHost, p->stream));
    if (out_ncycles && !rc) rc = fetch_expanded(p, p->d_ncycles, out_ncycles);
    if (out_status && !rc) rc = fetch_expanded(p, p->d_status, out_status);
    if (out_tpoint && !rc) rc = fetch_expanded(p, p->d_tpoint, out_tpoint);
    if (out_nrhs && !rc) rc = fetch_expanded(p, p->d_nfe, out_nrhs);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(p->stream));
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
    CUDA_TRY(dalloc(&d_cm, total));
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
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    memset(st, 0, sizeof(*st));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1])); st->ms_z0 = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[1], p->ev[2])); st->ms_integrate = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[2], p->ev[3])); st->ms_average = ms;
    const size_t n = p->n;
    std::vector<unsigned> nfe(n), nje(n), nst(n);
    std::vector<int> ncyc(n);
    CUDA_TRY(cudaMemcpy(nfe.data(), p->d_nfe, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nje.data(), p->d_nje, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(nst.data(), p->d_nsteps, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(ncyc.data(), p->d_ncycles, n * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
        st->n_rhs += nfe[i]; st->n_jac += nje[i]; st->n_steps += nst[i]; st->n_cycles += ncyc[i];
    }
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
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    SonicPlan* p = nullptr;
    int rc = sonic_plan_create_ex(device, radii, na, neuron_id, n, ia, f, A, Q, novertones, overtones, fs, nfs, &p);
    if (rc) return rc;
    const double ms_create = ms_since(t0);
    rc = sonic_plan_launch(p);
    if (!rc) rc = sonic_plan_fetch(p, out_tables, out_ncycles, out_status, out_tpoint, out_nrhs);
    const double ms_run = ms_since(t0) - ms_create;
    if (!rc && stats) {
        rc = sonic_plan_stats(p, stats);
        stats->ms_total = ms_since(t0);
    }
    const double ms_stats = ms_since(t0) - ms_create - ms_run;
    plan_free(p);
    if (getenv("SONIC_DEBUG"))
        fprintf(stderr, "[sonic] points_run n=%lld: create %.1f ms, launch+fetch %.1f ms, stats %.1f ms, free %.1f ms\n",
                (long long)n, ms_create, ms_run, ms_stats, ms_since(t0) - ms_create - ms_run - ms_stats);
    return rc;
}

int sonic_lookup_run(const SonicBlsParams* radii, int na, const double* f, int nf, const double* A,
                     int nA, const double* Q, int nQ, const double* fs, int nfs, int neuron_id,
                     uint32_t device_mask, double* out_tables, int32_t* out_ncycles,
                     uint32_t* out_status, double* out_tpoint, SonicStats* stats) {
    if (!radii || !f || !A || !Q || !fs || na <= 0 || nf <= 0 || nA <= 0 || nQ <= 0 || nfs <= 0 || !out_tables)
        return set_err(SONIC_E_ARG, "invalid argument (null pointer or empty dimension)");
    if (neuron_id < 0 || neuron_id >= SONIC_N_NEURONS)
        return set_err(SONIC_E_NEURON, "unknown neuron id %d", neuron_id);
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
    const int64_t n = (int64_t)na * nf * nA * nQ;
    const int nvar = 1 + SONIC_NEURON_NRATES[neuron_id];
    // flatten the grid in the reference's queue order: a > f > A > Q
    std::vector<int32_t> pia(n);
    std::vector<double> pf(n), pA(n), pQ(n);
    int64_t i = 0;
    for (int x = 0; x < na; x++)
        for (int y = 0; y < nf; y++)
            for (int z = 0; z < nA; z++)
                for (int w = 0; w < nQ; w++, i++) {
                    pia[i] = x; pf[i] = f[y]; pA[i] = A[z]; pQ[i] = Q[w];
                }
    const int nd = (int)devs.size();
    if (nd == 1) {
        int rc = sonic_points_run(devs[0], radii, na, neuron_id, n, pia.data(), pf.data(), pA.data(), pQ.data(),
                                  fs, nfs, out_tables, out_ncycles, out_status, out_tpoint, nullptr, stats);
        if (!rc && stats)
            stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }
    // several devices: deal the cost-sorted points round-robin so every device gets the same
    // cost profile, one host thread per device, host-side scatter of the results
    std::vector<int> order(n);
    std::vector<double> cost(n);
    for (int64_t k = 0; k < n; k++) {
        order[k] = (int)k;
        cost[k] = predict_log_cost(radii[pia[k]].a, pf[k], pA[k], pQ[k]);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    struct Shard {
        std::vector<int> idx;
        std::vector<int32_t> ia, ncyc;
        std::vector<double> f, A, Q, out, tp;
        std::vector<uint32_t> st;
        SonicStats stats;
        int rc = 0;
        std::string err;
    };
    std::vector<Shard> sh(nd);
    for (int64_t k = 0; k < n; k++) sh[k % nd].idx.push_back(order[k]);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        Shard& s = sh[d];
        const size_t m = s.idx.size();
        s.ia.resize(m); s.f.resize(m); s.A.resize(m); s.Q.resize(m);
        s.out.resize((size_t)nvar * m * nfs); s.ncyc.resize(m); s.st.resize(m); s.tp.resize(m);
        for (size_t k = 0; k < m; k++) {
            const int g = s.idx[k];
            s.ia[k] = pia[g]; s.f[k] = pf[g]; s.A[k] = pA[g]; s.Q[k] = pQ[g];
        }
        th.emplace_back([&, d]() {
            Shard& s2 = sh[d];
            if (s2.idx.empty()) return;
            s2.rc = sonic_points_run(devs[d], radii, na, neuron_id, (int64_t)s2.idx.size(), s2.ia.data(),
                                     s2.f.data(), s2.A.data(), s2.Q.data(), fs, nfs, s2.out.data(),
                                     s2.ncyc.data(), s2.st.data(), s2.tp.data(), nullptr, &s2.stats);
            if (s2.rc) s2.err = g_err;
        });
    }
    for (auto& t : th) t.join();
    SonicStats tot;
    memset(&tot, 0, sizeof(tot));
    for (int d = 0; d < nd; d++) {
        Shard& s = sh[d];
        if (s.rc) return set_err(s.rc, "device %d: %s", devs[d], s.err.c_str());
        const size_t m = s.idx.size();
        for (size_t k = 0; k < m; k++) {
            const int64_t g = s.idx[k];
            for (int v = 0; v < nvar; v++)
                for (int j = 0; j < nfs; j++)
                    out_tables[((int64_t)v * n + g) * nfs + j] = s.out[((size_t)v * m + k) * nfs + j];
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

int sonic_trim(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return SONIC_OK;
    for (int d = 0; d < ndev && d < 64; d++)
        for (int w = 0; w < 2; w++) {
            double* ptr = nullptr;
            {
                std::lock_guard<std::mutex> lk(g_pool_mutex);
                ptr = g_pool[d][w].ptr;
                g_pool[d][w].ptr = nullptr;
                g_pool[d][w].count = 0;
            }
            if (ptr) {
                cudaSetDevice(d);
                cudaFree(ptr);
            }
        }
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
    CUDA_TRY(dalloc(&d_out, (size_t)blocks * threads));
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
