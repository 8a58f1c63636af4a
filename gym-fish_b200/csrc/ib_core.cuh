// ib_core.cuh — immersed-boundary coupling kernels and their host-side state.
//
// BASELINE.json:5 (b): velocity interpolation from the Eulerian grid to Lagrangian markers with the
// 4-point Peskin delta, direct forcing, force spreading back, Guo force folded into the collide.
// Step order is SURVEY.md A7 (the oracle's ib_forces()):
//   index map -> unforced moments of the band cells -> interpolate -> F_k = 2 rho0 (U_d - U*) -> spread -> wrenches
//
// Band-sparse storage: only cells inside some marker's 4x4x4 stencil exist in the band arrays.
//   cellslot[idx] = 1 + band position (0 = not in band), band_cell[pos] = idx,
//   band_u[3][cap] unforced velocity, bandF[3][cap] spread force, rowflag[(nz+2)*ny] rows with band cells.
// The stream-collide kernel reads the force through cellslot only in flagged rows, so a step without
// markers nearby costs no extra HBM traffic.
#pragma once
#include "lbm_core.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace fg {

FG_HD float atomic_add_f(float *p, float v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const float o = *p; *p = o + v; return o;
#endif
}
FG_HD double atomic_add_d(double *p, double v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const double o = *p; *p = o + v; return o;
#endif
}
FG_HD int atomic_add_i(int *p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const int o = *p; *p = o + v; return o;
#endif
}
FG_HD int atomic_cas_i(int *p, int cmp, int val) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, cmp, val);
#else
    const int o = *p; if (o == cmp) *p = val; return o;
#endif
}

// SURVEY.md A6
FG_HD float peskin4(float r) {
    r = fabsf(r);
    if (r < 1.0f) return (3.0f - 2.0f * r + sqrtf(1.0f + 4.0f * r - 4.0f * r * r)) * 0.125f;
    if (r < 2.0f) return (5.0f - 2.0f * r - sqrtf(fmaxf(0.0f, -7.0f + 12.0f * r - 4.0f * r * r))) * 0.125f;
    return 0.0f;
}

FG_HD int wrap_or_skip(int v, int n, bool periodic) {
    if (v >= 0 && v < n) return v;
    if (!periodic) return -1;
    v %= n;
    return v < 0 ? v + n : v;
}

struct IbParams {
    Lattice L;
    Collision C;
    int n;                       // markers
    int per_x, per_y, per_z;     // periodic axes
    const float *X, *U, *dV;     // [n][3], [n][3], [n]
    const int *link;             // [n]
    int *base;                   // [n][3]
    int *owner;                  // [n]
    float *Fm, *Ustar;           // [n][3]
    float *mdfE;                 // [n][3] multi-direct forcing: the spread force gathered back at the markers (IbMdfGather), or nullptr
    int *cellslot;               // [(nz+2)*plane]
    int *band_cell;              // [cap]
    float *band_u;               // [3][cap]
    float *bandF;                // [3][cap]
    int *band_count;             // this step's band size (device counter)
    int *band_count_next;        // the other counter: previous step's size until IbClearBand ran, then zeroed for the next
    int band_cap;
    int band_ctas;               // grid of the band kernels (IbBandMoments, IbClearBand): they walk the band grid-stride
    uint8_t *rowflag;            // [(nz+2)*ny]
    // static bodies: (stencil weight, band slot) of every stencil node, cached while the marker set is not re-sent
    float *node_w;               // [n][64] or nullptr
    int *node_s;                 // [n][64]
    int cache_mode;              // 0: compute; 1: compute and fill the cache; 2: read the cache
    const float *origin;         // [links][3] torque reference points
    double *wrench;              // [links][6]
    int n_links;
    // ---- bodies crossing z-slab faces (n_ranks > 1, fg_peer_connect_all): see IbExchange below
    const int *gidx;             // [n] global marker id of local marker k (multi-rank culling), < 0: padding; nullptr: identity
    int rank, n_ranks;           // n_ranks <= 1: no exchange
    int cap, maxl;               // marker / link capacity (buffer strides)
    int *xside;                  // [n] 0: stencil inside my slab (or not mine at all), 1: also in my z-low neighbour, 2: z-high
    const int *xcounter;         // my exchange counter (device-resident, bumped once per IB step)
    const float *ustar_in;       // mine: [2 sets][2 sides][cap][3] partial U* pushed by the z-low (side 0) / z-high (side 1) neighbour
    float *peer_in_lo, *peer_in_hi;   // the same buffer of my z-low / z-high neighbour (peer memory) or nullptr
    const double *wrench_all;    // mine: [2 sets][n_ranks][maxl][6] partial wrenches pushed by every rank
    double *peer_wrench[8];      // wrench_all of every rank (peer memory; [rank] is my own)
};
constexpr int kMaxRanks = 8;

// storage index of stencil node (a,b,c) of a marker with base (i0,j0,k0), or -1 when the node is outside
FG_HD long long stencil_cell(const IbParams &p, int i0, int j0, int k0, int a, int b, int c) {
    const Lattice &L = p.L;
    const int zg = wrap_or_skip(k0 + c, L.nzg, p.per_z != 0);
    if (zg < 0) return -1;
    const int zl = zg - L.z0;
    if (zl < 0 || zl >= L.nz) return -1;
    const int y = wrap_or_skip(j0 + b, L.ny, p.per_y != 0);
    if (y < 0) return -1;
    const int x = wrap_or_skip(i0 + a, L.nx, p.per_x != 0);
    if (x < 0) return -1;
    return ((long long)(zl + 1) * L.ny + y) * L.nx + x;
}

// Thread mapping of the marker kernels: one thread per stencil node, 64 consecutive threads per marker
// (a = x offset fastest, so a warp touches runs of 4 consecutive cells), 2 markers per 128-thread CTA.
constexpr int kNodes = 64;
constexpr int kMarkersPerCta = 2;

// sum over the 32 lanes of a warp, result valid in lane 0 (device); identity in the host emulation, where every
// "thread" then adds its own partial atomically — same sum, different association
FG_HD float warp_sum(float v) {
#if defined(__CUDA_ARCH__)
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
#endif
    return v;
}
FG_HD double warp_sum(double v) {
#if defined(__CUDA_ARCH__)
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
#endif
    return v;
}
FG_HD bool reduction_leader(int tx) {
#if defined(__CUDA_ARCH__)
    return (tx & 31) == 0;
#else
    (void)tx;
    return true;
#endif
}

// (a3) marker -> grid index map (integer outputs bit-exact with the oracle) + band registration.
// Band positions come from ONE global counter.  CTA_AGG (the product's launch): the CTA counts its new cells in shared
// memory and reserves its range with a single atomic — 8 markers (512 threads) per CTA, so 1e5 markers cost 12k
// same-address atomics instead of one per warp (the first version spent 115 us there, profiles/r1_ncu_ib.csv).
template <bool CTA_AGG>
struct IbIndexMarkT {
    static constexpr int kMarkers = CTA_AGG ? 8 : kMarkersPerCta;
    static constexpr int kThreads = kNodes * kMarkers;
    static constexpr int kMinBlocks = CTA_AGG ? 2 : 8;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kMarkers + tx / kNodes, node = tx % kNodes;
        const bool live = k < p.n && !(p.gidx && p.gidx[k] < 0);
        const Lattice &L = p.L;
        long long cell = -1;
        if (live) {
            const float X = p.X[3 * k], Y = p.X[3 * k + 1], Z = p.X[3 * k + 2];
            const int i0 = int(floorf(X)) - 1, j0 = int(floorf(Y)) - 1, k0 = int(floorf(Z)) - 1;
            if (node == 0) {
                p.base[3 * k] = i0; p.base[3 * k + 1] = j0; p.base[3 * k + 2] = k0;
                int kc = wrap_or_skip(k0 + 1, L.nzg, p.per_z != 0);
                if (kc < 0) kc = k0 + 1 < 0 ? 0 : L.nzg - 1;
                p.owner[k] = kc / L.nz;
            }
            if (node < 3) p.Ustar[3 * k + node] = 0.0f;          // accumulated by IbInterpolate
            cell = stencil_cell(p, i0, j0, k0, node & 3, (node >> 2) & 3, node >> 4);
        }
        const bool won = cell >= 0 && p.cellslot[cell] == 0 && atomic_cas_i(&p.cellslot[cell], 0, -1) == 0;
        int pos = -1;
#if defined(__CUDA_ARCH__)
        if (CTA_AGG) {
            __shared__ int s_count, s_base;
            if (tx == 0) s_count = 0;
            __syncthreads();
            const int mine = won ? atomicAdd(&s_count, 1) : 0;
            __syncthreads();
            if (tx == 0 && s_count > 0) s_base = atomicAdd(p.band_count, s_count);
            __syncthreads();
            if (won) pos = s_base + mine;
        } else
#endif
        if (won) pos = atomic_add_i(p.band_count, 1);
        if (won && pos < p.band_cap) {
            p.band_cell[pos] = int(cell);
            p.cellslot[cell] = pos + 1;
            p.rowflag[cell / L.nx] = 1;
        }
    }
};
using IbIndexMark = IbIndexMarkT<false>;      // per-thread form: the cooperative single-kernel variant and the host emulation
#if defined(__CUDACC__)
using IbIndexMarkLaunch = IbIndexMarkT<true>;
#else
using IbIndexMarkLaunch = IbIndexMarkT<false>;
#endif

// (a4) unforced velocity of the band cells from the populations arriving at time t (parity aware).
// The band's size is a device counter the host only has a bound for (64 cells per marker; ~16 are distinct on a closed
// surface), so the kernel walks the band grid-stride with a grid sized for the SMs: a grid sized for the bound spent a
// quarter of its time dispatching CTAs that found nothing to do (49 536 CTAs for 1.6 M band cells at 1e5 markers).
template <int PARITY>
struct IbBandMoments {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    static constexpr int kCtasPerSm = 12;                     // 40 registers: 12 x 128 threads resident per SM
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        if (bx == 0 && tx == 0) *p.band_count_next = 0;       // the other counter is free again (IbClearBand has run)
        const int cnt = *p.band_count < p.band_cap ? *p.band_count : p.band_cap;
        const Lattice &L = p.L;
        for (long long pos = (long long)bx * kThreads + tx; pos < cnt; pos += (long long)(p.band_ctas > 0 ? p.band_ctas : 1) * kThreads) {
            const long long idx = p.band_cell[pos];
            const int x = int(idx % L.nx), y = int((idx / L.nx) % L.ny), zz = int(idx / L.plane);
            const Nbr nb = make_nbr(L, x, y, zz);
            float h[Q];
            if (L.solid && L.solid[idx]) {
                FG_UNROLL
                for (int i = 0; i < Q; ++i) h[i] = pop_ld(L.f + i * L.slot + idx);   // obstacles keep their initial state (as in the oracle)
            } else {
                load_arriving<PARITY, true>(h, L, p.C, nb, idx);
            }
            float dr, jx, jy, jz;
            moments(h, dr, jx, jy, jz);
            const float inv = 1.0f / (1.0f + dr);
            p.band_u[pos] = jx * inv;
            p.band_u[p.band_cap + pos] = jy * inv;
            p.band_u[2 * p.band_cap + pos] = jz * inv;
            p.bandF[pos] = 0.0f; p.bandF[p.band_cap + pos] = 0.0f; p.bandF[2 * p.band_cap + pos] = 0.0f;
        }
    }
};

FG_HD float node_weight(const IbParams &p, int k, int node, long long &cell, int &slot) {
    if (p.cache_mode == 2) {             // static body: weight and band slot of this node were stored by an earlier step
        slot = p.node_s[(size_t)k * kNodes + node];
        cell = slot >= 0 ? (long long)p.band_cell[slot] : -1;
        return p.node_w[(size_t)k * kNodes + node];
    }
    const float X = p.X[3 * k], Y = p.X[3 * k + 1], Z = p.X[3 * k + 2];
    const int i0 = p.base[3 * k], j0 = p.base[3 * k + 1], k0 = p.base[3 * k + 2];
    const int a = node & 3, b = (node >> 2) & 3, c = node >> 4;
    cell = stencil_cell(p, i0, j0, k0, a, b, c);
    slot = cell >= 0 ? p.cellslot[cell] - 1 : -1;
    const float w = peskin4(X - float(i0 + a)) * peskin4(Y - float(j0 + b)) * peskin4(Z - float(k0 + c));
    if (p.cache_mode == 1) { p.node_w[(size_t)k * kNodes + node] = w; p.node_s[(size_t)k * kNodes + node] = slot; }
    return w;
}

FG_HD float ld_cg(const float *p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);       // L2: the value was just produced by atomics of other warps of this CTA
#else
    return *p;
#endif
}

// (a5) U*_k = sum over the 64 nodes of w u*: warp-reduced, 2 x 3 atomics per marker
struct IbInterpolate {
    static constexpr int kThreads = kNodes * kMarkersPerCta;
    static constexpr int kMinBlocks = 8;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int gt = bx * kThreads + tx;
        if (gt < 6 * p.n_links) p.wrench[gt] = 0.0;           // accumulated by IbLinkReduce
        const int k = bx * kMarkersPerCta + tx / kNodes, node = tx % kNodes;
        if (k >= p.n || (p.gidx && p.gidx[k] < 0)) return;    // whole warps leave together (64 threads per marker)
        long long cell; int s;
        const float w = node_weight(p, k, node, cell, s);
        float u0 = 0.f, u1 = 0.f, u2 = 0.f;
        if (s >= 0) { u0 = w * p.band_u[s]; u1 = w * p.band_u[p.band_cap + s]; u2 = w * p.band_u[2 * p.band_cap + s]; }
        u0 = warp_sum(u0); u1 = warp_sum(u1); u2 = warp_sum(u2);
        if (reduction_leader(tx)) {
            atomic_add_f(&p.Ustar[3 * k], u0); atomic_add_f(&p.Ustar[3 * k + 1], u1); atomic_add_f(&p.Ustar[3 * k + 2], u2);
        }
    }
};

// (a6, a7) direct forcing F_k = 2 rho0 (U_d - U*) and spreading F(x) += F_k w dV: one reduction atomic per node and
// component; nodes of one marker are distinct cells, so a warp never collides with itself
struct IbForceSpread {
    static constexpr int kThreads = kNodes * kMarkersPerCta;
    static constexpr int kMinBlocks = 8;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kMarkersPerCta + tx / kNodes, node = tx % kNodes;
        if (k >= p.n || (p.gidx && p.gidx[k] < 0)) return;
        const float f0 = 2.0f * (p.U[3 * k] - ld_cg(&p.Ustar[3 * k])), f1 = 2.0f * (p.U[3 * k + 1] - ld_cg(&p.Ustar[3 * k + 1])),
                    f2 = 2.0f * (p.U[3 * k + 2] - ld_cg(&p.Ustar[3 * k + 2]));
        if (node == 0) { p.Fm[3 * k] = f0; p.Fm[3 * k + 1] = f1; p.Fm[3 * k + 2] = f2; }
        long long cell; int s;
        const float w = node_weight(p, k, node, cell, s) * p.dV[k];
        if (s < 0) return;
        atomic_add_f(&p.bandF[s], w * f0);
        atomic_add_f(&p.bandF[p.band_cap + s], w * f1);
        atomic_add_f(&p.bandF[2 * p.band_cap + s], w * f2);
    }
};

// (a6, optional) multi-direct forcing, SURVEY.md A7 (4) "n_iter > 1" (FgConfig.ib_iterations; the oracle's loop in ib_forces()).
// The collide works with u = u* + F/(2 rho0) (Guo half-force), so after one direct-forcing pass the markers see
// U*_k + E_k/2 with E_k = sum_x F(x) delta_h(x - X_k) instead of U_d,k.  Every further pass is two launches over the same
// stencils: IbMdfGather collects E_k from the band force spread so far (all markers, before any of them spreads again),
// IbMdfSpread adds the Jacobi correction dF_k = 2 rho0 (U_d,k - U*_k) - E_k to the marker force and spreads it.
struct IbMdfGather {
    static constexpr int kThreads = kNodes * kMarkersPerCta;
    static constexpr int kMinBlocks = 8;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kMarkersPerCta + tx / kNodes, node = tx % kNodes;
        if (k >= p.n || (p.gidx && p.gidx[k] < 0)) return;    // whole warps leave together (64 threads per marker)
        long long cell; int s;
        const float w = node_weight(p, k, node, cell, s);
        float e0 = 0.f, e1 = 0.f, e2 = 0.f;
        if (s >= 0) { e0 = w * p.bandF[s]; e1 = w * p.bandF[p.band_cap + s]; e2 = w * p.bandF[2 * p.band_cap + s]; }
        e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2);
        if (reduction_leader(tx)) {
            atomic_add_f(&p.mdfE[3 * k], e0); atomic_add_f(&p.mdfE[3 * k + 1], e1); atomic_add_f(&p.mdfE[3 * k + 2], e2);
        }
    }
};
struct IbMdfSpread {
    static constexpr int kThreads = kNodes * kMarkersPerCta;
    static constexpr int kMinBlocks = 8;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kMarkersPerCta + tx / kNodes, node = tx % kNodes;
        if (k >= p.n || (p.gidx && p.gidx[k] < 0)) return;
        const float d0 = 2.0f * (p.U[3 * k] - p.Ustar[3 * k]) - p.mdfE[3 * k], d1 = 2.0f * (p.U[3 * k + 1] - p.Ustar[3 * k + 1]) - p.mdfE[3 * k + 1],
                    d2 = 2.0f * (p.U[3 * k + 2] - p.Ustar[3 * k + 2]) - p.mdfE[3 * k + 2];
        if (node == 0) { p.Fm[3 * k] += d0; p.Fm[3 * k + 1] += d1; p.Fm[3 * k + 2] += d2; }   // no other thread reads Fm in this launch
        long long cell; int s;
        const float w = node_weight(p, k, node, cell, s) * p.dV[k];
        if (s < 0) return;
        atomic_add_f(&p.bandF[s], w * d0);
        atomic_add_f(&p.bandF[p.band_cap + s], w * d1);
        atomic_add_f(&p.bandF[2 * p.band_cap + s], w * d2);
    }
};

// (a5-a7) in one launch (single-rank path; across slabs the exchange sits between the two kernels above).
// run(): the two phases as the host emulation executes them.  cta(): what the GPU runs — ONE WARP PER MARKER, two stencil
// nodes per lane (planes c and c + 2 of the 4 x 4 x 4 stencil).  U*_k is then a butterfly reduction inside the warp: no
// atomics on U*, no fence, no CTA barrier between interpolation and spreading — the round-1 form (64 threads per marker,
// two warps meeting through L2 atomics + __threadfence + __syncthreads) was a chain of four dependent L2 round trips
// per CTA and ran at 117 us for 1e5 markers with the memory system idle (profiles/r2_summary.md).  The three 1-D delta
// weights of a node come by shuffle from the 12 lanes that evaluated them; static bodies read (weight, slot) of their
// nodes from the per-node cache instead (cache_mode 2).
struct IbInterpSpread {
    static constexpr int kThreads = 128;
    static constexpr int kMarkers = kThreads / 32;       // markers per CTA on the GPU
    static constexpr int kMinBlocks = 16;                // 28 registers: the full 2048 threads per SM
    static constexpr int kBlockPhases = 2;
    // host emulation: launched with the grid of the per-thread kernels (2 markers x 64 nodes per block)
    FG_HD static void run(const IbParams &p, int bx, int by, int bz, int tx, int phase) {
        if (phase == 0) IbInterpolate::run(p, bx, by, bz, tx);
        else IbForceSpread::run(p, bx, by, bz, tx);
    }
#if defined(__CUDACC__)
    __device__ __forceinline__ static float warp_all_sum(float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
    }
    // interpolation + direct forcing of marker k by one warp: (weight, band slot) of the lane's two nodes and the marker force
    struct Nodes { float w0, w1, f0, f1, f2; int s0, s1; };
    __device__ __forceinline__ static Nodes marker_force(const IbParams &p, int k, int lane) {
        Nodes r;
        float w0, w1;
        int s0, s1;
        const size_t nb = (size_t)k * kNodes;
        if (p.cache_mode == 2) {
            // static body: two coalesced 8-byte reads per lane instead of the delta weights, the wrap logic and two
            // scattered cellslot gathers (cellslot is a dense 4 B-per-cell array: every gather costs a DRAM sector)
            w0 = p.node_w[nb + lane]; w1 = p.node_w[nb + 32 + lane];
            s0 = p.node_s[nb + lane]; s1 = p.node_s[nb + 32 + lane];
        } else {
            const int sel = lane % 12, axis = sel >> 2, j = sel & 3;
            const float w1d = peskin4(p.X[3 * k + axis] - float(p.base[3 * k + axis] + j));
            const int a = lane & 3, b = (lane >> 2) & 3, c = lane >> 4;          // node `lane`: plane c; node `lane + 32`: plane c + 2
            const float wxy = __shfl_sync(0xffffffffu, w1d, a) * __shfl_sync(0xffffffffu, w1d, 4 + b);
            w0 = wxy * __shfl_sync(0xffffffffu, w1d, 8 + c);
            w1 = wxy * __shfl_sync(0xffffffffu, w1d, 10 + c);
            const int i0 = p.base[3 * k], j0 = p.base[3 * k + 1], k0 = p.base[3 * k + 2];
            const long long c0 = stencil_cell(p, i0, j0, k0, a, b, c), c1 = stencil_cell(p, i0, j0, k0, a, b, c + 2);
            s0 = c0 >= 0 ? p.cellslot[c0] - 1 : -1;
            s1 = c1 >= 0 ? p.cellslot[c1] - 1 : -1;
            if (p.cache_mode == 1) {
                p.node_w[nb + lane] = w0; p.node_w[nb + 32 + lane] = w1;
                p.node_s[nb + lane] = s0; p.node_s[nb + 32 + lane] = s1;
            }
        }
        float u0 = 0.f, u1 = 0.f, u2 = 0.f;
        if (s0 >= 0) { u0 = w0 * p.band_u[s0]; u1 = w0 * p.band_u[p.band_cap + s0]; u2 = w0 * p.band_u[2 * p.band_cap + s0]; }
        if (s1 >= 0) { u0 += w1 * p.band_u[s1]; u1 += w1 * p.band_u[p.band_cap + s1]; u2 += w1 * p.band_u[2 * p.band_cap + s1]; }
        u0 = warp_all_sum(u0); u1 = warp_all_sum(u1); u2 = warp_all_sum(u2);          // every lane holds U*_k
        r.f0 = 2.0f * (p.U[3 * k] - u0); r.f1 = 2.0f * (p.U[3 * k + 1] - u1); r.f2 = 2.0f * (p.U[3 * k + 2] - u2);
        if (lane == 0) {
            p.Ustar[3 * k] = u0; p.Ustar[3 * k + 1] = u1; p.Ustar[3 * k + 2] = u2;
            p.Fm[3 * k] = r.f0; p.Fm[3 * k + 1] = r.f1; p.Fm[3 * k + 2] = r.f2;
        }
        const float dV = p.dV[k];
        r.w0 = w0 * dV; r.w1 = w1 * dV; r.s0 = s0; r.s1 = s1;
        return r;
    }
    __device__ __forceinline__ static void cta(const IbParams &p, int bx, int tx) {
        const int gt = bx * kThreads + tx;
        if (gt < 6 * p.n_links) p.wrench[gt] = 0.0;           // accumulated by IbLinkReduce
        const int k = bx * kMarkers + (tx >> 5), lane = tx & 31;
        if (k >= p.n || (p.gidx && p.gidx[k] < 0)) return;    // warp-uniform
        const Nodes r = marker_force(p, k, lane);
        if (r.s0 >= 0) {
            atomicAdd(&p.bandF[r.s0], r.w0 * r.f0); atomicAdd(&p.bandF[p.band_cap + r.s0], r.w0 * r.f1); atomicAdd(&p.bandF[2 * p.band_cap + r.s0], r.w0 * r.f2);
        }
        if (r.s1 >= 0) {
            atomicAdd(&p.bandF[r.s1], r.w1 * r.f0); atomicAdd(&p.bandF[p.band_cap + r.s1], r.w1 * r.f1); atomicAdd(&p.bandF[2 * p.band_cap + r.s1], r.w1 * r.f2);
        }
    }
#endif
};

// A/B variant (FG_FLAG_IB_TILE_SPREAD; BASELINE.json:5 (b) "shared-memory-staged tile accumulation"): the spread of a CTA's
// markers is first accumulated in a shared-memory table keyed by band slot (open addressing, 512 entries for at most
// 4 x 64 nodes), then every occupied entry goes to global memory with ONE reduction per component.  Markers that are
// neighbours on the body surface share most of their 4 x 4 x 4 stencils, so a spatially ordered marker list needs ~2.3 x
// fewer global atomics; a list without locality (the Fibonacci spheres of bench.py in generation order) pays the staging
// and saves nothing.  A warp holds ONE marker whose 64 nodes are distinct cells, so warp-level aggregation
// (__match_any_sync) has nothing to merge by construction.  Measured A/B: profiles/r2_summary.md.
struct IbInterpSpreadTile {
    static constexpr int kThreads = 128;
    static constexpr int kMarkers = kThreads / 32;
    static constexpr int kMinBlocks = 12;
    static constexpr int kBlockPhases = 2;
    static constexpr int kSlots = 512;
    FG_HD static void run(const IbParams &p, int bx, int by, int bz, int tx, int phase) { IbInterpSpread::run(p, bx, by, bz, tx, phase); }
#if defined(__CUDACC__)
    __device__ __forceinline__ static void stage(int *keys, float (*vals)[kSlots], int s, float a, float b, float c) {
        unsigned h = (unsigned(s) * 2654435761u) >> 23;
        for (;;) {
            const int old = atomicCAS(&keys[h], -1, s);
            if (old == -1 || old == s) break;
            h = (h + 1) & (kSlots - 1);
        }
        atomicAdd(&vals[0][h], a); atomicAdd(&vals[1][h], b); atomicAdd(&vals[2][h], c);
    }
    __device__ __forceinline__ static void cta(const IbParams &p, int bx, int tx) {
        __shared__ int keys[kSlots];
        __shared__ float vals[3][kSlots];
        for (int i = tx; i < kSlots; i += kThreads) { keys[i] = -1; vals[0][i] = 0.f; vals[1][i] = 0.f; vals[2][i] = 0.f; }
        const int gt = bx * kThreads + tx;
        if (gt < 6 * p.n_links) p.wrench[gt] = 0.0;
        __syncthreads();
        const int k = bx * kMarkers + (tx >> 5), lane = tx & 31;
        if (k < p.n && !(p.gidx && p.gidx[k] < 0)) {          // warp-uniform
            const IbInterpSpread::Nodes r = IbInterpSpread::marker_force(p, k, lane);
            if (r.s0 >= 0) stage(keys, vals, r.s0, r.w0 * r.f0, r.w0 * r.f1, r.w0 * r.f2);
            if (r.s1 >= 0) stage(keys, vals, r.s1, r.w1 * r.f0, r.w1 * r.f1, r.w1 * r.f2);
        }
        __syncthreads();
        for (int i = tx; i < kSlots; i += kThreads) {
            const int s = keys[i];
            if (s >= 0) { atomicAdd(&p.bandF[s], vals[0][i]); atomicAdd(&p.bandF[p.band_cap + s], vals[1][i]); atomicAdd(&p.bandF[2 * p.band_cap + s], vals[2][i]); }
        }
    }
#endif
};

// (a8) hydrodynamic wrench ON each link = minus what its markers exert on the fluid; markers arrive sorted by link,
// so a warp usually holds one link: shuffle-reduce in fp64, 6 atomics per warp
struct IbLinkReduce {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kThreads + tx;
        const bool live = k < p.n && !(p.gidx && p.gidx[k] < 0);
        int l = live ? p.link[k] : -1;
        if (l >= p.n_links) l = -1;
        if (live && p.n_ranks > 1 && p.owner[k] != p.rank) l = -1;   // across slabs every marker is reduced by its owner only
        double v[6] = {0, 0, 0, 0, 0, 0};
        if (l >= 0) {
            const double dV = p.dV[k];
            const double fx = -double(p.Fm[3 * k]) * dV, fy = -double(p.Fm[3 * k + 1]) * dV, fz = -double(p.Fm[3 * k + 2]) * dV;
            const double rx = double(p.X[3 * k]) - double(p.origin[3 * l]), ry = double(p.X[3 * k + 1]) - double(p.origin[3 * l + 1]),
                         rz = double(p.X[3 * k + 2]) - double(p.origin[3 * l + 2]);
            v[0] = fx; v[1] = fy; v[2] = fz;
            v[3] = ry * fz - rz * fy; v[4] = rz * fx - rx * fz; v[5] = rx * fy - ry * fx;
        }
#if defined(__CUDA_ARCH__)
        const int l0 = __shfl_sync(0xffffffffu, l, 0);
        if (__all_sync(0xffffffffu, l == l0 || l < 0) ) {
            if (l0 < 0) { /* lane 0 idle: fall through to the per-lane path below for the stragglers */ }
            else {
                FG_UNROLL
                for (int c = 0; c < 6; ++c) v[c] = warp_sum(v[c]);
                if ((tx & 31) == 0) {
                    FG_UNROLL
                    for (int c = 0; c < 6; ++c) atomicAdd(p.wrench + 6 * l0 + c, v[c]);
                }
                return;
            }
        }
#endif
        if (l >= 0)
            for (int c = 0; c < 6; ++c) atomic_add_d(p.wrench + 6 * l + c, v[c]);
    }
};

// fluid state at arbitrary points (fg_probe): one thread per stencil node gathers the arriving populations of its
// cell (parity aware), the 64 nodes of a probe are reduced with shuffles + 2 x 4 atomics
struct ProbeParams {
    Lattice L;
    Collision C;
    int n, per_x, per_y, per_z;
    const float *X;     // [n][3]
    float *out;         // [n][4], zeroed before the launch
};
template <int PARITY>
struct ProbeMoments {
    static constexpr int kThreads = kNodes * kMarkersPerCta;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const ProbeParams &p, int bx, int, int, int tx) {
        const int k = bx * kMarkersPerCta + tx / kNodes, node = tx % kNodes;
        if (k >= p.n) return;
        const Lattice &L = p.L;
        const float X = p.X[3 * k], Y = p.X[3 * k + 1], Z = p.X[3 * k + 2];
        const int i0 = int(floorf(X)) - 1, j0 = int(floorf(Y)) - 1, k0 = int(floorf(Z)) - 1;
        const int a = node & 3, b = (node >> 2) & 3, c = node >> 4;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        const int zg = wrap_or_skip(k0 + c, L.nzg, p.per_z != 0);
        const int zl = zg < 0 ? -1 : zg - L.z0;
        const int y = wrap_or_skip(j0 + b, L.ny, p.per_y != 0), x = wrap_or_skip(i0 + a, L.nx, p.per_x != 0);
        if (zl >= 0 && zl < L.nz && y >= 0 && x >= 0) {
            const int zz = zl + 1;
            const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
            const Nbr nb = make_nbr(L, x, y, zz);
            float h[Q];
            if (L.solid && L.solid[idx]) {
                FG_UNROLL
                for (int i = 0; i < Q; ++i) h[i] = pop_ld(L.f + i * L.slot + idx);
            } else {
                load_arriving<PARITY, true>(h, L, p.C, nb, idx);
            }
            float dr, jx, jy, jz;
            moments(h, dr, jx, jy, jz);
            const float inv = 1.0f / (1.0f + dr);
            const float w = peskin4(X - float(i0 + a)) * peskin4(Y - float(j0 + b)) * peskin4(Z - float(k0 + c));
            v0 = w * (1.0f + dr); v1 = w * jx * inv; v2 = w * jy * inv; v3 = w * jz * inv;
        }
        v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
        if (reduction_leader(tx)) {
            atomic_add_f(&p.out[4 * k], v0); atomic_add_f(&p.out[4 * k + 1], v1);
            atomic_add_f(&p.out[4 * k + 2], v2); atomic_add_f(&p.out[4 * k + 3], v3);
        }
    }
};

// forget the band of the previous step (grid-stride like IbBandMoments)
struct IbClearBand {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int cnt = *p.band_count_next < p.band_cap ? *p.band_count_next : p.band_cap;   // the previous step's counter
        for (long long pos = (long long)bx * kThreads + tx; pos < cnt; pos += (long long)(p.band_ctas > 0 ? p.band_ctas : 1) * kThreads) {
            const long long idx = p.band_cell[pos];
            p.cellslot[idx] = 0;
            p.rowflag[idx / p.L.nx] = 0;
        }
    }
};

// ---------------------------------------------------------------- bodies across z-slab faces
// Every rank holds the same marker list.  A 4-wide stencil spans at most two slabs: the partial sum U*_k of each rank
// is pushed into the other rank's ustar_in (peer store over NVLink), double-buffered by the parity of the exchange
// counter so that a rank one step ahead never overwrites what its neighbour still reads.  Link wrenches are reduced
// by the marker's owner and all-gathered by peer stores, then summed in rank order, so the (replicated) host body
// integrators of all ranks see bit-identical totals.
FG_HD int marker_rank(const IbParams &p, int zplane) {
    const int zg = wrap_or_skip(zplane, p.L.nzg, p.per_z != 0);
    return zg < 0 ? -1 : zg / p.L.nz;
}

struct IbPushPartial {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kThreads + tx;
        if (k >= p.n) return;
        const int gk = p.gidx ? p.gidx[k] : k;               // exchange buffers are indexed by the GLOBAL marker id
        if (gk < 0) { p.xside[k] = 0; return; }
        const int k0 = p.base[3 * k + 2];
        const int ra = marker_rank(p, k0), rb = marker_rank(p, k0 + 3);
        int xs = 0;
        if (ra != rb) {
            if (rb == p.rank && ra >= 0) xs = 1;        // the lower part of the stencil lies in my z-low neighbour
            else if (ra == p.rank && rb >= 0) xs = 2;   // the upper part lies in my z-high neighbour
        }
        p.xside[k] = xs;
        const int set = (p.xcounter[0] + 1) & 1;
        float *dst = xs == 1 ? p.peer_in_lo : (xs == 2 ? p.peer_in_hi : nullptr);
        if (!dst) return;
        const int side = xs == 1 ? 1 : 0;               // I am the receiver's z-high (1) or z-low (0) neighbour
        float *q = dst + ((size_t)(set * 2 + side) * p.cap + gk) * 3;
        q[0] = p.Ustar[3 * k]; q[1] = p.Ustar[3 * k + 1]; q[2] = p.Ustar[3 * k + 2];
    }
};

struct IbAddPartial {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kThreads + tx;
        if (k >= p.n) return;
        const int xs = p.xside[k];
        if (xs == 0) return;
        const int gk = p.gidx ? p.gidx[k] : k;
        const int set = (p.xcounter[0] + 1) & 1, side = xs == 1 ? 0 : 1;   // what my z-low (0) / z-high (1) neighbour sent
        const float *q = p.ustar_in + ((size_t)(set * 2 + side) * p.cap + gk) * 3;
        p.Ustar[3 * k] += q[0]; p.Ustar[3 * k + 1] += q[1]; p.Ustar[3 * k + 2] += q[2];
    }
};

struct IbPushWrench {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int i = bx * kThreads + tx;
        if (i >= 6 * p.n_links) return;
        const int set = (p.xcounter[0] + 1) & 1;
        const double v = p.wrench[i];
        for (int r = 0; r < p.n_ranks; ++r)
            if (p.peer_wrench[r]) p.peer_wrench[r][((size_t)(set * p.n_ranks + p.rank) * p.maxl) * 6 + i] = v;
    }
};

struct IbSumWrench {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int i = bx * kThreads + tx;
        if (i >= 6 * p.n_links) return;
        const int set = (p.xcounter[0] + 1) & 1;
        double sum = 0.0;
        for (int r = 0; r < p.n_ranks; ++r) sum += p.wrench_all[((size_t)(set * p.n_ranks + r) * p.maxl) * 6 + i];
        p.wrench[i] = sum;
    }
};

// The whole IB pipeline as ONE cooperative launch: the phases are the kernel bodies above, separated by grid-wide
// barriers instead of kernel boundaries (6 launches of ~3-5 us each were a fifth of the step on the 256x128x128
// workload).  Phase work counts are rounded up to whole CTAs so warps stay intact for the shuffle reductions.
template <int PARITY>
struct IbFused {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 8;
    static constexpr int kPhases = 6;
    FG_HD static long long round_cta(long long n) { return (n + kThreads - 1) / kThreads * kThreads; }
    FG_HD static long long items(const IbParams &p, int phase) {
        const long long nodes = round_cta((long long)(p.n + kMarkersPerCta - 1) / kMarkersPerCta * kThreads);
        switch (phase) {
            case 0: { const int c = *p.band_count_next; return round_cta(c < p.band_cap ? c : p.band_cap); }
            case 1: return nodes;
            case 2: { const int c = *p.band_count; return round_cta((c < p.band_cap ? c : p.band_cap) > 0 ? (c < p.band_cap ? c : p.band_cap) : 1); }
            case 3: { const long long w = round_cta(6ll * p.n_links); return nodes > w ? nodes : w; }
            case 4: return nodes;
            default: return round_cta(p.n);
        }
    }
    FG_HD static void item(const IbParams &p, int phase, long long i) {
        const int bx = int(i / kThreads), tx = int(i % kThreads);
        switch (phase) {
            case 0: IbClearBand::run(p, bx, 0, 0, tx); break;
            case 1: IbIndexMark::run(p, bx, 0, 0, tx); break;
            case 2: IbBandMoments<PARITY>::run(p, bx, 0, 0, tx); break;
            case 3: IbInterpolate::run(p, bx, 0, 0, tx); break;
            case 4: IbForceSpread::run(p, bx, 0, 0, tx); break;
            default: IbLinkReduce::run(p, bx, 0, 0, tx); break;
        }
    }
};

// ---------------------------------------------------------------- host-side state
// PCIe traffic per substep (BASELINE.json:5 (c)): one pinned H2D message [X 3n | U 3n | dV n | link n | origins 3L]
// (32 B per marker) and one pinned D2H message [wrench 6L doubles | band counters], both asynchronous on the
// handle's stream; the host waits on an event recorded right after the IB kernels, i.e. BEFORE the stream-collide of
// the same substep has run, so body integration and the next upload overlap the fluid kernel.
template <class Dev>
class IbState {
public:
    enum { EV_STAGE0 = 0, EV_STAGE1 = 1, EV_WRENCH = 2 };

    bool ready() const { return cap_ > 0; }
    int n_markers() const { return xchg_ ? n_total_ : n_; }
    int n_local() const { return n_; }
    int n_links() const { return nl_; }
    int band_cells() const { return band_cells_; }
    const double *wrench_ptr() const { return h_out_; }
    const double *origin_ptr() const { return h_origin_.data(); }
    void set_fused(bool on) { fused_ = on; }
    void set_reuse_static(bool on) { reuse_static_ = on; }
    void set_tile_spread(bool on) { tile_spread_ = on; }
    void clear_wrenches() { if (h_out_) std::fill(h_out_, h_out_ + 6 * size_t(maxl_), 0.0); }

    int create(Dev &dev, const FgConfig &cfg, const Lattice &L, std::string &err) {
        cap_ = cfg.max_markers;
        maxl_ = std::max(cfg.max_links, 1);
        if (const char *e = std::getenv("FG_IB_CACHE_MIN")) kCacheMinMarkers = std::atoi(e);   // tests: exercise the static-body cache on small clouds
        const long long cells = L.slot;
        band_cap_ = int(std::min<long long>(64ll * cap_, (long long)L.plane * L.nz));
        per_[0] = cfg.bc[FG_XLO] == FG_BC_PERIODIC; per_[1] = cfg.bc[FG_YLO] == FG_BC_PERIODIC; per_[2] = cfg.bc[FG_ZLO] == FG_BC_PERIODIC;
        auto A = [&](size_t bytes) { void *p = dev.alloc(bytes, err); if (p && !dev.zero(p, bytes)) { err = dev.err; p = nullptr; } return p; };
        cap_pad_ = cfg.n_ranks > 1 ? (cap_ + kPad - 1) / kPad * kPad : cap_;
        nzl_ = L.nz; nzg_ = L.nzg;
        msg_floats_ = 9 * size_t(cap_pad_) + 3 * size_t(maxl_);
        dmsg_ = (float *)A(sizeof(float) * msg_floats_);
        dbase_ = (int *)A(sizeof(int) * 3 * cap_pad_); downer_ = (int *)A(sizeof(int) * cap_pad_);
        dF_ = (float *)A(sizeof(float) * 3 * cap_pad_); dUs_ = (float *)A(sizeof(float) * 3 * cap_pad_);
        cellslot_ = (int *)A(sizeof(int) * cells);
        band_cell_ = (int *)A(sizeof(int) * band_cap_);
        band_u_ = (float *)A(sizeof(float) * 3 * band_cap_);
        bandF_ = (float *)A(sizeof(float) * 3 * band_cap_);
        band_count_ = (int *)A(2 * sizeof(int));
        rowflag_ = (uint8_t *)A(size_t(L.nz + 2) * L.ny);
        dwrench_ = (double *)A(sizeof(double) * 6 * maxl_);
        iters_ = std::max(1, int(cfg.ib_iterations));
        if (iters_ > 1 && !(dE_ = (float *)A(sizeof(float) * 3 * cap_pad_))) return FG_ENOMEM;
        rank_ = cfg.rank; n_ranks_ = cfg.n_ranks;
        if (cfg.n_ranks > 1) {
            x_bytes_ = x_off_wrench() + sizeof(double) * 2 * size_t(cfg.n_ranks) * maxl_ * 6;
            xbuf_ = (char *)A(x_bytes_);
            dxside_ = (int *)A(sizeof(int) * cap_pad_);
            if (!xbuf_ || !dxside_) return FG_ENOMEM;
        }
        for (int i = 0; i < 2; ++i) h_stage_[i] = (float *)dev.alloc_host(sizeof(float) * msg_floats_, err);
        h_out_ = (double *)dev.alloc_host(sizeof(double) * 6 * maxl_ + 2 * sizeof(int), err);
        if (!dmsg_ || !dbase_ || !downer_ || !dF_ || !dUs_ || !cellslot_ || !band_cell_ || !band_u_ || !bandF_ || !band_count_ ||
            !rowflag_ || !dwrench_ || !h_stage_[0] || !h_stage_[1] || !h_out_)
            return FG_ENOMEM;
        std::memset(h_out_, 0, sizeof(double) * 6 * maxl_ + 2 * sizeof(int));
        h_origin_.assign(3 * size_t(maxl_), 0.0);
        return FG_OK;
    }

    void destroy(Dev &dev) {
        void *ps[] = {dmsg_, dbase_, downer_, dF_, dUs_, cellslot_, band_cell_, band_u_, bandF_, band_count_, rowflag_, dwrench_, xbuf_, dxside_, node_w_, node_s_, dE_};
        xbuf_ = nullptr; dxside_ = nullptr; xchg_ = false; node_w_ = nullptr; node_s_ = nullptr; cache_valid_ = false; dE_ = nullptr;
        for (void *p : ps) dev.free(p);
        dev.free_host(h_stage_[0]); dev.free_host(h_stage_[1]); dev.free_host(h_out_);
        h_stage_[0] = h_stage_[1] = nullptr; h_out_ = nullptr;
        cap_ = 0;
    }

    // ---- exchange across slab faces
    size_t xbuf_bytes() const { return x_bytes_; }
    void *xbuf() const { return xbuf_; }
    bool exchange_on() const { return xchg_; }
    int iterations() const { return iters_; }
    // layout of the exchange buffer: [ints: 4 + kMaxRanks][ustar_in floats][wrench_all doubles]
    static size_t x_off_ustar() { return 64; }
    size_t x_off_wrench() const { return (x_off_ustar() + sizeof(float) * 2 * 2 * size_t(cap_) * 3 + 15) / 16 * 16; }
    void enable_exchange(int rank, int n_ranks, void *lo_x, void *hi_x, void *const *all_x) {
        rank_ = rank; n_ranks_ = n_ranks; xchg_ = true;
        peer_lo_x_ = static_cast<char *>(lo_x); peer_hi_x_ = static_cast<char *>(hi_x);
        for (int r = 0; r < kMaxRanks; ++r) peer_all_x_[r] = r < n_ranks ? static_cast<char *>(all_x[r]) : nullptr;
    }

    // origins: nullptr keeps the current torque reference points
    // The planes (local zz, ghost offset included) that marker stencils of this rank touch.  Cheap on purpose: one
    // floor per marker; a body that wraps around the periodic z axis simply reports "everywhere".
    void update_range(int n, const float *X) {
        int kmin = 0x7fffffff, kmax = -0x7fffffff;
        const int z0 = rank_ * nzl_;
        for (int k = 0; k < n; ++k) {
            const int k0 = int(std::floor(X[3 * k + 2])) - 1;
            if (xchg_ && slab_of(k0) != rank_ && slab_of(k0 + 3) != rank_) continue;
            kmin = std::min(kmin, k0); kmax = std::max(kmax, k0 + 3);
        }
        z_any_ = kmin <= kmax;
        z_all_ = z_any_ && per_[2] && (kmin < 0 || kmax >= nzg_);
        if (!z_any_) { zmin_ = 1; zmax_ = 0; return; }        // no stencil on this slab
        zmin_ = std::max(kmin, z0) - z0 + 1;
        zmax_ = std::min(kmax, z0 + nzl_ - 1) - z0 + 1;
    }
    // planes [za, zb) contain every cell the IB kernels of this step read or write.  In the AA pattern every storage
    // location is read and written by exactly ONE cell per step, and IbBandMoments reads precisely the locations its
    // band cells own, so the collide of any cell outside the band may run beside the IB kernels; one plane of margin
    // is kept anyway.  Snapped outwards to multiples of 8 so that a slowly moving body keeps its launch geometry
    // (and CUDA graph) for many steps.
    // the planes (local, ghost offset included) marker stencils touch, unsnapped; false: none, or wrapped around a periodic z
    bool stencil_planes(int &zmin, int &zmax) const {
        if (z_all_ || !z_any_ || zmax_ < zmin_) return false;
        zmin = zmin_; zmax = zmax_;
        return true;
    }
    // no marker stencil touches the first or the last plane of this slab (so their collide needs no IB force)
    bool boundary_planes_free(int nz) const {
        if (z_all_) return false;
        if (!z_any_ || zmax_ < zmin_) return true;
        return zmin_ > 1 && zmax_ < nz;
    }
    bool near_planes(int &za, int &zb) const {
        if (z_all_) return false;
        if (!z_any_ || zmax_ < zmin_) { za = zb = 1; return true; }   // no stencil on this slab: everything is far
        constexpr int q = 8;
        const int a = zmin_ - 1, b = zmax_ + 2;       // [a, b)
        za = a <= 1 ? 1 : 1 + (a - 1) / q * q;
        zb = 1 + (b - 1 + q - 1) / q * q;
        return true;
    }

    // origins: nullptr keeps the current torque reference points
    int set_markers(Dev &dev, int n, const float *X, const float *U, const float *dV, const int32_t *link, const double *origins,
                    int n_origins, std::string &err, bool range_known = false) {
        // validate first: a rejected call leaves the handle as it was
        if (n > cap_) { err = "more markers than FgConfig.max_markers"; return FG_EINVAL; }
        if (n_origins > maxl_) { err = "more links than FgConfig.max_links"; return FG_EINVAL; }
        int nl = n > 0 ? 1 : 0;
        if (link)
            for (int k = 0; k < n; ++k) nl = std::max(nl, link[k] + 1);
        if (nl > maxl_) { err = "link id exceeds FgConfig.max_links"; return FG_EINVAL; }
        if (!range_known) update_range(n, X);
        if (origins) {
            for (int i = 0; i < 3 * n_origins; ++i) h_origin_[i] = origins[i];
            nl_origins_ = n_origins;
        }
        // multi-rank: keep only the markers whose 4-wide stencil touches this slab, padded to a multiple of kPad so that
        // the launch geometry (and the CUDA-graph key) does not change every time a marker crosses a face
        act_.clear();
        if (xchg_) {
            for (int k = 0; k < n; ++k) {
                const int k0 = int(std::floor(X[3 * k + 2])) - 1;
                // ... or that this rank owns: a marker whose stencil lies entirely outside a non-periodic box touches no slab,
                // yet its force 2 (U_d - 0) still counts in its link's wrench (as on one rank, and in the oracle)
                if (slab_of(k0) == rank_ || slab_of(k0 + 3) == rank_ || owner_of(k0) == rank_) act_.push_back(k);
            }
        }
        const int m = xchg_ ? int((act_.size() + kPad - 1) / kPad * kPad) : n;
        if (m > cap_pad_) { err = "more markers than FgConfig.max_markers"; return FG_EINVAL; }
        // pack the message in a pinned staging buffer (double-buffered: wait for the copy issued two calls ago)
        const int sb = stage_next_;
        stage_next_ ^= 1;
        if (stage_used_[sb] && !dev.ev_sync(EV_STAGE0 + sb)) { err = dev.err; return FG_ECUDA; }
        float *msg = h_stage_[sb];
        const size_t M = size_t(m);
        if (!xchg_) {
            if (n > 0) {
                std::memcpy(msg, X, sizeof(float) * 3 * M);
                std::memcpy(msg + 3 * M, U, sizeof(float) * 3 * M);
                std::memcpy(msg + 6 * M, dV, sizeof(float) * M);
                if (link) std::memcpy(msg + 7 * M, link, sizeof(int) * M);
                else std::memset(msg + 7 * M, 0, sizeof(int) * M);
            }
        } else {
            int *lk = reinterpret_cast<int *>(msg + 7 * M), *gi = reinterpret_cast<int *>(msg + 8 * M);
            for (size_t j = 0; j < M; ++j) {
                const bool real = j < act_.size();
                const int k = real ? act_[j] : (act_.empty() ? 0 : act_[0]);
                for (int d = 0; d < 3; ++d) {
                    msg[3 * j + d] = n > 0 ? X[3 * k + d] : 0.f;
                    msg[3 * M + 3 * j + d] = n > 0 ? U[3 * k + d] : 0.f;
                }
                msg[6 * M + j] = real ? dV[k] : 0.f;
                lk[j] = real ? (link ? link[k] : 0) : -1;
                gi[j] = real ? k : -1;
            }
        }
        const size_t head = (xchg_ ? 9 : 8) * M;
        for (size_t i = 0; i < 3 * size_t(maxl_); ++i) msg[head + i] = float(h_origin_[i]);
        pending_bytes_ = sizeof(float) * (head + 3 * size_t(maxl_));
        if (!range_known) {
            // fg_set_markers (a caller that re-sends its markers every step): the copy starts NOW on the copy stream, beside
            // the collide of the previous step that is still running — the call came after fg_step returned, i.e. after the
            // last IB pass (the only reader of the device message) finished.  3.2 MB at 1e5 markers: ~0.1 ms off the step.
            if (!dev.upload_early(dmsg_, msg, pending_bytes_, EV_STAGE0 + sb, EV_WRENCH)) { err = dev.err; return FG_ECUDA; }
            stage_used_[sb] = true;
            pending_sb_ = -1;
        } else {
            // host-integrated bodies (inside fg_step, possibly inside a graph capture): the copy is queued by flush_upload()
            // at the start of the next compute_forces(), i.e. AFTER the step has launched the far-plane collide, which
            // therefore does not wait for the upload
            pending_sb_ = sb;
        }
        n_ = m;
        n_total_ = n;
        markers_dirty_ = true;
        if (xchg_) {        // host copy of the coordinates for the read-outs of markers this rank does not hold
            hX_.assign(X, X + 3 * size_t(n));
        }
        nl_ = std::max(nl, nl_origins_);
        forces_valid_ = false;
        return FG_OK;
    }

    // queue the staged marker message (if any) on the handle's stream
    bool flush_upload(Dev &dev) {
        if (pending_sb_ < 0) return true;
        const int sb = pending_sb_;
        pending_sb_ = -1;
        if (!dev.h2d_async(dmsg_, h_stage_[sb], pending_bytes_) || !dev.ev_record(EV_STAGE0 + sb)) return false;
        stage_used_[sb] = true;
        return true;
    }

    int set_link_origins(Dev &dev, int n, const double *o, std::string &err) {
        if (n > maxl_) { err = "more links than FgConfig.max_links"; return FG_EINVAL; }
        if (!flush_upload(dev)) { err = dev.err; return FG_ECUDA; }
        for (int i = 0; i < 3 * n; ++i) h_origin_[i] = o[i];
        nl_origins_ = n;
        nl_ = std::max(nl_, n);
        // rare path (prescribed markers): small synchronous upload behind the marker message
        std::vector<float> of(3 * size_t(maxl_));
        for (size_t i = 0; i < of.size(); ++i) of[i] = float(h_origin_[i]);
        if (!dev.h2d(dmsg_ + (xchg_ ? 9 : 8) * size_t(n_), of.data(), sizeof(float) * of.size())) { err = dev.err; return FG_ECUDA; }
        return FG_OK;
    }

    IbParams params(const Lattice &L, const Collision &C) const {
        IbParams p{};
        const size_t N = size_t(n_);
        p.L = L; p.C = C; p.n = n_;
        p.per_x = per_[0]; p.per_y = per_[1]; p.per_z = per_[2];
        p.X = dmsg_; p.U = dmsg_ + 3 * N; p.dV = dmsg_ + 6 * N; p.link = reinterpret_cast<const int *>(dmsg_ + 7 * N);
        p.gidx = xchg_ ? reinterpret_cast<const int *>(dmsg_ + 8 * N) : nullptr;
        p.origin = dmsg_ + (xchg_ ? 9 : 8) * N;
        p.base = dbase_; p.owner = downer_; p.Fm = dF_; p.Ustar = dUs_; p.mdfE = dE_;
        p.cellslot = cellslot_; p.band_cell = band_cell_; p.band_u = band_u_; p.bandF = bandF_;
        p.band_count = band_count_ + cur_; p.band_count_next = band_count_ + (cur_ ^ 1);
        p.band_cap = band_cap_; p.rowflag = rowflag_;
        p.wrench = dwrench_; p.n_links = nl_;
        p.node_w = node_w_; p.node_s = node_s_; p.cache_mode = cache_mode_;
        p.rank = rank_; p.n_ranks = xchg_ ? n_ranks_ : 1; p.cap = cap_; p.maxl = maxl_;
        if (xchg_) {
            p.xside = dxside_;
            p.xcounter = reinterpret_cast<const int *>(xbuf_);
            p.ustar_in = reinterpret_cast<const float *>(xbuf_ + x_off_ustar());
            p.peer_in_lo = peer_lo_x_ ? reinterpret_cast<float *>(peer_lo_x_ + x_off_ustar()) : nullptr;
            p.peer_in_hi = peer_hi_x_ ? reinterpret_cast<float *>(peer_hi_x_ + x_off_ustar()) : nullptr;
            p.wrench_all = reinterpret_cast<const double *>(xbuf_ + x_off_wrench());
            for (int r = 0; r < kMaxRanks; ++r)
                p.peer_wrench[r] = peer_all_x_[r] ? reinterpret_cast<double *>(peer_all_x_[r] + x_off_wrench()) : nullptr;
        }
        return p;
    }

    // SURVEY.md A7 (2)-(5),(7) on the device; the collide that follows reads force_view()
    int compute_forces(Dev &dev, const Lattice &L, const Collision &C, int parity, std::string &err) {
        // static bodies: the marker set was not re-sent since the last step, so its index map and band are still valid
        const bool use_fused = fused_ && !xchg_ && iters_ == 1 && dev.supports_phased();
        const bool rebuild = markers_dirty_ || !band_live_ || !reuse_static_ || use_fused;
        if (rebuild) cur_ ^= 1;                               // this step's counter; the other one still holds the old size
        // static bodies with many markers: from the second step without a re-send on, stencil weights and band slots come
        // from a per-node cache (8 B per node, allocated on first use) — filled by one step, read by the following ones
        cache_mode_ = 0;
        if (rebuild) cache_valid_ = false;
        else if (!xchg_ && !fused_ && n_ >= kCacheMinMarkers && cache_ok_) {
            if (!node_w_) {
                std::string e;
                node_w_ = static_cast<float *>(dev.alloc(sizeof(float) * size_t(cap_) * kNodes, e));
                node_s_ = static_cast<int *>(dev.alloc(sizeof(int) * size_t(cap_) * kNodes, e));
                if (!node_w_ || !node_s_) { dev.free(node_w_); dev.free(node_s_); node_w_ = nullptr; node_s_ = nullptr; cache_ok_ = false; }
            }
            if (node_w_) { cache_mode_ = cache_valid_ ? 2 : 1; cache_valid_ = true; }
        }
        IbParams p = params(L, C);
        bool ok = flush_upload(dev);
        const int nb = (n_ + kMarkersPerCta - 1) / kMarkersPerCta;
        // band kernels: a grid for the SMs (or for the bound of the band's size when that is smaller), walked grid-stride
        const auto band_grid = [&](long long bound) { return int(std::max<long long>(1, std::min<long long>((bound + 127) / 128, (long long)dev.sm_count() * IbBandMoments<0>::kCtasPerSm))); };
        if (use_fused) {
            // upper bound of any phase's work, for the launch geometry (the kernel loops grid-stride)
            const long long most = std::max<long long>((long long)nb * 128, std::min<long long>(band_cap_, (long long)kNodes * std::max(n_, n_prev_)));
            p.band_ctas = (band_cap_ + 127) / 128;                // one band cell per work item: the band kernels' own stride is never taken
            ok = ok && (parity == 0 ? dev.template launch_phased<IbFused<0>>(most, p) : dev.template launch_phased<IbFused<1>>(most, p));
        } else {
            if (rebuild) {
                if (band_live_) {
                    p.band_ctas = band_grid(std::min<long long>(band_cap_, (long long)kNodes * std::max(n_prev_, 1)));
                    ok = ok && dev.template launch<IbClearBand>(Dim3x(p.band_ctas), p);
                }
                ok = ok && dev.template launch<IbIndexMarkLaunch>(Dim3x((n_ + IbIndexMarkLaunch::kMarkers - 1) / IbIndexMarkLaunch::kMarkers), p);
            } else {
                ok = ok && dev.zero(dUs_, sizeof(float) * 3 * size_t(n_));   // IbIndexMark would have cleared the U* accumulators
            }
            p.band_ctas = band_grid(std::min<long long>(band_cap_, (long long)kNodes * n_));
            ok = ok && (parity == 0 ? dev.template launch<IbBandMoments<0>>(Dim3x(p.band_ctas), p)
                                    : dev.template launch<IbBandMoments<1>>(Dim3x(p.band_ctas), p));
            if (!xchg_) {
                const Dim3 gs = Dim3x(std::max(dev.interp_spread_blocks(n_), (6 * nl_ + 127) / 128));
                ok = ok && (tile_spread_ ? dev.template launch_block_phased<IbInterpSpreadTile>(gs, p) : dev.template launch_block_phased<IbInterpSpread>(gs, p));
                if (iters_ > 1) {
                    // multi-direct forcing: gather the spread force at every marker, then spread the corrections (IbMdfGather)
                    IbParams q = p;
                    if (q.cache_mode == 1) q.cache_mode = 2;          // the pass above has filled the per-node cache
                    for (int it = 1; it < iters_; ++it) {
                        ok = ok && dev.zero(dE_, sizeof(float) * 3 * size_t(n_));
                        ok = ok && dev.template launch<IbMdfGather>(Dim3x(nb), q);
                        ok = ok && dev.template launch<IbMdfSpread>(Dim3x(nb), q);
                    }
                }
                ok = ok && dev.template launch<IbLinkReduce>(Dim3x((n_ + 127) / 128), p);
            } else {
                // bodies across slab faces: the exchange of partial U* sits between interpolation and spreading
                int *mine = reinterpret_cast<int *>(xbuf_);
                ok = ok && dev.template launch<IbInterpolate>(Dim3x(std::max(nb, (6 * nl_ + 127) / 128)), p);
                {   // partial U* to the face neighbours, wait for theirs (counter value c+1 of this exchange)
                    int *sig[2] = {peer_lo_x_ ? reinterpret_cast<int *>(peer_lo_x_) + 2 : nullptr,    // I am its z-high neighbour
                                   peer_hi_x_ ? reinterpret_cast<int *>(peer_hi_x_) + 1 : nullptr};   // I am its z-low neighbour
                    int *wt[2] = {peer_lo_x_ ? mine + 1 : nullptr, peer_hi_x_ ? mine + 2 : nullptr};
                    ok = ok && dev.template launch<IbPushPartial>(Dim3x((n_ + 127) / 128), p);
                    ok = ok && dev.signal_counters(mine, sig, 2, false) && dev.wait_counters(mine, wt, 2);
                    ok = ok && dev.template launch<IbAddPartial>(Dim3x((n_ + 127) / 128), p);
                }
                ok = ok && dev.template launch<IbForceSpread>(Dim3x(nb), p);
                if (iters_ > 1) {
                    // Multi-direct forcing across slab faces: every pass gathers E_k over the own cells, completes it with the
                    // face neighbour's part (the U* exchange kernels, pointed at E) and spreads the correction into the own
                    // cells.  Each exchange is an epoch of its own: the counter is bumped after it, so consecutive exchanges
                    // use the two buffer sets in turn (a rank can write set s again only after its neighbour has signalled the
                    // exchange in between, i.e. after that neighbour has read set s) and the counter values keep increasing.
                    int *sig[2] = {peer_lo_x_ ? reinterpret_cast<int *>(peer_lo_x_) + 2 : nullptr, peer_hi_x_ ? reinterpret_cast<int *>(peer_hi_x_) + 1 : nullptr};
                    int *wt[2] = {peer_lo_x_ ? mine + 1 : nullptr, peer_hi_x_ ? mine + 2 : nullptr};
                    IbParams q = p;
                    q.Ustar = dE_;                                    // what IbPushPartial / IbAddPartial move
                    for (int it = 1; it < iters_; ++it) {
                        ok = ok && dev.signal_counters(mine, nullptr, 0, true);   // the previous exchange (U*, or the last pass) is complete
                        ok = ok && dev.zero(dE_, sizeof(float) * 3 * size_t(n_));
                        ok = ok && dev.template launch<IbMdfGather>(Dim3x(nb), p);
                        ok = ok && dev.template launch<IbPushPartial>(Dim3x((n_ + 127) / 128), q);
                        ok = ok && dev.signal_counters(mine, sig, 2, false) && dev.wait_counters(mine, wt, 2);
                        ok = ok && dev.template launch<IbAddPartial>(Dim3x((n_ + 127) / 128), q);
                        ok = ok && dev.template launch<IbMdfSpread>(Dim3x(nb), p);
                    }
                    ok = ok && dev.signal_counters(mine, nullptr, 0, true);       // the wrench exchange below is the next epoch
                }
                ok = ok && dev.template launch<IbLinkReduce>(Dim3x((n_ + 127) / 128), p);
                {   // owner-reduced link wrenches to every rank, summed in rank order (bit-identical totals everywhere)
                    int *sig[kMaxRanks], *wt[kMaxRanks];
                    for (int r = 0; r < kMaxRanks; ++r) {
                        sig[r] = peer_all_x_[r] ? reinterpret_cast<int *>(peer_all_x_[r]) + 4 + rank_ : nullptr;
                        wt[r] = peer_all_x_[r] ? mine + 4 + r : nullptr;
                    }
                    const int g6 = (6 * nl_ + 127) / 128;
                    ok = ok && dev.template launch<IbPushWrench>(Dim3x(g6), p);
                    ok = ok && dev.signal_counters(mine, sig, kMaxRanks, false) && dev.wait_counters(mine, wt, kMaxRanks);
                    ok = ok && dev.template launch<IbSumWrench>(Dim3x(g6), p);
                    ok = ok && dev.signal_counters(mine, nullptr, 0, true);   // bump my counter: this exchange is complete
                }
            }
        }
        markers_dirty_ = false;
        // results the host needs, queued right behind the IB kernels (not behind the collide that follows)
        ok = ok && dev.d2h_async(h_out_, dwrench_, sizeof(double) * 6 * maxl_) &&
             dev.d2h_async(reinterpret_cast<char *>(h_out_) + sizeof(double) * 6 * maxl_, band_count_, 2 * sizeof(int)) &&
             dev.ev_record(EV_WRENCH);
        if (!ok) { err = dev.err; return FG_ECUDA; }
        band_live_ = true; n_prev_ = n_; forces_valid_ = true; wrench_fetched_ = false;
        return FG_OK;
    }

    // what changes the IB launches of the next substep (CUDA-graph cache key, sim.hpp substep_key)
    void graph_key(uint64_t (&w)[3]) const {
        w[0] = uint64_t(cur_) | (uint64_t(stage_next_) << 1) | (uint64_t(band_live_) << 2) | (uint64_t(n_ > 0) << 3) | (uint64_t(fused_) << 4) |
               (uint64_t(markers_dirty_ || !reuse_static_) << 5) | (uint64_t(xchg_) << 6) | (uint64_t(next_cache_mode()) << 7);
        w[1] = uint64_t(uint32_t(n_)) | (uint64_t(uint32_t(n_prev_)) << 32);
        w[2] = uint64_t(uint32_t(nl_));
    }
    // the cache mode the NEXT compute_forces will pick (part of the CUDA-graph key: it selects launch arguments)
    int next_cache_mode() const {
        const bool rebuild = markers_dirty_ || !band_live_ || !reuse_static_;
        if (rebuild || fused_ || xchg_ || n_ < kCacheMinMarkers || !cache_ok_) return 0;
        return cache_valid_ ? 2 : 1;
    }
    ForceField force_view() const { return ForceField{cellslot_, bandF_, band_cap_, rowflag_}; }

    int fetch_wrenches(Dev &dev, std::string &err) {
        if (!forces_valid_ || wrench_fetched_) return FG_OK;
        if (!dev.ev_sync(EV_WRENCH)) { err = dev.err; return FG_ECUDA; }
        const int *cnts = reinterpret_cast<const int *>(reinterpret_cast<const char *>(h_out_) + sizeof(double) * 6 * maxl_);
        if (cnts[cur_] > band_cap_) { err = "IB band overflow"; return FG_ENOMEM; }
        band_cells_ = cnts[cur_];
        wrench_fetched_ = true;
        return FG_OK;
    }

    // read-outs for tests and observations
    // Per-marker read-outs and link wrenches describe the LAST STEP's immersed-boundary pass.  Between fg_set_markers and
    // the next fg_step there is none for the new marker set: everything reads as zero (the oracle does the same).
    bool forces_valid() const { return forces_valid_; }
    int get_index_map(Dev &dev, int32_t *base3, int32_t *owner, std::string &err) {
        if (n_markers() == 0) return FG_OK;
        if (!forces_valid_) {
            std::fill(base3, base3 + 3 * size_t(n_markers()), 0);
            std::fill(owner, owner + size_t(n_markers()), 0);
            return FG_OK;
        }
        if (!xchg_) {
            if (!dev.sync() || !dev.d2h(base3, dbase_, sizeof(int) * 3 * n_) || !dev.d2h(owner, downer_, sizeof(int) * n_)) { err = dev.err; return FG_ECUDA; }
            return FG_OK;
        }
        // culled: the device holds this rank's markers only; the map is pure integer arithmetic on the fp32 coordinates
        // (same formula as IbIndexMark), so the others are filled in on the host and the held ones come from the device
        for (int k = 0; k < n_total_; ++k) {
            for (int d = 0; d < 3; ++d) base3[3 * k + d] = int(std::floor(hX_[3 * k + d])) - 1;
            owner[k] = owner_of(base3[3 * k + 2]);
        }
        std::vector<int> b(3 * size_t(n_)), o(n_);
        if (!dev.sync() || !dev.d2h(b.data(), dbase_, sizeof(int) * 3 * n_) || !dev.d2h(o.data(), downer_, sizeof(int) * n_)) { err = dev.err; return FG_ECUDA; }
        for (size_t j = 0; j < act_.size(); ++j) {
            for (int d = 0; d < 3; ++d) base3[3 * act_[j] + d] = b[3 * j + d];
            owner[act_[j]] = o[j];
        }
        return FG_OK;
    }
    int get_marker_array(Dev &dev, bool forces, float *out, std::string &err) {
        if (n_markers() == 0) return FG_OK;
        if (!forces_valid_) { std::fill(out, out + 3 * size_t(n_markers()), 0.f); return FG_OK; }
        if (!xchg_) {
            if (!dev.sync() || !dev.d2h(out, forces ? dF_ : dUs_, sizeof(float) * 3 * n_)) { err = dev.err; return FG_ECUDA; }
            return FG_OK;
        }
        std::vector<float> loc(3 * size_t(n_));
        if (!dev.sync() || !dev.d2h(loc.data(), forces ? dF_ : dUs_, sizeof(float) * 3 * n_)) { err = dev.err; return FG_ECUDA; }
        std::fill(out, out + 3 * size_t(n_total_), 0.f);       // markers this rank does not hold read as zero
        for (size_t j = 0; j < act_.size(); ++j)
            for (int d = 0; d < 3; ++d) out[3 * act_[j] + d] = loc[3 * j + d];
        return FG_OK;
    }
    int get_force_field(Dev &dev, const Lattice &L, float *F, std::string &err) {
        const size_t nloc = size_t(L.plane) * L.nz;
        std::fill(F, F + 3 * nloc, 0.f);
        if (!band_live_) return FG_OK;
        int cnt = 0;
        if (!dev.sync() || !dev.d2h(&cnt, band_count_ + cur_, sizeof(int))) { err = dev.err; return FG_ECUDA; }
        cnt = std::min(cnt, band_cap_);
        std::vector<int> cells(cnt);
        std::vector<float> bf(3 * size_t(band_cap_));
        if (cnt == 0) return FG_OK;
        if (!dev.d2h(cells.data(), band_cell_, sizeof(int) * cnt) || !dev.d2h(bf.data(), bandF_, sizeof(float) * bf.size())) { err = dev.err; return FG_ECUDA; }
        for (int i = 0; i < cnt; ++i) {
            const long long l = (long long)cells[i] - L.plane;   // drop the ghost plane offset
            if (l < 0 || l >= (long long)nloc) continue;
            F[l] = bf[i]; F[nloc + l] = bf[band_cap_ + i]; F[2 * nloc + l] = bf[2 * size_t(band_cap_) + i];
        }
        return FG_OK;
    }
    int get_markers(Dev &dev, float *X, float *U, int32_t *link, int cap, std::string &err) {
        if (xchg_) {
            if (cap > 0) { err = "fg_get_markers: not available with culled markers (fg_peer_connect_all); count only"; return cap >= n_total_ && !X && !U && !link ? n_total_ : FG_ENOTSUP; }
            return n_total_;
        }
        const int n = std::min(cap, n_);
        const size_t N = size_t(n_);
        bool ok = flush_upload(dev) && dev.sync();
        if (n > 0 && X) ok = ok && dev.d2h(X, dmsg_, sizeof(float) * 3 * n);
        if (n > 0 && U) ok = ok && dev.d2h(U, dmsg_ + 3 * N, sizeof(float) * 3 * n);
        if (n > 0 && link) ok = ok && dev.d2h(link, dmsg_ + 7 * N, sizeof(int) * n);
        if (!ok) { err = dev.err; return FG_ECUDA; }
        return n_;
    }

private:
    static Dim3 Dim3x(int x) { Dim3 d; d.x = std::max(x, 1); return d; }
    int cap_ = 0, maxl_ = 1, n_ = 0, nl_ = 0, nl_origins_ = 0, n_prev_ = 0, band_cap_ = 0, band_cells_ = 0;
    static constexpr int kPad = 512;
    int slab_plane(int z) const {
        if (z >= 0 && z < nzg_) return z;
        if (!per_[2]) return -1;
        z %= nzg_;
        return z < 0 ? z + nzg_ : z;
    }
    int slab_of(int z) const { const int zg = slab_plane(z); return zg < 0 ? -1 : zg / nzl_; }
    int owner_of(int k0) const {        // IbIndexMark's owner: the slab of the stencil's second plane, clamped into the box
        int kc = slab_plane(k0 + 1);
        if (kc < 0) kc = k0 + 1 < 0 ? 0 : nzg_ - 1;
        return kc / nzl_;
    }
    int cap_pad_ = 0, n_total_ = 0, nzl_ = 1, nzg_ = 1;
    std::vector<int> act_;
    std::vector<float> hX_;
    int cur_ = 0, stage_next_ = 0;
    int rank_ = 0, n_ranks_ = 1;
    bool xchg_ = false;
    size_t x_bytes_ = 0;
    char *xbuf_ = nullptr, *peer_lo_x_ = nullptr, *peer_hi_x_ = nullptr, *peer_all_x_[kMaxRanks] = {};
    int *dxside_ = nullptr;
    size_t msg_floats_ = 0;
    int per_[3] = {1, 1, 1};
    bool band_live_ = false, forces_valid_ = false, wrench_fetched_ = true, fused_ = false;
    bool markers_dirty_ = true, reuse_static_ = true, tile_spread_ = false;
    bool z_any_ = false, z_all_ = true;
    int zmin_ = 1, zmax_ = 0;
    bool stage_used_[2] = {false, false};
    int pending_sb_ = -1;           // staging buffer holding a marker message that has not been queued yet
    size_t pending_bytes_ = 0;
    float *dmsg_ = nullptr, *dF_ = nullptr, *dUs_ = nullptr, *band_u_ = nullptr, *bandF_ = nullptr;
    float *dE_ = nullptr;          // multi-direct forcing (FgConfig.ib_iterations > 1): gathered force per marker
    int iters_ = 1;
    int *dbase_ = nullptr, *downer_ = nullptr, *cellslot_ = nullptr, *band_cell_ = nullptr, *band_count_ = nullptr;
    uint8_t *rowflag_ = nullptr;
    double *dwrench_ = nullptr;
    float *node_w_ = nullptr;                  // per-node cache of static bodies (lazily allocated): weight ...
    int *node_s_ = nullptr;                    // ... and band slot
    int cache_mode_ = 0;
    bool cache_valid_ = false, cache_ok_ = true;
    int kCacheMinMarkers = 8192;               // below this the IB kernels are launch-latency bound and the cache buys nothing
    float *h_stage_[2] = {nullptr, nullptr};   // pinned
    double *h_out_ = nullptr;                  // pinned: [6 maxl doubles][2 ints]
    std::vector<double> h_origin_;
};

}  // namespace fg
