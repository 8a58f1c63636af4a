// lbm_core.cuh — device code of the fused D3Q19 stream-collide path.
//
// Path and boundary: /root/reference/README.md:2-3 names the coupled agent-fluid step; there is no
// reference source, the scheme is BASELINE.json:5 (a): single-lattice in-place streaming
// (AA pattern), SoA fp32 populations, MRT moment transform in registers, bounce-back and Guo
// forcing fused into the collide.
//
// Data layout in HBM: one array f[19][nz+2][ny][nx] of fp32 (x fastest).  Planes zz=0 and zz=nz+1
// are ghost planes: EVERY z-face condition (periodic wrap, slab neighbour, inlet, outlet) is a
// plane operation on 5 populations between steps (ZFaceOp below), so the main kernel never
// wraps in z.  Populations are stored SHIFTED, h_i = f_i - w_i (SURVEY.md §7 "DDF shifting"), which
// is what keeps fp32 within 1e-5 of the fp64 oracle.
//
// AA pattern.  Even step: read slot i at x, collide, write slot opp(i) at x.
// Odd step: read f_i from slot opp(i) at x-c_i and f_opp(i) from slot i at x+c_i, collide, write
// f*_i to slot i at x+c_i and f*_opp(i) to slot opp(i) at x-c_i.  Each thread reads and writes the
// same 19 addresses, so the update is in place and race-free.  Half-way bounce-back redirects a
// blocked link to the cell's own opposite slot on both the read and the write side.
//
// Everything here is __host__ __device__: tests/emu compiles the same functions with g++ and
// loops over the grid on the CPU (test infrastructure only; the product library is CUDA-only).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#define FG_UNROLL _Pragma("unroll")
#else
#define FG_HD inline
#define FG_UNROLL
#endif

// Population storage type, fixed at compile time.  The product library stores fp32 (BASELINE.json:5); the same sources
// built with -DFG_POP16 give libfishgym_cuda_f16.so: 16-bit storage of the SHIFTED populations, scaled by 2^12 so that
// near-rest flows stay out of the subnormal range, fp32 arithmetic — 76 B per cell update instead of 152.
#if defined(FG_POP16)
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif
#endif

namespace fg {

#if defined(FG_POP16)
#if defined(__CUDACC__)
typedef __half pop_t;
FG_HD float pop_dec(pop_t v) { return __half2float(v) * (1.0f / 4096.0f); }
FG_HD pop_t pop_enc(float v) { return __float2half_rn(v * 4096.0f); }
#else
typedef _Float16 pop_t;
FG_HD float pop_dec(pop_t v) { return float(v) * (1.0f / 4096.0f); }
FG_HD pop_t pop_enc(float v) { return pop_t(v * 4096.0f); }
#endif
FG_HD float pop_ld(const pop_t *p) { return pop_dec(*p); }
FG_HD void pop_st(pop_t *p, float v) { *p = pop_enc(v); }
#define FG_POP_NAME "f16"
#else
typedef float pop_t;
FG_HD float pop_ld(const pop_t *p) {
#if defined(__CUDA_ARCH__) && defined(FG_LDCS)
    return __ldcs(p);       // experiment: streaming (evict-first) loads
#elif defined(__CUDA_ARCH__) && defined(FG_LDCG)
    return __ldcg(p);       // experiment: L2 only
#else
    return *p;
#endif
}
FG_HD void pop_st(pop_t *p, float v) {
#if defined(__CUDA_ARCH__) && defined(FG_STCS)
    __stcs(p, v);           // experiment: streaming stores
#elif defined(__CUDA_ARCH__) && defined(FG_STCG)
    __stcg(p, v);
#else
    *p = v;
#endif
}
FG_HD float pop_dec(pop_t v) { return v; }
FG_HD pop_t pop_enc(float v) { return v; }
#define FG_POP_NAME "f32"
#endif
constexpr int kPopBytes = int(sizeof(pop_t));

constexpr int Q = 19;
// SURVEY.md A1 (d'Humieres ordering)
constexpr int CXT[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
constexpr int CYT[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
constexpr int CZT[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
constexpr int OPPT[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
constexpr float W0 = 1.0f / 3.0f, W1 = 1.0f / 18.0f, W2 = 1.0f / 36.0f;
constexpr double WD[Q] = {1.0 / 3, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18,
                          1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36,
                          1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
// the five populations crossing a z face
constexpr int ZPT[5] = {5, 11, 12, 15, 16};   // c_z = +1
constexpr int ZMT[5] = {6, 13, 14, 17, 18};   // c_z = -1

// run-time indexed copies (namespace-scope constexpr arrays are not addressable in device code)
FG_HD int cxr(int i) { constexpr int t[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0}; return t[i]; }
FG_HD int cyr(int i) { constexpr int t[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return t[i]; }
FG_HD int czr(int i) { constexpr int t[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1}; return t[i]; }
FG_HD int oppr(int i) { constexpr int t[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15}; return t[i]; }
FG_HD int zpr(int k) { constexpr int t[5] = {5, 11, 12, 15, 16}; return t[k]; }
FG_HD int zmr(int k) { constexpr int t[5] = {6, 13, 14, 17, 18}; return t[k]; }

struct Dim3 { int x = 1, y = 1, z = 1; };
struct GraphKey {            // everything that selects the kernels of a substep or changes their arguments
    uint64_t w[4];
    uint64_t &operator[](int i) { return w[i]; }
    bool operator<(const GraphKey &o) const {
        for (int i = 0; i < 4; ++i) if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
};   // launch grid (blocks); block size is the kernel's kThreads

enum : int { BC_PERIODIC = 0, BC_WALL = 1, BC_INLET = 2, BC_OUTLET = 3, BC_PEER = 4 };
enum : int { F_XLO = 0, F_XHI = 1, F_YLO = 2, F_YHI = 3, F_ZLO = 4, F_ZHI = 5 };

struct Lattice {
    pop_t *f;             // slot q lives at f + q*slot
    long long slot;       // (nz+2)*ny*nx
    int nx, ny, nz;       // local slab (interior planes zz = 1..nz)
    int plane;            // nx*ny
    int z0, nzg;          // first owned global plane, global height
    int wall_x, wall_y;   // 1: both faces of the axis are walls; 0: periodic
    int bc_zlo, bc_zhi;   // GLOBAL z-face conditions (BC_*), relevant where the slab touches them
    const uint8_t *solid; // [(nz+2)*plane] interior obstacles (incl. ghost planes) or nullptr
};

struct Collision {
    float omega;          // 1/tau
    float rate[Q];        // MRT relaxation rates (conserved rows 0)
    float g[3];           // uniform body force density
    float wallterm[6][Q]; // 6 w_i c_i.u_wall(face): moving-wall bounce-back correction
    float heq_in[Q];      // shifted inlet equilibrium h_eq,i(inlet_rho, inlet_u)
};

// Eulerian IB force, band-sparse: cellslot[idx] = 1 + position in the band arrays, 0 = no force
struct ForceField {
    const int *cellslot;   // [(nz+2)*plane] or nullptr (no immersed boundary)
    const float *bandF;    // [3][band_cap]
    int band_cap;
    const uint8_t *rowflag; // [(nz+2)*ny] rows that contain band cells (lets whole rows skip the lookup)
};

struct StepParams {
    Lattice L;
    Collision C;
    ForceField F;
    int zz_begin, zz_stride;  // planes this launch covers: zz = zz_begin + blockIdx.z * zz_stride ...
    int zz_skip_begin, zz_skip_len;   // ... and from zz_skip_begin on, zz_skip_len planes further up (a launch with a hole)
    int zz_flip;              // >= 0: blockIdx.z is replaced by zz_flip - blockIdx.z, i.e. the planes are swept downwards
    int y0, ystride;       // rows this launch covers: y = y0 + blockIdx.y * ystride
    // two-cell kernels on rows narrower than a CTA's 256 cells (nx = 128, 64): a CTA covers 256 / nx consecutive rows of the
    // launch (the NARROW instantiations), row = (blockIdx.y << rows_per_cta_log2) + (cell offset >> nx_log2)
    int rows_per_cta_log2, nx_log2, rows;
    long long kz[Q][3];    // byte offset of (slot S, plane z-1 / z / z+1) from a cell of plane z: kPopBytes*(S*slot + dz*plane)
};

// ---------------------------------------------------------------- collide (registers only)
// shifted populations: rho = 1 + sum h, j = sum c h, u = (j + F/2)/rho  

// 3 w_i for the momentum-carrying (odd) part of the equilibrium and of the Guo source.  In fp32, 18*fl(1/18) != 1:
// with rounded weights the equilibrium carries (1+delta) j and BGK drifts momentum by omega*delta per step — a
// SYSTEMATIC error that reaches 1e-5 within ~1000 steps.  These two constants satisfy 2*A1 + 8*A2 == 1 exactly in
// fp32 (A2 = fl(1/12), A1 = 1/2 - 4*A2 is representable), and the odd part is built from j itself (no division),
// so sum_i c_i h*_i == j + F up to unbiased rounding.
constexpr float A2 = 1.0f / 12.0f;
constexpr float A1 = 0.5f - 4.0f * A2;
static_assert(2.0f * A1 + 8.0f * A2 == 1.0f, "odd-part weights must sum to exactly 1");

template <int I> struct Dir {
    static constexpr int cx = CXT[I], cy = CYT[I], cz = CZT[I], opp = OPPT[I];
    static constexpr float w = I == 0 ? W0 : (I < 7 ? W1 : W2);
    static constexpr float a = I < 7 ? A1 : A2;   // 3 w_i, see above
};

template <int I>
FG_HD void bgk_pair(float (&h)[Q], float dr, float rho, float ux, float uy, float uz, float uu15, float Fx, float Fy,
                    float Fz, float uF3, float om, float kf, float jx, float jy, float jz) {
    using D = Dir<I>;
    constexpr int J = D::opp;
    const float cu = D::cx * ux + D::cy * uy + D::cz * uz;
    const float cF = D::cx * Fx + D::cy * Fy + D::cz * Fz;
    const float sym = D::w * (dr + rho * (4.5f * cu * cu - uu15));
    const float asym = D::a * (D::cx * jx + D::cy * jy + D::cz * jz);   // 3 w rho c.u with rho u = j + F/2 taken as is
    const float psym = D::w * (9.0f * cu * cF - uF3);
    const float pasym = D::a * cF;
    h[I] = h[I] - om * (h[I] - (sym + asym)) + kf * (psym + pasym);
    h[J] = h[J] - om * (h[J] - (sym - asym)) + kf * (psym - pasym);
}

FG_HD void collide_bgk(float (&h)[Q], float Fx, float Fy, float Fz, const Collision &c) {
    const float sx = h[1] + h[2], sy = h[3] + h[4], sz = h[5] + h[6];
    const float sxy = (h[7] + h[8]) + (h[9] + h[10]), sxz = (h[11] + h[12]) + (h[13] + h[14]),
                syz = (h[15] + h[16]) + (h[17] + h[18]);
    const float dr = h[0] + ((sx + sy) + sz) + ((sxy + sxz) + syz);
    const float jx = (h[1] - h[2]) + ((h[7] - h[8]) + (h[9] - h[10])) + ((h[11] - h[12]) + (h[13] - h[14]));
    const float jy = (h[3] - h[4]) + ((h[7] + h[8]) - (h[9] + h[10])) + ((h[15] - h[16]) + (h[17] - h[18]));
    const float jz = (h[5] - h[6]) + ((h[11] + h[12]) - (h[13] + h[14])) + ((h[15] + h[16]) - (h[17] + h[18]));
    const float rho = 1.0f + dr, inv = 1.0f / rho;
    const float ex = jx + 0.5f * Fx, ey = jy + 0.5f * Fy, ez = jz + 0.5f * Fz;   // rho u
    const float ux = ex * inv, uy = ey * inv, uz = ez * inv;
    const float uu15 = 1.5f * (ux * ux + uy * uy + uz * uz);
    const float uF3 = 3.0f * (ux * Fx + uy * Fy + uz * Fz);
    const float om = c.omega, kf = 1.0f - 0.5f * om;
    h[0] = h[0] - om * (h[0] - W0 * (dr - rho * uu15)) - kf * W0 * uF3;
    bgk_pair<1>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<3>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<5>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<7>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<8>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<11>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<12>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<15>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
    bgk_pair<16>(h, dr, rho, ux, uy, uz, uu15, Fx, Fy, Fz, uF3, om, kf, ex, ey, ez);
}

// MRT in the d'Humieres basis, closed forms for M h, m_eq, M*source and M^-1.
// Works on shifted moments: M w = (1,-11,3,0,...) only offsets rho, e, eps, which cancel in m - m_eq.
FG_HD void collide_mrt(float (&h)[Q], float Fx, float Fy, float Fz, const Collision &c) {
    const float sx = h[1] + h[2], sy = h[3] + h[4], sz = h[5] + h[6];
    const float sxy = (h[7] + h[8]) + (h[9] + h[10]), sxz = (h[11] + h[12]) + (h[13] + h[14]),
                syz = (h[15] + h[16]) + (h[17] + h[18]);
    const float saxis = (sx + sy) + sz, sedge = (sxy + sxz) + syz;
    const float ax = h[1] - h[2], ay = h[3] - h[4], az = h[5] - h[6];
    const float exy_x = (h[7] - h[8]) + (h[9] - h[10]);    // sum c_x over xy edges
    const float exz_x = (h[11] - h[12]) + (h[13] - h[14]); // sum c_x over xz edges
    const float exy_y = (h[7] + h[8]) - (h[9] + h[10]);
    const float eyz_y = (h[15] - h[16]) + (h[17] - h[18]);
    const float exz_z = (h[11] + h[12]) - (h[13] + h[14]);
    const float eyz_z = (h[15] + h[16]) - (h[17] + h[18]);
    // moments of h
    const float dr = h[0] + saxis + sedge;
    const float m_e = -30.0f * h[0] - 11.0f * saxis + 8.0f * sedge;
    const float m_eps = 12.0f * h[0] - 4.0f * saxis + sedge;
    const float ex = exy_x + exz_x, ey = exy_y + eyz_y, ez = exz_z + eyz_z;
    const float jx = ax + ex, jy = ay + ey, jz = az + ez;
    const float qx = ex - 4.0f * ax, qy = ey - 4.0f * ay, qz = ez - 4.0f * az;
    const float pa = 2.0f * sx - sy - sz, pe = sxy + sxz - 2.0f * syz;
    const float m_pxx = pa + pe, m_pixx = pe - 2.0f * pa;
    const float wa = sy - sz, we = sxy - sxz;
    const float m_pww = wa + we, m_piww = we - 2.0f * wa;
    const float m_pxy = (h[7] - h[8]) - (h[9] - h[10]);
    const float m_pyz = (h[15] - h[16]) - (h[17] - h[18]);
    const float m_pxz = (h[11] - h[12]) - (h[13] - h[14]);
    const float m_mx = exy_x - exz_x, m_my = eyz_y - exy_y, m_mz = exz_z - eyz_z;
    // macroscopic
    const float rho = 1.0f + dr, inv = 1.0f / rho;
    const float rux = jx + 0.5f * Fx, ruy = jy + 0.5f * Fy, ruz = jz + 0.5f * Fz;   // rho u, no division
    const float ux = rux * inv, uy = ruy * inv, uz = ruz * inv;
    const float ruxx = rux * ux, ruyy = ruy * uy, ruzz = ruz * uz;
    const float ruu = ruxx + ruyy + ruzz;
    const float uF = ux * Fx + uy * Fy + uz * Fz;
    const float fxx = 2.0f * ux * Fx - uy * Fy - uz * Fz;   // source of 3p_xx / 2
    const float fww = uy * Fy - uz * Fz;
    const float *s = c.rate;
    // d_k = [ -s_k (m_k - m_eq,k) + (1 - s_k/2) Psi_k ] / |row_k|^2
    const float d1 = (-s[1] * (m_e - (-11.0f * dr + 19.0f * ruu)) + (1.0f - 0.5f * s[1]) * 38.0f * uF) * (1.0f / 2394.0f);
    const float d2 = (-s[2] * (m_eps - (3.0f * dr - 5.5f * ruu)) - (1.0f - 0.5f * s[2]) * 11.0f * uF) * (1.0f / 252.0f);
    const float d3 = Fx * 0.1f, d5 = Fy * 0.1f, d7 = Fz * 0.1f;
    const float k23 = 2.0f / 3.0f;
    const float d4 = (-s[4] * (qx + k23 * rux) - (1.0f - 0.5f * s[4]) * k23 * Fx) * 0.025f;
    const float d6 = (-s[6] * (qy + k23 * ruy) - (1.0f - 0.5f * s[6]) * k23 * Fy) * 0.025f;
    const float d8 = (-s[8] * (qz + k23 * ruz) - (1.0f - 0.5f * s[8]) * k23 * Fz) * 0.025f;
    const float pxx_eq = 2.0f * ruxx - ruyy - ruzz, pww_eq = ruyy - ruzz;
    const float d9 = (-s[9] * (m_pxx - pxx_eq) + (1.0f - 0.5f * s[9]) * 2.0f * fxx) * (1.0f / 36.0f);
    const float d10 = (-s[10] * (m_pixx + 0.5f * pxx_eq) - (1.0f - 0.5f * s[10]) * fxx) * (1.0f / 72.0f);
    const float d11 = (-s[11] * (m_pww - pww_eq) + (1.0f - 0.5f * s[11]) * 2.0f * fww) * (1.0f / 12.0f);
    const float d12 = (-s[12] * (m_piww + 0.5f * pww_eq) - (1.0f - 0.5f * s[12]) * fww) * (1.0f / 24.0f);
    const float d13 = (-s[13] * (m_pxy - rux * uy) + (1.0f - 0.5f * s[13]) * (ux * Fy + uy * Fx)) * 0.25f;
    const float d14 = (-s[14] * (m_pyz - ruy * uz) + (1.0f - 0.5f * s[14]) * (uy * Fz + uz * Fy)) * 0.25f;
    const float d15 = (-s[15] * (m_pxz - rux * uz) + (1.0f - 0.5f * s[15]) * (ux * Fz + uz * Fx)) * 0.25f;
    const float d16 = -s[16] * m_mx * 0.125f, d17 = -s[17] * m_my * 0.125f, d18 = -s[18] * m_mz * 0.125f;
    // h* = h + M^T d
    h[0] += -30.0f * d1 + 12.0f * d2;
    const float ca = -11.0f * d1 - 4.0f * d2;
    const float cax = ca + 2.0f * d9 - 4.0f * d10;
    const float cay = ca - d9 + 2.0f * d10 + d11 - 2.0f * d12;
    const float caz = ca - d9 + 2.0f * d10 - d11 + 2.0f * d12;
    const float vx = d3 - 4.0f * d4, vy = d5 - 4.0f * d6, vz = d7 - 4.0f * d8;
    h[1] += cax + vx; h[2] += cax - vx;
    h[3] += cay + vy; h[4] += cay - vy;
    h[5] += caz + vz; h[6] += caz - vz;
    const float ce = 8.0f * d1 + d2;
    const float cxy = ce + d9 + d10 + d11 + d12, cxz = ce + d9 + d10 - d11 - d12, cyz = ce - 2.0f * (d9 + d10);
    const float tx = d3 + d4, ty = d5 + d6, tz = d7 + d8;
    {   // xy edges 7:(+,+) 8:(-,+) 9:(+,-) 10:(-,-)
        const float a = tx + d16, b = ty - d17;
        h[7] += cxy + a + b + d13; h[8] += cxy - a + b - d13;
        h[9] += cxy + a - b - d13; h[10] += cxy - a - b + d13;
    }
    {   // xz edges 11:(+,+) 12:(-,+) 13:(+,-) 14:(-,-)
        const float a = tx - d16, b = tz + d18;
        h[11] += cxz + a + b + d15; h[12] += cxz - a + b - d15;
        h[13] += cxz + a - b - d15; h[14] += cxz - a - b + d15;
    }
    {   // yz edges 15:(+,+) 16:(-,+) 17:(+,-) 18:(-,-)
        const float a = ty + d17, b = tz - d18;
        h[15] += cyz + a + b + d14; h[16] += cyz - a + b - d14;
        h[17] += cyz + a - b - d14; h[18] += cyz - a - b + d14;
    }
}

// rho - 1 and momentum of shifted populations (no force)
FG_HD void moments(const float (&h)[Q], float &dr, float &jx, float &jy, float &jz) {
    const float sx = h[1] + h[2], sy = h[3] + h[4], sz = h[5] + h[6];
    const float sxy = (h[7] + h[8]) + (h[9] + h[10]), sxz = (h[11] + h[12]) + (h[13] + h[14]),
                syz = (h[15] + h[16]) + (h[17] + h[18]);
    dr = h[0] + ((sx + sy) + sz) + ((sxy + sxz) + syz);
    jx = (h[1] - h[2]) + ((h[7] - h[8]) + (h[9] - h[10])) + ((h[11] - h[12]) + (h[13] - h[14]));
    jy = (h[3] - h[4]) + ((h[7] + h[8]) - (h[9] + h[10])) + ((h[15] - h[16]) + (h[17] - h[18]));
    jz = (h[5] - h[6]) + ((h[11] + h[12]) - (h[13] + h[14])) + ((h[15] + h[16]) - (h[17] + h[18]));
}

// ---------------------------------------------------------------- neighbourhood of one cell
struct Nbr {
    int dxm, dxp, dym, dyp, dzm, dzp;       // index offsets to x-1, x+1, y-1, y+1, z-1, z+1 (periodic wrap in x,y)
    bool wxm, wxp, wym, wyp, wzm, wzp;      // the link towards that side crosses a wall face

    template <int CX, int CY, int CZ> FG_HD int off() const {
        return (CX > 0 ? dxp : (CX < 0 ? dxm : 0)) + (CY > 0 ? dyp : (CY < 0 ? dym : 0)) + (CZ > 0 ? dzp : (CZ < 0 ? dzm : 0));
    }
    // first wall face crossed by a link in direction (CX,CY,CZ), priority x, y, z (oracle: pull()), -1 if none
    template <int CX, int CY, int CZ> FG_HD int wall() const {
        if (CX > 0 && wxp) return F_XHI;
        if (CX < 0 && wxm) return F_XLO;
        if (CY > 0 && wyp) return F_YHI;
        if (CY < 0 && wym) return F_YLO;
        if (CZ > 0 && wzp) return F_ZHI;
        if (CZ < 0 && wzm) return F_ZLO;
        return -1;
    }
};

FG_HD Nbr make_nbr(const Lattice &L, int x, int y, int zz) {
    Nbr n;
    n.dxm = x == 0 ? L.nx - 1 : -1;
    n.dxp = x == L.nx - 1 ? -(L.nx - 1) : 1;
    n.dym = y == 0 ? (L.ny - 1) * L.nx : -L.nx;
    n.dyp = y == L.ny - 1 ? -(L.ny - 1) * L.nx : L.nx;
    n.dzm = -L.plane;
    n.dzp = L.plane;
    const int zg = L.z0 + zz - 1;
    n.wxm = L.wall_x && x == 0;
    n.wxp = L.wall_x && x == L.nx - 1;
    n.wym = L.wall_y && y == 0;
    n.wyp = L.wall_y && y == L.ny - 1;
    n.wzm = L.bc_zlo == BC_WALL && zg == 0;
    n.wzp = L.bc_zhi == BC_WALL && zg == L.nzg - 1;
    return n;
}

// ---------------------------------------------------------------- AA-pattern loads / stores
// Load the populations ARRIVING at the cell (f_i(x,t)) for the given storage parity.
template <int I, bool CHECK>
FG_HD void odd_load_pair(float (&h)[Q], const Lattice &L, const Collision &C, const Nbr &nb, long long idx) {
    using D = Dir<I>;
    constexpr int J = D::opp;
    const pop_t *fI = L.f + I * L.slot, *fJ = L.f + J * L.slot;
    const int om = nb.off<-D::cx, -D::cy, -D::cz>();   // towards x - c_I
    const int op = nb.off<D::cx, D::cy, D::cz>();      // towards x + c_I
    if (CHECK) {
        const int wm = nb.wall<-D::cx, -D::cy, -D::cz>();
        const int wp = nb.wall<D::cx, D::cy, D::cz>();
        const bool bm = wm >= 0 || (L.solid && L.solid[idx + om]);
        const bool bp = wp >= 0 || (L.solid && L.solid[idx + op]);
        h[I] = bm ? pop_ld(fI + idx) + (wm >= 0 ? C.wallterm[wm][I] : 0.0f) : pop_ld(fJ + idx + om);
        h[J] = bp ? pop_ld(fJ + idx) + (wp >= 0 ? C.wallterm[wp][J] : 0.0f) : pop_ld(fI + idx + op);
    } else {
        h[I] = pop_ld(fJ + idx + om);
        h[J] = pop_ld(fI + idx + op);
    }
}

template <int I, bool CHECK>
FG_HD void odd_store_pair(const float (&h)[Q], const Lattice &L, const Collision &C, const Nbr &nb, long long idx) {
    using D = Dir<I>;
    constexpr int J = D::opp;
    pop_t *fI = L.f + I * L.slot, *fJ = L.f + J * L.slot;
    const int om = nb.off<-D::cx, -D::cy, -D::cz>();
    const int op = nb.off<D::cx, D::cy, D::cz>();
    if (CHECK) {
        const int wm = nb.wall<-D::cx, -D::cy, -D::cz>();
        const int wp = nb.wall<D::cx, D::cy, D::cz>();
        const bool bm = wm >= 0 || (L.solid && L.solid[idx + om]);
        const bool bp = wp >= 0 || (L.solid && L.solid[idx + op]);
        if (bp) pop_st(fJ + idx, h[I] + (wp >= 0 ? C.wallterm[wp][J] : 0.0f)); else pop_st(fI + idx + op, h[I]);
        if (bm) pop_st(fI + idx, h[J] + (wm >= 0 ? C.wallterm[wm][I] : 0.0f)); else pop_st(fJ + idx + om, h[J]);
    } else {
        pop_st(fI + idx + op, h[I]);
        pop_st(fJ + idx + om, h[J]);
    }
}

#define FG_FOR_PAIRS(X) X(1) X(3) X(5) X(7) X(8) X(11) X(12) X(15) X(16)

template <int PARITY, bool CHECK>
FG_HD void load_arriving(float (&h)[Q], const Lattice &L, const Collision &C, const Nbr &nb, long long idx) {
    if (PARITY == 0) {
        FG_UNROLL
        for (int i = 0; i < Q; ++i) h[i] = pop_ld(L.f + i * L.slot + idx);
    } else {
        h[0] = pop_ld(L.f + idx);
#define FG_X(I) odd_load_pair<I, CHECK>(h, L, C, nb, idx);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
    }
}

template <int PARITY, bool CHECK>
FG_HD void store_departing(const float (&h)[Q], const Lattice &L, const Collision &C, const Nbr &nb, long long idx) {
    if (PARITY == 0) {
        pop_st(L.f + idx, h[0]);
#define FG_X(I) pop_st(L.f + Dir<I>::opp * L.slot + idx, h[I]); pop_st(L.f + I * L.slot + idx, h[Dir<I>::opp]);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
    } else {
        pop_st(L.f + idx, h[0]);
#define FG_X(I) odd_store_pair<I, CHECK>(h, L, C, nb, idx);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
    }
}

// ---------------------------------------------------------------- kernel bodies
// grid: (ceil(nx/threads), rows, planes); one thread per cell, x fastest => coalesced 128 B per warp per slot.
// MODE is decided by the HOST per launch (sim.hpp launch_collide partitions the slab), so the bulk kernel carries no
// boundary code and fits 64 registers (8 CTAs of 128 threads per SM):
//   CHECK_NONE  no link can be blocked in the rows / planes of this launch
//   CHECK_XEDGE only the two cells at the ends of each row can (x walls): predicated pointer selects, no divergence
//   CHECK_XWARP the same, but only the two warps at the ends of a row run the predicated code (warp-uniform branch)
//   CHECK_ALL   wall rows (y walls), wall planes (z walls), or obstacles anywhere
enum : int { CHECK_NONE = 0, CHECK_ALL = 1, CHECK_XEDGE = 2, CHECK_XWARP = 3 };

// ---- bulk path addressing.  ncu showed the first version of the odd kernel issue-limited by 64-bit address
// arithmetic (4-5 integer instructions per access, more than the MRT itself).  Here a thread builds the 9 pointers to
// its (x+a, y+b) neighbours in the own plane once (periodic wrap folded in), and every access adds a launch-constant
// 64-bit byte offset kz[slot][dz] that the host precomputed and that sits in the constant bank: 2 integer
// instructions per access.  X walls (CHECK_XEDGE) are handled by predicated pointer selects on the ten links with
// c_x != 0 instead of a divergent checked path.
struct NbrPtrs {
    char *p[3][3];      // [a+1][b+1] -> slot 0 of cell (x+a, y+b, z)
    bool bxm, bxp;      // x wall on the low / high side of this cell (CHECK_XEDGE only)
};

template <bool XEDGE>
FG_HD NbrPtrs make_ptrs(const Lattice &L, int x, int y, long long idx) {
    NbrPtrs n;
    char *pc = reinterpret_cast<char *>(L.f + idx);
    const int dxm = (x == 0 ? L.nx - 1 : -1) * kPopBytes, dxp = (x == L.nx - 1 ? -(L.nx - 1) : 1) * kPopBytes;
    const int dym = (y == 0 ? (L.ny - 1) * L.nx : -L.nx) * kPopBytes, dyp = (y == L.ny - 1 ? -(L.ny - 1) * L.nx : L.nx) * kPopBytes;
    n.p[1][1] = pc;
    n.p[0][1] = pc + dxm; n.p[2][1] = pc + dxp;
    n.p[1][0] = pc + dym; n.p[1][2] = pc + dyp;
    n.p[0][0] = pc + (dxm + dym); n.p[2][0] = pc + (dxp + dym);
    n.p[0][2] = pc + (dxm + dyp); n.p[2][2] = pc + (dxp + dyp);
    n.bxm = XEDGE && x == 0;
    n.bxp = XEDGE && x == L.nx - 1;
    return n;
}

template <int I, bool XEDGE>
FG_HD void fast_odd_load_pair(float (&h)[Q], const StepParams &p, const NbrPtrs &n) {
    using D = Dir<I>;
    constexpr int J = D::opp;
    // f_I arrives from x - c_I (stored in slot J there); f_J arrives from x + c_I (stored in slot I there)
    const char *am = n.p[1 - D::cx][1 - D::cy] + p.kz[J][1 - D::cz];
    const char *ap = n.p[1 + D::cx][1 + D::cy] + p.kz[I][1 + D::cz];
    if (XEDGE && D::cx != 0) {
        const bool bm = D::cx > 0 ? n.bxm : n.bxp;   // the cell at x - c_I is behind the wall
        const bool bp = D::cx > 0 ? n.bxp : n.bxm;   // the cell at x + c_I is behind the wall
        const char *own_i = n.p[1][1] + p.kz[I][1], *own_j = n.p[1][1] + p.kz[J][1];
        const float wi = p.C.wallterm[D::cx > 0 ? F_XLO : F_XHI][I], wj = p.C.wallterm[D::cx > 0 ? F_XHI : F_XLO][J];
        h[I] = pop_ld(reinterpret_cast<const pop_t *>(bm ? own_i : am)) + (bm ? wi : 0.0f);
        h[J] = pop_ld(reinterpret_cast<const pop_t *>(bp ? own_j : ap)) + (bp ? wj : 0.0f);
    } else {
        h[I] = pop_ld(reinterpret_cast<const pop_t *>(am));
        h[J] = pop_ld(reinterpret_cast<const pop_t *>(ap));
    }
}

template <int I, bool XEDGE>
FG_HD void fast_odd_store_pair(const float (&h)[Q], const StepParams &p, const NbrPtrs &n) {
    using D = Dir<I>;
    constexpr int J = D::opp;
    char *am = n.p[1 - D::cx][1 - D::cy] + p.kz[J][1 - D::cz];
    char *ap = n.p[1 + D::cx][1 + D::cy] + p.kz[I][1 + D::cz];
    if (XEDGE && D::cx != 0) {
        const bool bm = D::cx > 0 ? n.bxm : n.bxp;
        const bool bp = D::cx > 0 ? n.bxp : n.bxm;
        char *own_i = n.p[1][1] + p.kz[I][1], *own_j = n.p[1][1] + p.kz[J][1];
        const float wi = p.C.wallterm[D::cx > 0 ? F_XLO : F_XHI][I], wj = p.C.wallterm[D::cx > 0 ? F_XHI : F_XLO][J];
        pop_st(reinterpret_cast<pop_t *>(bp ? own_j : ap), h[I] + (bp ? wj : 0.0f));   // f*_I bounces into slot J of this cell
        pop_st(reinterpret_cast<pop_t *>(bm ? own_i : am), h[J] + (bm ? wi : 0.0f));
    } else {
        pop_st(reinterpret_cast<pop_t *>(ap), h[I]);
        pop_st(reinterpret_cast<pop_t *>(am), h[J]);
    }
}

#if !defined(FG_CTA)
#define FG_CTA 128          // threads per stream-collide CTA (experiments: 64 with 18 CTAs/SM, 256 with 4)
#endif
constexpr int kCollideThreads = FG_CTA;
template <int PARITY, bool MRT, int MODE, int OCC = 9>
struct StreamCollide {
    static constexpr int kThreads = kCollideThreads;
    // 9 CTAs of 128 threads per SM: ptxas then fits every bulk variant into 56 registers without a spill (9 x 128 x 56 =
    // 64 512 of the 65 536 registers).  Measured against 8 (64 registers) and 10 (48 registers, 32-80 B of spills) in
    // gpu pass 21: 9 is 1.7-3.3 % faster than 8 on every workload, 10 is 3-4 % slower.
    // On z-slabs (boundary planes, halo push and IB chain at high priority beside the interior) the SAME comparison comes
    // out the other way — 0.1174 ms/step with 8 against 0.1214 with 9 on two GPUs, although the kernel alone is faster
    // with 9 (multi-GPU pass 6) — so peered handles launch the OCC = 8 instantiation.
#if defined(FG_OCC)
    static constexpr int kMinBlocks = FG_OCC;
#else
    static constexpr int kMinBlocks = OCC * 128 / kCollideThreads;
#endif

    FG_HD static void force_at(const StepParams &p, int y, int zz, long long idx, float &Fx, float &Fy, float &Fz) {
        Fx = p.C.g[0]; Fy = p.C.g[1]; Fz = p.C.g[2];
        if (p.F.cellslot && p.F.rowflag[zz * p.L.ny + y]) {
            const int s = p.F.cellslot[idx];
            if (s > 0) {
                Fx += p.F.bandF[s - 1];
                Fy += p.F.bandF[p.F.band_cap + s - 1];
                Fz += p.F.bandF[2 * p.F.band_cap + s - 1];
            }
        }
    }

    // general path: every link checked against walls and obstacles
    FG_HD static void checked_cell(const StepParams &p, int x, int y, int zz) {
        const Lattice &L = p.L;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
        if (L.solid && L.solid[idx]) return;
        const Nbr nb = make_nbr(L, x, y, zz);
        float h[Q];
        load_arriving<PARITY, true>(h, L, p.C, nb, idx);
        float Fx, Fy, Fz;
        force_at(p, y, zz, idx, Fx, Fy, Fz);
        if (MRT) collide_mrt(h, Fx, Fy, Fz, p.C); else collide_bgk(h, Fx, Fy, Fz, p.C);
        store_departing<PARITY, true>(h, L, p.C, nb, idx);
    }

    // bulk path: no blocked link except (XEDGE) across the x walls
    template <bool XEDGE>
    FG_HD static void bulk_cell(const StepParams &p, int x, int y, int zz) {
        const Lattice &L = p.L;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
        float h[Q];
        if (PARITY == 0) {
            // even step: 19 aligned, purely local accesses; plain indexing compiles to the leanest code here
            const Nbr nb{};
            load_arriving<0, false>(h, L, p.C, nb, idx);
            float Fx, Fy, Fz;
            force_at(p, y, zz, idx, Fx, Fy, Fz);
            if (MRT) collide_mrt(h, Fx, Fy, Fz, p.C); else collide_bgk(h, Fx, Fy, Fz, p.C);
            store_departing<0, false>(h, L, p.C, nb, idx);
        } else {
            const NbrPtrs n = make_ptrs<XEDGE>(L, x, y, idx);
            h[0] = pop_ld(reinterpret_cast<const pop_t *>(n.p[1][1] + p.kz[0][1]));
#define FG_X(I) fast_odd_load_pair<I, XEDGE>(h, p, n);
            FG_FOR_PAIRS(FG_X)
#undef FG_X
            float Fx, Fy, Fz;
            force_at(p, y, zz, idx, Fx, Fy, Fz);
            if (MRT) collide_mrt(h, Fx, Fy, Fz, p.C); else collide_bgk(h, Fx, Fy, Fz, p.C);
            pop_st(reinterpret_cast<pop_t *>(n.p[1][1] + p.kz[0][1]), h[0]);
#define FG_X(I) fast_odd_store_pair<I, XEDGE>(h, p, n);
            FG_FOR_PAIRS(FG_X)
#undef FG_X
        }
    }

    FG_HD static void run(const StepParams &p, int bx, int by, int bz, int tx) {
        const int x = bx * kThreads + tx, y = p.y0 + by * p.ystride;
        // CTAs are dispatched in blockIdx order: odd steps sweep the planes downwards, so a step starts on the planes
        // the previous one touched last and finds them in the 126 MB L2 (a third of a 256x128x128 lattice)
        int zz = p.zz_begin + (p.zz_flip >= 0 ? p.zz_flip - bz : bz) * p.zz_stride;
        if (zz >= p.zz_skip_begin) zz += p.zz_skip_len;
        if (x >= p.L.nx) return;
        if (MODE == CHECK_ALL) checked_cell(p, x, y, zz);
        else if (MODE == CHECK_XEDGE) bulk_cell<true>(p, x, y, zz);
        else if (MODE == CHECK_XWARP) {
            // only the warps that hold the first / last cell of a row pay for the wall selects (warp-uniform branch)
            const int w0 = x & ~31;
            if (w0 == 0 || w0 + 32 >= p.L.nx) bulk_cell<true>(p, x, y, zz);
            else bulk_cell<false>(p, x, y, zz);
        }
        else bulk_cell<false>(p, x, y, zz);
    }
};

// ---------------------------------------------------------------- even step, V cells per thread, 16-byte accesses
// BASELINE.json:5 (a) names "128-bit vectorised coalesced loads".  In the EVEN step of the AA pattern all 19 accesses of a
// cell are local and aligned, so a thread can take V = 4 (or 2) consecutive cells of a row with one LDG.128 / STG.128
// (LDG.64) per slot: 19 + 19 memory instructions per V cells instead of 19 V + 19 V.  The V cells are collided one after
// the other in the same registers (19 V population registers + the collide's temporaries), so the register budget is
// what limits occupancy: ptxas numbers and the measured A/B against the scalar kernel are in profiles/r2_summary.md.
// Same arithmetic per cell as StreamCollide<0>: bit-identical results (tests).  Opt-in: FG_FLAG_EVEN_VEC4 / _VEC2.
template <int V> struct alignas(sizeof(pop_t) * V) VecF { pop_t a[V]; };      // V consecutive cells of one slot, in the storage type

template <bool MRT, int V, bool NARROW = false>
struct StreamCollideEvenVec {
    static constexpr int kThreads = kCollideThreads;
#if defined(FG_VEC_OCC)
    static constexpr int kMinBlocks = FG_VEC_OCC;
#else
    static constexpr int kMinBlocks = V == 4 ? 4 : 6;
#endif
    using Scalar = StreamCollide<0, MRT, CHECK_NONE>;

    template <int C>
    FG_HD static void cell(VecF<V> (&v)[Q], const StepParams &p, int y, int zz, long long idx) {
        float h[Q];
        FG_UNROLL
        for (int i = 0; i < Q; ++i) h[i] = pop_dec(v[i].a[C]);
        float Fx, Fy, Fz;
        Scalar::force_at(p, y, zz, idx + C, Fx, Fy, Fz);
        if (MRT) collide_mrt(h, Fx, Fy, Fz, p.C); else collide_bgk(h, Fx, Fy, Fz, p.C);
        v[0].a[C] = pop_enc(h[0]);
#define FG_X(I) v[Dir<I>::opp].a[C] = pop_enc(h[I]); v[I].a[C] = pop_enc(h[Dir<I>::opp]);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
    }

    // grid: (ceil(nx / V / threads), rows, planes); requires nx % V == 0
    FG_HD static void run(const StepParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        int x0 = (bx * kThreads + tx) * V, row = by;
        if (NARROW) {                           // narrow rows: this CTA holds several of them
            row = (by << p.rows_per_cta_log2) + (x0 >> p.nx_log2);
            x0 &= L.nx - 1;
            if (row >= p.rows) return;
        }
        const int y = p.y0 + row * p.ystride;
        int zz = p.zz_begin + (p.zz_flip >= 0 ? p.zz_flip - bz : bz) * p.zz_stride;
        if (zz >= p.zz_skip_begin) zz += p.zz_skip_len;
        if (x0 >= L.nx) return;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x0;
        VecF<V> v[Q];
        FG_UNROLL
        for (int i = 0; i < Q; ++i) v[i] = *reinterpret_cast<const VecF<V> *>(L.f + i * L.slot + idx);
        cell<0>(v, p, y, zz, idx);
        cell<1>(v, p, y, zz, idx);
        if (V == 4) {
            cell<V == 4 ? 2 : 0>(v, p, y, zz, idx);
            cell<V == 4 ? 3 : 0>(v, p, y, zz, idx);
        }
        FG_UNROLL
        for (int i = 0; i < Q; ++i) *reinterpret_cast<VecF<V> *>(L.f + i * L.slot + idx) = v[i];
    }
};

// ---------------------------------------------------------------- odd step, 2 cells per thread (bulk rows)
// The odd step reads f_i from slot opp(i) of the neighbour at x - c_i and writes f*_i to slot i of the neighbour at x + c_i.
// For the 9 slots with c_x = 0 the two cells (x0, x0 + 1) of a thread are an aligned pair in the neighbour row: one 64-bit
// access.  The 10 slots with c_x != 0 are one float off the pair boundary and stay 32-bit accesses.  Same arithmetic per
// cell and the same 19 + 19 locations per cell as StreamCollide<1>: bit-identical results.  A/B: profiles/r2_summary.md.
// XWALL (rows between x walls: the tank, the school): the first cell of a row has no neighbour at x - 1 and the last none at
// x + 1; the links across those faces bounce back into the cell's own opposite slot (+ the moving-wall term), on the read
// and on the write side, exactly as CHECK_XEDGE does for one cell per thread.  Only the two warps at the ends of a row run
// that predicated code (warp-uniform branch, as CHECK_XWARP).
template <bool MRT, bool XWALL = false, bool NARROW = false>
struct StreamCollideOddVec2 {
    static constexpr int kThreads = kCollideThreads;
#if defined(FG_ODDVEC_OCC)
    static constexpr int kMinBlocks = FG_ODDVEC_OCC;
#else
    static constexpr int kMinBlocks = 6;      // 80 registers, no spills (ptxas); at 5 (<= 102 registers) the BGK variant spills
#endif
    using Scalar = StreamCollide<1, MRT, CHECK_NONE>;

    template <int I>
    FG_HD static void addr(const StepParams &p, char *pc, int dym, int dyp, char *&am, char *&ap) {
        using D = Dir<I>;
        constexpr int J = D::opp;
        am = pc + (D::cy > 0 ? dym : (D::cy < 0 ? dyp : 0)) + p.kz[J][1 - D::cz];      // slot J, row y - c_y, plane z - c_z, at x0
        ap = pc + (D::cy > 0 ? dyp : (D::cy < 0 ? dym : 0)) + p.kz[I][1 + D::cz];      // slot I, row y + c_y, plane z + c_z, at x0
    }
    FG_HD static float lds(const char *q) { return pop_ld(reinterpret_cast<const pop_t *>(q)); }
    FG_HD static void sts(char *q, float v) { pop_st(reinterpret_cast<pop_t *>(q), v); }
    // b0: cell x0 is the first of its row and x - 1 is behind a wall; b1: cell x0 + 1 is the last and x + 1 is behind a wall
    template <int I, bool XW>
    FG_HD static void load_pair(float (&h0)[Q], float (&h1)[Q], const StepParams &p, char *pc, int dym, int dyp, int oxm, int oxp, bool b0, bool b1) {
        using D = Dir<I>;
        constexpr int J = D::opp;
        constexpr int B = kPopBytes;
        char *am, *ap;
        addr<I>(p, pc, dym, dyp, am, ap);
        if (D::cx == 0) {
            const VecF<2> vm = *reinterpret_cast<const VecF<2> *>(am), vp = *reinterpret_cast<const VecF<2> *>(ap);
            h0[I] = pop_dec(vm.a[0]); h1[I] = pop_dec(vm.a[1]); h0[J] = pop_dec(vp.a[0]); h1[J] = pop_dec(vp.a[1]);
        } else if (D::cx > 0) {
            if (XW) {
                h0[I] = lds(b0 ? pc + p.kz[I][1] : am + oxm) + (b0 ? p.C.wallterm[F_XLO][I] : 0.0f);
                h1[J] = lds(b1 ? pc + B + p.kz[J][1] : ap + oxp) + (b1 ? p.C.wallterm[F_XHI][J] : 0.0f);
            } else {
                h0[I] = lds(am + oxm); h1[J] = lds(ap + oxp);
            }
            h1[I] = lds(am);
            h0[J] = lds(ap + B);
        } else {
            if (XW) {
                h1[I] = lds(b1 ? pc + B + p.kz[I][1] : am + oxp) + (b1 ? p.C.wallterm[F_XHI][I] : 0.0f);
                h0[J] = lds(b0 ? pc + p.kz[J][1] : ap + oxm) + (b0 ? p.C.wallterm[F_XLO][J] : 0.0f);
            } else {
                h1[I] = lds(am + oxp); h0[J] = lds(ap + oxm);
            }
            h0[I] = lds(am + B);
            h1[J] = lds(ap);
        }
    }
    // f*_I goes where f_J came from and f*_J where f_I came from
    template <int I, bool XW>
    FG_HD static void store_pair(const float (&h0)[Q], const float (&h1)[Q], const StepParams &p, char *pc, int dym, int dyp, int oxm, int oxp, bool b0, bool b1) {
        using D = Dir<I>;
        constexpr int J = D::opp;
        constexpr int B = kPopBytes;
        char *am, *ap;
        addr<I>(p, pc, dym, dyp, am, ap);
        if (D::cx == 0) {
            VecF<2> vp, vm;
            vp.a[0] = pop_enc(h0[I]); vp.a[1] = pop_enc(h1[I]); vm.a[0] = pop_enc(h0[J]); vm.a[1] = pop_enc(h1[J]);
            *reinterpret_cast<VecF<2> *>(ap) = vp; *reinterpret_cast<VecF<2> *>(am) = vm;
        } else if (D::cx > 0) {
            sts(ap + B, h0[I]);
            sts(am, h1[J]);
            if (XW) {
                sts(b1 ? pc + B + p.kz[J][1] : ap + oxp, h1[I] + (b1 ? p.C.wallterm[F_XHI][J] : 0.0f));
                sts(b0 ? pc + p.kz[I][1] : am + oxm, h0[J] + (b0 ? p.C.wallterm[F_XLO][I] : 0.0f));
            } else {
                sts(ap + oxp, h1[I]); sts(am + oxm, h0[J]);
            }
        } else {
            sts(ap, h1[I]);
            sts(am + B, h0[J]);
            if (XW) {
                sts(b0 ? pc + p.kz[J][1] : ap + oxm, h0[I] + (b0 ? p.C.wallterm[F_XLO][J] : 0.0f));
                sts(b1 ? pc + B + p.kz[I][1] : am + oxp, h1[J] + (b1 ? p.C.wallterm[F_XHI][I] : 0.0f));
            } else {
                sts(ap + oxm, h0[I]); sts(am + oxp, h1[J]);
            }
        }
    }

    template <bool XW>
    FG_HD static void body(const StepParams &p, int x0, int y, int zz) {
        const Lattice &L = p.L;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x0;
        char *pc = reinterpret_cast<char *>(L.f + idx);
        const int dym = (y == 0 ? (L.ny - 1) * L.nx : -L.nx) * kPopBytes, dyp = (y == L.ny - 1 ? -(L.ny - 1) * L.nx : L.nx) * kPopBytes;
        const int oxm = (x0 == 0 ? L.nx - 1 : -1) * kPopBytes, oxp = (x0 + 2 == L.nx ? -(L.nx - 2) : 2) * kPopBytes;   // cells x0 - 1 and x0 + 2 (periodic)
        const bool b0 = XW && x0 == 0, b1 = XW && x0 + 2 == L.nx;
        float h0[Q], h1[Q];
        {
            const VecF<2> v = *reinterpret_cast<const VecF<2> *>(pc + p.kz[0][1]);
            h0[0] = pop_dec(v.a[0]); h1[0] = pop_dec(v.a[1]);
        }
#define FG_X(I) load_pair<I, XW>(h0, h1, p, pc, dym, dyp, oxm, oxp, b0, b1);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
        float Fx, Fy, Fz;
        Scalar::force_at(p, y, zz, idx, Fx, Fy, Fz);
        if (MRT) collide_mrt(h0, Fx, Fy, Fz, p.C); else collide_bgk(h0, Fx, Fy, Fz, p.C);
        Scalar::force_at(p, y, zz, idx + 1, Fx, Fy, Fz);
        if (MRT) collide_mrt(h1, Fx, Fy, Fz, p.C); else collide_bgk(h1, Fx, Fy, Fz, p.C);
        {
            VecF<2> v;
            v.a[0] = pop_enc(h0[0]); v.a[1] = pop_enc(h1[0]);
            *reinterpret_cast<VecF<2> *>(pc + p.kz[0][1]) = v;
        }
#define FG_X(I) store_pair<I, XW>(h0, h1, p, pc, dym, dyp, oxm, oxp, b0, b1);
        FG_FOR_PAIRS(FG_X)
#undef FG_X
    }

    // grid: (nx / 2 / threads, rows, planes); requires nx % 2 == 0, x periodic or (XWALL) between walls, no other blocked
    // link in the rows of the launch
    FG_HD static void run(const StepParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        int x0 = (bx * kThreads + tx) * 2, row = by;
        if (NARROW) {                           // narrow rows: this CTA holds several of them
            row = (by << p.rows_per_cta_log2) + (x0 >> p.nx_log2);
            x0 &= L.nx - 1;
            if (row >= p.rows) return;
        }
        const int y = p.y0 + row * p.ystride;
        int zz = p.zz_begin + (p.zz_flip >= 0 ? p.zz_flip - bz : bz) * p.zz_stride;
        if (zz >= p.zz_skip_begin) zz += p.zz_skip_len;
        if (x0 >= L.nx) return;
        if (XWALL) {
            const int w0 = x0 & ~63;                                   // first cell of this warp (32 lanes x 2 cells)
            if (w0 == 0 || w0 + 64 >= L.nx) body<true>(p, x0, y, zz);
            else body<false>(p, x0, y, zz);
        } else {
            body<false>(p, x0, y, zz);
        }
    }
};

// ---------------------------------------------------------------- two steps in one launch (L2-resident wavefront)
// StreamCollidePair runs the EVEN step s and the ODD step s+1 of a range of planes in ONE launch: the odd step follows
// the even step `lag` planes behind, so it finds the populations the even step has just written in the 126 MB L2 and
// DRAM sees one read and one write per cell for the two updates instead of two of each.
//   * CTAs take tickets (atomic counter) in launch order and map them to (phase, plane, row, x-block) along the
//     schedule E0 .. E(lag-1), E(lag), O(0), E(lag+1), O(1), ...; a lower ticket has always started, so waiting on
//     lower tickets cannot deadlock.
//   * an odd-phase CTA of plane v waits until every even-phase CTA of planes v-1, v, v+1 has finished (per-plane
//     completion counters, release/acquire at gpu scope); with lag >= 3 the wait is almost never taken.
//   * the odd phase leaves out planes whose neighbours are not part of the launch (range ends, hole edges): the host
//     steps those in the following odd substep, after the z-face operations of the even step.
// Per cell the arithmetic is exactly that of StreamCollide<0> followed by StreamCollide<1>: results are bit-identical.
// MEASURED (r1 passes 11-13, 256^3 and 512^3 periodic boxes): DRAM traffic per pair drops by 43 % (L2 hit rate 64 %),
// but the pair takes 973 us against 411 + 419 us for the two plain launches: a CTA now has to wait for its stores to be
// acknowledged before it can count itself done (fence), and for the ticket before its first load, which costs more
// occupancy than the saved DRAM time returns (stall reasons: barrier 3.9, membar 0.6 warps per issue).  A variant with a
// static blockIdx schedule and per-warp counters was 2.5x slower still.  Hence OPT-IN (FG_FLAG_FUSED_PAIRS), off by default.
struct PairParams {
    StepParams s;           // planes: zz = s.zz_begin + v (+ hole); rows: y = s.y0 + r * s.ystride (EVEN phase: every row)
    int planes, rows, xblocks;
    int lag;
    int odd_lo, odd_hi;     // virtual planes [odd_lo, odd_hi) take the odd step here, except ...
    int odd_skip_v;         // ... the two planes next to the hole: v == odd_skip_v - 1 and v == odd_skip_v (< 0: no hole)
    int odd_y_lo, odd_y_hi; // rows [odd_y_lo, odd_y_hi) take the BULK odd step, the others (y-wall rows) the checked one
    int *ticket;            // [1], zero at launch
    int *done;              // [planes] even-phase CTAs finished per plane, zero at launch
};

template <bool MRT, int MODE>
struct StreamCollidePair {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 8;
    static constexpr int kGridPhases = 2;
    using Even = StreamCollide<0, MRT, CHECK_NONE>;
    using Odd = StreamCollide<1, MRT, MODE>;
    using OddChecked = StreamCollide<1, MRT, CHECK_ALL>;

    // schedule position q (in units of whole planes) -> phase (0 even, 1 odd) and virtual plane
    FG_HD static void decode(const PairParams &p, int q, int &phase, int &v) {
        const int k = p.lag, pairs = p.planes - k;
        if (q < k) { phase = 0; v = q; return; }
        const int q2 = q - k;
        if (q2 < 2 * pairs) { phase = q2 & 1; v = (q2 >> 1) + (phase ? 0 : k); return; }
        phase = 1; v = pairs + (q2 - 2 * pairs);
    }
    FG_HD static bool odd_valid(const PairParams &p, int v) {
        return v >= p.odd_lo && v < p.odd_hi && !(p.odd_skip_v >= 0 && (v == p.odd_skip_v - 1 || v == p.odd_skip_v));
    }
    FG_HD static int plane_of(const PairParams &p, int v) {
        int zz = p.s.zz_begin + v;
        if (zz >= p.s.zz_skip_begin) zz += p.s.zz_skip_len;
        return zz;
    }
    FG_HD static void cell(const PairParams &p, int phase, int v, int r, int bx, int tx) {
        const int x = bx * kThreads + tx, y = p.s.y0 + r * p.s.ystride, zz = plane_of(p, v);
        if (x >= p.s.L.nx) return;
        if (phase == 0) Even::template bulk_cell<false>(p.s, x, y, zz);
        else if (y >= p.odd_y_lo && y < p.odd_y_hi) Odd::template bulk_cell<MODE == CHECK_XEDGE>(p.s, x, y, zz);
        else checked_row(p.s, x, y, zz);
    }
#if defined(__CUDACC__)
    __device__ __noinline__
#endif
    static void checked_row(const StepParams &s, int x, int y, int zz) { OddChecked::checked_cell(s, x, y, zz); }

    // host emulation: grid (xblocks, rows, planes), all of phase 0 before phase 1
    FG_HD static void run(const PairParams &p, int bx, int by, int bz, int tx, int phase) {
        if (phase == 1 && !odd_valid(p, bz)) return;
        cell(p, phase, bz, by, bx, tx);
    }
#if defined(__CUDACC__)
    __device__ __forceinline__ static void cta(const PairParams &p, int tx) {
        __shared__ int s_ticket;
        if (tx == 0) s_ticket = atomicAdd(p.ticket, 1);
        __syncthreads();
        const int per_plane = p.rows * p.xblocks;
        const int q = s_ticket / per_plane, w = s_ticket - q * per_plane;
        const int r = w / p.xblocks, bx = w - r * p.xblocks;
        int phase, v;
        decode(p, q, phase, v);
        if (phase == 1) {
            if (!odd_valid(p, v)) return;
            if (tx == 0) {
                for (int d = -1; d <= 1; ++d) {
                    const int *c = p.done + v + d;
                    int n;
                    for (;;) {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(n) : "l"(c) : "memory");
                        if (n >= per_plane) break;
                        __nanosleep(100);
                    }
                }
                __threadfence();
            }
            __syncthreads();
        }
        cell(p, phase, v, r, bx, tx);
        if (phase == 0) {
            __syncthreads();
            if (tx == 0) {
                __threadfence();
                atomicAdd(p.done + v, 1);
            }
        }
    }
#endif
};

// f <- shifted equilibrium of (rho, u) given per cell, or of a constant state: natural layout, all planes
struct InitParams {
    Lattice L;
    const float *rho, *u;   // device [nz][ny][nx], [3][nz][ny][nx] (local, no ghosts) or nullptr => (1, 0)
};
FG_HD void shifted_equilibrium(float dr, float ux, float uy, float uz, float (&h)[Q]) {
    const float rho = 1.0f + dr, uu15 = 1.5f * (ux * ux + uy * uy + uz * uz);
    FG_UNROLL
    for (int i = 0; i < Q; ++i) {
        const float cu = cxr(i) * ux + cyr(i) * uy + czr(i) * uz;
        const float w = i == 0 ? W0 : (i < 7 ? W1 : W2);
        h[i] = w * (dr + rho * (3.0f * cu + 4.5f * cu * cu - uu15));
    }
}
struct InitEquilibrium {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const InitParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int x = bx * kThreads + tx, y = by, zz = bz;   // all nz+2 planes
        if (x >= L.nx) return;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
        float dr = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
        if (p.rho && zz >= 1 && zz <= L.nz) {
            const long long l = ((long long)(zz - 1) * L.ny + y) * L.nx + x, n = (long long)L.nz * L.plane;
            dr = p.rho[l] - 1.0f; ux = p.u[l]; uy = p.u[n + l]; uz = p.u[2 * n + l];
        }
        float h[Q];
        shifted_equilibrium(dr, ux, uy, uz, h);
        FG_UNROLL
        for (int i = 0; i < Q; ++i) pop_st(L.f + i * L.slot + idx, h[i]);
    }
};

// arriving populations / moments for read-out (parity aware): out19 natural layout [19][nz][ny][nx] (shifted),
// or mom [4][nz][ny][nx] = (rho-1, u)
struct GatherParams {
    Lattice L;
    Collision C;
    float *out19;
    float *mom;
};
template <int PARITY>
struct GatherArriving {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const GatherParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int x = bx * kThreads + tx, y = by, zz = bz + 1;
        if (x >= L.nx) return;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
        const long long l = ((long long)(zz - 1) * L.ny + y) * L.nx + x, n = (long long)L.nz * L.plane;
        const Nbr nb = make_nbr(L, x, y, zz);
        float h[Q];
        if (L.solid && L.solid[idx]) {
            FG_UNROLL
            for (int i = 0; i < Q; ++i) h[i] = pop_ld(L.f + i * L.slot + idx);   // solid cells keep whatever they hold
        } else {
            load_arriving<PARITY, true>(h, L, p.C, nb, idx);
        }
        if (p.out19) {
            FG_UNROLL
            for (int i = 0; i < Q; ++i) p.out19[i * n + l] = h[i];
        }
        if (p.mom) {
            float dr, jx, jy, jz;
            moments(h, dr, jx, jy, jz);
            const float inv = 1.0f / (1.0f + dr);
            p.mom[l] = dr; p.mom[n + l] = jx * inv; p.mom[2 * n + l] = jy * inv; p.mom[3 * n + l] = jz * inv;
        }
    }
};

// natural layout <- gathered populations (the inverse read-out): with GatherArriving<1> before it this completes the
// streaming an even step has left pending, WITHOUT a collision — what fg_set_solid needs before the obstacle mask may change
struct ScatterNatural {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const GatherParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int x = bx * kThreads + tx, y = by, zz = bz + 1;
        if (x >= L.nx) return;
        const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
        const long long l = ((long long)(zz - 1) * L.ny + y) * L.nx + x, n = (long long)L.nz * L.plane;
        FG_UNROLL
        for (int i = 0; i < Q; ++i) pop_st(L.f + i * L.slot + idx, p.out19[i * n + l]);
    }
};

// fg_get_solid_force (momentum exchange): one thread per cell gathers the populations arriving at its cell with the
// parity-aware loader — a population that was reflected off an obstacle arrives from the cell's own opposite slot — and every
// such link hands the obstacle -2 c_i f_i.  "Off an obstacle" is the loader's own test (no wall face crossed, obstacle flag of the
// source cell set; the flags of the ghost planes encode periodic wrap / outlet clamping, sim.hpp set_solid).  Warp-reduced in
// fp64, six atomics per warp that saw a link.  A read-out, not part of a step.
struct SolidForceParams {
    Lattice L;
    Collision C;
    double *out;            // [6], zeroed before the launch
    double ox, oy, oz;      // torque reference point (global z)
    int torque;
};
FG_HD double warp_sum_d(double v, bool &leader, int tx) {
#if defined(__CUDA_ARCH__)
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
    leader = (tx & 31) == 0;
#else
    (void)tx;
    leader = true;      // host emulation: every "thread" adds its own part
#endif
    return v;
}
FG_HD void atomic_add_f64(double *p, double v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
template <int PARITY>
struct SolidForce {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    template <int I>
    FG_HD static void link(const float (&h)[Q], const SolidForceParams &p, const Nbr &nb, long long idx, double rx, double ry, double rz, double (&a)[6]) {
        using D = Dir<I>;
        constexpr int J = D::opp;
        const Lattice &L = p.L;
        // f_I arrives from x - c_I, f_J from x + c_I (odd_load_pair)
        const bool bm = nb.wall<-D::cx, -D::cy, -D::cz>() < 0 && L.solid[idx + nb.off<-D::cx, -D::cy, -D::cz>()];
        const bool bp = nb.wall<D::cx, D::cy, D::cz>() < 0 && L.solid[idx + nb.off<D::cx, D::cy, D::cz>()];
        constexpr double w = I < 7 ? 1.0 / 18 : 1.0 / 36;      // lattice weight of both directions of the pair (populations are stored shifted by it)
        if (bm) add(2.0 * (double(h[I]) + w), -D::cx, -D::cy, -D::cz, p, rx, ry, rz, a);
        if (bp) add(2.0 * (double(h[J]) + w), D::cx, D::cy, D::cz, p, rx, ry, rz, a);
    }
    // a population of mass m/2 reflected at the link midpoint x + e/2 (e points from the cell to the obstacle): the obstacle takes m e
    FG_HD static void add(double m, int ex, int ey, int ez, const SolidForceParams &p, double rx, double ry, double rz, double (&a)[6]) {
        const double fx = ex * m, fy = ey * m, fz = ez * m;
        a[0] += fx; a[1] += fy; a[2] += fz;
        if (p.torque) {
            const double x = rx + 0.5 * ex, y = ry + 0.5 * ey, z = rz + 0.5 * ez;
            a[3] += y * fz - z * fy; a[4] += z * fx - x * fz; a[5] += x * fy - y * fx;
        }
    }
    FG_HD static void run(const SolidForceParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int x = bx * kThreads + tx, y = by, zz = bz + 1;
        double a[6] = {0, 0, 0, 0, 0, 0};
        bool any = false;
        if (x < L.nx) {
            const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
            if (!L.solid[idx]) {
                const Nbr nb = make_nbr(L, x, y, zz);
                float h[Q];
                load_arriving<PARITY, true>(h, L, p.C, nb, idx);
                const double rx = x - p.ox, ry = y - p.oy, rz = (L.z0 + zz - 1) - p.oz;
#define FG_X(I) link<I>(h, p, nb, idx, rx, ry, rz, a);
                FG_FOR_PAIRS(FG_X)
#undef FG_X
                any = a[0] != 0 || a[1] != 0 || a[2] != 0 || a[3] != 0 || a[4] != 0 || a[5] != 0;
            }
        }
#if defined(__CUDA_ARCH__)
        if (!__any_sync(0xffffffffu, any)) return;      // most warps see no obstacle
#else
        if (!any) return;
#endif
        FG_UNROLL
        for (int c = 0; c < 6; ++c) {
            bool leader;
            const double v = warp_sum_d(a[c], leader, tx);
            if (leader && v != 0) atomic_add_f64(p.out + c, v);
        }
    }
};

// divergence guard of the fluid (fg_check_finite): one thread per cell reads the REST population only (4 B per cell) and
// counts cells where it is not finite or has left |h_0| <= 4 (the density would be off by more than 12).  A NaN in any
// population of a cell reaches h_0 with the cell's next collision, so a blow-up is seen one step later at most.
struct FiniteParams {
    Lattice L;
    unsigned long long *bad;    // [1], zeroed before the launch
};
struct FiniteCheck {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 8;
    FG_HD static void run(const FiniteParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int x = bx * kThreads + tx, y = by, zz = bz + 1;
        bool bad = false;
        if (x < L.nx) {
            const long long idx = ((long long)zz * L.ny + y) * L.nx + x;
            if (!(L.solid && L.solid[idx])) bad = !(fabsf(pop_ld(L.f + idx)) <= 4.0f);
        }
#if defined(__CUDA_ARCH__)
        const unsigned m = __ballot_sync(0xffffffffu, bad);
        if (m && (tx & 31) == 0) atomicAdd(p.bad, (unsigned long long)__popc(m));
#else
        if (bad) ++*p.bad;
#endif
    }
};

// ---------------------------------------------------------------- z-face plane operations (a10)
// After every step the 5 populations crossing each z face are moved between boundary / ghost planes
// (SURVEY.md A8 "slot subtlety for halos").  `dst` may be this rank's own lattice (periodic wrap on one GPU),
// or a z-neighbour's lattice mapped over NVLink (peer stores), so this kernel IS the halo exchange.
//   after EVEN step (parity_done = 0):  boundary plane, slots opp(movers leaving)  ->  neighbour ghost plane
//   after ODD  step (parity_done = 1):  own ghost plane, slots of movers that left  ->  neighbour boundary plane
// Inlet / outlet are the same operation with a constant or the adjacent plane as the source.
struct FaceOp {
    int mode;                  // BC_WALL: nothing; BC_PEER: copy src -> dst; BC_INLET; BC_OUTLET
    int hi;                    // which face of the SENDER (copy) / of this lattice (inlet, outlet) the op is for
    const pop_t *src;          // copy source: element (k, c) = src[k*src_slot + src_off + c]
    long long src_slot, src_off;
    int src_by_index;          // 1: k = 0..4 (a packed message); 0: k = lattice slot number
    pop_t *dst;                // copy destination lattice base (own, or a z-neighbour's over NVLink); slot stride L.slot
    long long dst_off;
    const uint8_t *sender_solid;   // solid flags of the sender's boundary plane (plane-sized) or nullptr
    const uint8_t *receiver_solid; // solid flags of the receiver's boundary plane (plane-sized) or nullptr
};
struct HaloParams {
    Lattice L;
    Collision C;
    FaceOp op[2];
    int parity_done;
};

// does the link that brings population i into cell (x,y) cross an x/y wall?  (then bounce-back owns that slot)
FG_HD bool link_from_wall(const Lattice &L, int i, int x, int y) {
    const int sx = x - cxr(i), sy = y - cyr(i);
    return (L.wall_x && (sx < 0 || sx >= L.nx)) || (L.wall_y && (sy < 0 || sy >= L.ny));
}

struct ZFaceOp {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    // grid: (ceil(plane/threads), 5 slots, 2 ops)
    FG_HD static void run(const HaloParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const int c = bx * kThreads + tx;
        if (c >= L.plane) return;
        const FaceOp &op = p.op[bz];
        if (op.mode == BC_WALL) return;
        const bool hi = op.hi != 0;
        const int y = c / L.nx, x = c - y * L.nx;
        if (op.mode == BC_PEER) {
            // even: movers leaving through the sender's face sit in slot opp(i) of its boundary plane (hi: slots ZM)
            // odd : pushes that crossed the face landed in the sender's ghost plane, natural slot i (hi: slots ZP)
            const int slot = p.parity_done == 0 ? (hi ? zmr(by) : zpr(by)) : (hi ? zpr(by) : zmr(by));
            if (p.parity_done == 1) {
                if (link_from_wall(L, slot, x, y)) return;   // nothing was pushed here; the receiver's bounce-back owns the slot
                if (op.sender_solid) {
                    int sx = x - cxr(slot), sy = y - cyr(slot);   // the would-be sender; solid cells do not push
                    sx = sx < 0 ? sx + L.nx : (sx >= L.nx ? sx - L.nx : sx);
                    sy = sy < 0 ? sy + L.ny : (sy >= L.ny ? sy - L.ny : sy);
                    if (op.sender_solid[sy * L.nx + sx]) return;
                }
            }
            // An obstacle cell of the receiving boundary plane keeps what it holds (nothing reads it but the immersed-
            // boundary moments, which see the initial state there, like the oracle).
            if (p.parity_done == 1 && op.receiver_solid && op.receiver_solid[c]) return;
            const int k = op.src_by_index ? by : slot;
            op.dst[slot * L.slot + op.dst_off + c] = op.src[k * op.src_slot + op.src_off + c];
            return;
        }
        // inlet / outlet: populations ENTERING through this face, direction i (lo: ZP, hi: ZM)
        const int i = hi ? zmr(by) : zpr(by);
        if (p.parity_done == 0) {
            // the next (odd) step pulls them from ghost slot opp(i)
            const int gs = oppr(i);
            pop_t *ghost = L.f + gs * L.slot + (long long)(hi ? L.nz + 1 : 0) * L.plane + c;
            if (op.mode == BC_INLET) pop_st(ghost, p.C.heq_in[i]);
            else *ghost = L.f[gs * L.slot + (long long)(hi ? L.nz : 1) * L.plane + c];   // outlet: copy of the last plane
        } else {
            // the next (even) step reads them from natural slot i of the boundary plane
            if (link_from_wall(L, i, x, y)) return;
            if (L.solid && L.solid[(long long)(hi ? L.nz : 1) * L.plane + c]) return;   // obstacle cells keep what they hold
            if (op.mode == BC_OUTLET && L.solid) {
                // the clamped pull source (x-cx, y-cy) of the same plane is an obstacle: bounce-back owns the slot
                int sx = x - cxr(i), sy = y - cyr(i);
                sx = sx < 0 ? sx + L.nx : (sx >= L.nx ? sx - L.nx : sx);
                sy = sy < 0 ? sy + L.ny : (sy >= L.ny ? sy - L.ny : sy);
                if (L.solid[((long long)(hi ? L.nz : 1) * L.ny + sy) * L.nx + sx]) return;
                // The plane inside is an obstacle at (x,y): the source cell's push towards it was bounced into the
                // source's own slot opp(i) and nothing arrived below, so the post-collision value is taken from there.
                if (L.solid[(long long)(hi ? L.nz - 1 : 2) * L.plane + c]) {
                    L.f[i * L.slot + (long long)(hi ? L.nz : 1) * L.plane + c] =
                        L.f[oppr(i) * L.slot + ((long long)(hi ? L.nz : 1) * L.ny + sy) * L.nx + sx];
                    return;
                }
            }
            pop_t *cell = L.f + i * L.slot + (long long)(hi ? L.nz : 1) * L.plane + c;
            if (op.mode == BC_INLET) pop_st(cell, p.C.heq_in[i]);
            else *cell = L.f[i * L.slot + (long long)(hi ? L.nz - 1 : 2) * L.plane + c];   // outlet: what the plane inside received
        }
    }
};

// Four cells per thread for the plain plane copies of ZFaceOp (periodic wrap on one GPU, halo push to a z-neighbour): without x / y
// walls and obstacles every cell of the plane is copied unconditionally, so a thread moves 16 bytes (8 in the 16-bit build) per
// slot — at 512^2 the one-cell form took 17 us per step for 21 MB on one GPU.  The host launches it only when both face operations
// are copies (or absent) and plane % 4 == 0; everything else (inlet / outlet, walls, obstacles, packed messages) stays with ZFaceOp.
struct ZFaceCopy4 {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    // grid: (ceil(plane / 4 / threads), 5 slots, 2 ops)
    FG_HD static void run(const HaloParams &p, int bx, int by, int bz, int tx) {
        const Lattice &L = p.L;
        const long long c = 4ll * (bx * kThreads + tx);
        if (c >= L.plane) return;
        const FaceOp &op = p.op[bz];
        if (op.mode != BC_PEER) return;
        const bool hi = op.hi != 0;
        const int slot = p.parity_done == 0 ? (hi ? zmr(by) : zpr(by)) : (hi ? zpr(by) : zmr(by));
        *reinterpret_cast<VecF<4> *>(op.dst + slot * L.slot + op.dst_off + c) =
            *reinterpret_cast<const VecF<4> *>(op.src + slot * op.src_slot + op.src_off + c);
    }
};

}  // namespace fg
