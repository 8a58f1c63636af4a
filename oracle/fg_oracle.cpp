// fg_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A small, readable fp64 OpenMP restatement of the coupled fluid step: D3Q19 BGK / MRT
// lattice-Boltzmann (two-lattice, textbook collide-then-pull-stream), Guo forcing, half-way
// bounce-back, equilibrium inlet / zero-gradient outlet on the z faces, Peskin 4-point
// immersed-boundary direct forcing and per-link wrench reductions.  It exports the same C ABI as
// the CUDA library (include/fishgym.h) so the same Python and the same tests drive either.
//
// What it restates: nothing in /root/reference can be followed — the checkout is README.md:1-15
// only (README.md:2-3 names the coupled agent-fluid path; no source, tests or golden vectors).
// The algorithm is the one BASELINE.json:5 prescribes, written out in SURVEY.md Appendix A1-A9.
// PARITY IS UNPINNED BY THE REFERENCE; it is pinned by the analytic results in tests/
// (Taylor-Green decay, Poiseuille parabola, sphere drag) and the algebraic identities of A1-A6.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product path never does.
//
// Deliberately different in construction from the CUDA path: unshifted populations, explicit
// 19x19 moment matrix (m_eq = M f_eq and M*source evaluated numerically, no closed forms),
// separate collide and stream passes, two lattices, pull streaming.
#include "../include/fishgym.h"
#include "oracle_body.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

constexpr int Q = 19;
// d'Humieres ordering
constexpr int CX[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
constexpr int CY[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
constexpr int CZ[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
constexpr int OPP[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
constexpr double W[Q] = {1.0 / 3, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18,
                         1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36,
                         1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
// the 5 populations crossing a z face (A1): +z movers and -z movers
constexpr int ZP[5] = {5, 11, 12, 15, 16};
constexpr int ZM[5] = {6, 13, 14, 17, 18};

thread_local std::string g_create_error;

struct Moments {
    double M[Q][Q];
    double Minv[Q][Q];
    Moments() {
        // rows as polynomials of c
        for (int i = 0; i < Q; ++i) {
            const double x = CX[i], y = CY[i], z = CZ[i];
            const double c2 = x * x + y * y + z * z;
            M[0][i] = 1;
            M[1][i] = 19 * c2 - 30;
            M[2][i] = (21 * c2 * c2 - 53 * c2 + 24) / 2;
            M[3][i] = x;
            M[4][i] = (5 * c2 - 9) * x;
            M[5][i] = y;
            M[6][i] = (5 * c2 - 9) * y;
            M[7][i] = z;
            M[8][i] = (5 * c2 - 9) * z;
            M[9][i] = 3 * x * x - c2;
            M[10][i] = (3 * c2 - 5) * (3 * x * x - c2);
            M[11][i] = y * y - z * z;
            M[12][i] = (3 * c2 - 5) * (y * y - z * z);
            M[13][i] = x * y;
            M[14][i] = y * z;
            M[15][i] = x * z;
            M[16][i] = (y * y - z * z) * x;
            M[17][i] = (z * z - x * x) * y;
            M[18][i] = (x * x - y * y) * z;
        }
        for (int k = 0; k < Q; ++k) {
            double n2 = 0;
            for (int i = 0; i < Q; ++i) n2 += M[k][i] * M[k][i];
            for (int i = 0; i < Q; ++i) Minv[i][k] = M[k][i] / n2;   // rows orthogonal => M^-1 = M^T diag(1/|row|^2)
        }
    }
};
const Moments &moments() {
    static const Moments m;
    return m;
}

inline void equilibrium(double rho, double ux, double uy, double uz, double *feq) {
    // SURVEY.md A2
    const double uu = ux * ux + uy * uy + uz * uz;
    for (int i = 0; i < Q; ++i) {
        const double cu = CX[i] * ux + CY[i] * uy + CZ[i] * uz;
        feq[i] = W[i] * rho * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * uu);
    }
}

// Peskin 4-point kernel
inline double peskin4(double r) {
    r = std::fabs(r);
    if (r < 1.0) return (3.0 - 2.0 * r + std::sqrt(1.0 + 4.0 * r - 4.0 * r * r)) / 8.0;
    if (r < 2.0) return (5.0 - 2.0 * r - std::sqrt(-7.0 + 12.0 * r - 4.0 * r * r)) / 8.0;
    return 0.0;
}

}  // namespace

struct FgSim {
    FgConfig cfg{};
    int nx = 0, ny = 0, nz = 0;   // local slab
    int nzg = 0, z0 = 0;          // global height, first owned global plane
    size_t plane = 0, ncell = 0;  // ncell includes 2 ghost planes
    std::vector<double> f;        // [Q][nz+2][ny][nx] populations ARRIVING at time t (ghost planes unused)
    std::vector<double> fs;       // post-collision f*; ghost planes receive neighbours' f* (halo)
    std::vector<uint8_t> solid;   // [nz+2][ny][nx] incl. ghost planes
    bool has_solid = false;
    double rates[Q]{};
    double AT[Q][Q]{}, BT[Q][Q]{};   // transposed relaxation matrices A = M^-1 S M, B = M^-1 (I - S/2) M
    double omega = 1.25;
    double feq_in[Q]{};
    // immersed boundary
    int n_markers = 0, n_links = 0, n_origins = 0;
    std::vector<float> mX, mU, mdV;
    std::vector<int32_t> mlink;
    std::vector<int32_t> mbase, mowner;
    std::vector<double> mF, mUstar;        // [n][3]
    std::vector<double> link_origin;       // [links][3]
    std::vector<double> wrench;            // [links][6]
    std::vector<double> Fx, Fy, Fz;        // Eulerian IB force (dense, incl. ghosts)
    bool force_dirty = false;
    int band_cells = 0;
    // bodies
    std::vector<obody::Fish> fish;
    std::vector<float> action;
    // halo state (multi-rank host-staged protocol)
    bool stream_pending = false;
    int faces_received = 0;
    // stats
    int64_t steps = 0;
    double last_ms = 0, last_mlups = 0;
    std::string err;

    size_t idx(int x, int y, int zl) const { return (size_t(zl + 1) * ny + y) * nx + x; }  // zl in [-1, nz]
    double *F(std::vector<double> &a, int i) { return a.data() + size_t(i) * ncell; }
    const double *F(const std::vector<double> &a, int i) const { return a.data() + size_t(i) * ncell; }
    bool internal_face(int face) const {
        if (cfg.n_ranks <= 1) return false;
        if (face == FG_ZLO) return cfg.rank > 0 || cfg.bc[FG_ZLO] == FG_BC_PERIODIC;
        return cfg.rank < cfg.n_ranks - 1 || cfg.bc[FG_ZHI] == FG_BC_PERIODIC;
    }
    int n_internal_faces() const { return int(internal_face(FG_ZLO)) + int(internal_face(FG_ZHI)); }
};

namespace {

int fail(FgSim *s, int code, const std::string &msg) {
    if (s) s->err = msg; else g_create_error = msg;
    return code;
}

void set_equilibrium_everywhere(FgSim *s, double rho, double ux, double uy, double uz) {
    double feq[Q];
    equilibrium(rho, ux, uy, uz, feq);
    for (int i = 0; i < Q; ++i) std::fill(s->F(s->f, i), s->F(s->f, i) + s->ncell, feq[i]);
}

// -------- collide: f -> fs on owned cells --------
// m* = m - S (m - m_eq) + (I - S/2) M Phi  and  f* = M^-1 m*  collapse into two constant matrices,
//   f* = f - A (f - f_eq) + B Phi,   A = M^-1 S M,   B = M^-1 (I - S/2) M     (BGK: A = omega I, B = (1 - omega/2) I)
// stored transposed so that the products run in vectorisable axpy form (no reassociation needed).
void build_relaxation(FgSim *s) {
    const Moments &mm = moments();
    for (int k = 0; k < Q; ++k)
        for (int i = 0; i < Q; ++i) {
            double a = 0, b = 0;
            for (int r = 0; r < Q; ++r) {
                a += mm.Minv[i][r] * s->rates[r] * mm.M[r][k];
                b += mm.Minv[i][r] * (1.0 - 0.5 * s->rates[r]) * mm.M[r][k];
            }
            s->AT[k][i] = a;   // AT[k][i] = A[i][k]
            s->BT[k][i] = b;
        }
}

void collide(FgSim *s) {
    const bool mrt = s->cfg.collision == FG_MRT;
    const double gx = s->cfg.body_force[0], gy = s->cfg.body_force[1], gz = s->cfg.body_force[2];
    const bool ibf = s->force_dirty;
    const bool any_force = ibf || gx != 0 || gy != 0 || gz != 0;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < s->nz; ++z)
        for (int y = 0; y < s->ny; ++y)
            for (int x = 0; x < s->nx; ++x) {
                const size_t c = s->idx(x, y, z);
                double fi[Q], feq[Q], phi[Q], d[Q], out[Q];
                for (int i = 0; i < Q; ++i) fi[i] = s->F(s->f, i)[c];
                if (s->has_solid && s->solid[c]) {
                    for (int i = 0; i < Q; ++i) s->F(s->fs, i)[c] = fi[i];
                    continue;
                }
                double Fxc = gx, Fyc = gy, Fzc = gz;
                if (ibf) { Fxc += s->Fx[c]; Fyc += s->Fy[c]; Fzc += s->Fz[c]; }
                double rho = 0, jx = 0, jy = 0, jz = 0;
                for (int i = 0; i < Q; ++i) {
                    rho += fi[i];
                    jx += CX[i] * fi[i];
                    jy += CY[i] * fi[i];
                    jz += CZ[i] * fi[i];
                }
                const double ux = (jx + 0.5 * Fxc) / rho, uy = (jy + 0.5 * Fyc) / rho, uz = (jz + 0.5 * Fzc) / rho;
                equilibrium(rho, ux, uy, uz, feq);
                for (int i = 0; i < Q; ++i) {
                    const double cu = CX[i] * ux + CY[i] * uy + CZ[i] * uz;
                    phi[i] = W[i] * ((3.0 * (CX[i] - ux) + 9.0 * cu * CX[i]) * Fxc +
                                     (3.0 * (CY[i] - uy) + 9.0 * cu * CY[i]) * Fyc +
                                     (3.0 * (CZ[i] - uz) + 9.0 * cu * CZ[i]) * Fzc);
                    d[i] = fi[i] - feq[i];
                }
                if (!mrt) {
                    const double om = s->omega;
                    for (int i = 0; i < Q; ++i) out[i] = fi[i] - om * d[i] + (1.0 - 0.5 * om) * phi[i];
                } else {
                    for (int i = 0; i < Q; ++i) out[i] = fi[i];
                    for (int k = 0; k < Q; ++k) {
                        const double dk = d[k];
                        for (int i = 0; i < Q; ++i) out[i] -= s->AT[k][i] * dk;
                    }
                    if (any_force)
                        for (int k = 0; k < Q; ++k) {
                            const double pk = phi[k];
                            for (int i = 0; i < Q; ++i) out[i] += s->BT[k][i] * pk;
                        }
                }
                for (int i = 0; i < Q; ++i) s->F(s->fs, i)[c] = out[i];
            }
}

// -------- the pull rule: where does f_i(x, t+1) come from? --------
// returns the arriving population i at owned cell (x,y,z) given post-collision field `fs`
inline double pull(const FgSim *s, const std::vector<double> &fs, int i, int x, int y, int z) {
    const FgConfig &c = s->cfg;
    int sx = x - CX[i], sy = y - CY[i], sz = z - CZ[i];
    int wall_face = -1;
    if (sx < 0) { if (c.bc[FG_XLO] == FG_BC_PERIODIC) sx = s->nx - 1; else wall_face = FG_XLO; }
    else if (sx >= s->nx) { if (c.bc[FG_XHI] == FG_BC_PERIODIC) sx = 0; else wall_face = FG_XHI; }
    if (sy < 0) { if (c.bc[FG_YLO] == FG_BC_PERIODIC) sy = s->ny - 1; else if (wall_face < 0) wall_face = FG_YLO; }
    else if (sy >= s->ny) { if (c.bc[FG_YHI] == FG_BC_PERIODIC) sy = 0; else if (wall_face < 0) wall_face = FG_YHI; }
    bool inlet = false;
    const int zg = s->z0 + sz;
    if (zg < 0 || zg >= s->nzg) {
        const int face = zg < 0 ? FG_ZLO : FG_ZHI;
        switch (c.bc[face]) {
            case FG_BC_PERIODIC:
                if (c.n_ranks <= 1) sz = zg < 0 ? s->nz - 1 : 0;   // else: ghost plane filled by the ring neighbour
                break;
            case FG_BC_WALL: if (wall_face < 0) wall_face = face; break;
            case FG_BC_INLET: inlet = true; break;
            case FG_BC_OUTLET: sz = z; break;
        }
    }
    const size_t here = s->idx(x, y, z);
    if (wall_face >= 0) {
        const double *uw = c.wall_u[wall_face];
        return s->F(fs, OPP[i])[here] + 6.0 * W[i] * (CX[i] * uw[0] + CY[i] * uw[1] + CZ[i] * uw[2]);
    }
    if (inlet) return s->feq_in[i];
    const size_t src = s->idx(sx, sy, sz);
    if (s->has_solid && s->solid[src]) return s->F(fs, OPP[i])[here];
    return s->F(fs, i)[src];
}

void stream(FgSim *s) {
#pragma omp parallel for schedule(static)
    for (int z = 0; z < s->nz; ++z)
        for (int y = 0; y < s->ny; ++y)
            for (int x = 0; x < s->nx; ++x) {
                const size_t c = s->idx(x, y, z);
                if (s->has_solid && s->solid[c]) continue;
                const int zg = s->z0 + z;
                if (x > 0 && x < s->nx - 1 && y > 0 && y < s->ny - 1 && zg > 0 && zg < s->nzg - 1) {
                    // interior cell: no face is crossed, the pull rule reduces to f_i(x) = f*_i(x - c_i), or to the cell's own
                    // opposite population where x - c_i is an obstacle (what pull() returns for such a cell)
                    if (!s->has_solid) {
                        for (int i = 0; i < Q; ++i)
                            s->F(s->f, i)[c] = s->F(s->fs, i)[c - (ptrdiff_t(CZ[i]) * s->ny + CY[i]) * s->nx - CX[i]];
                    } else {
                        for (int i = 0; i < Q; ++i) {
                            const size_t src = c - (ptrdiff_t(CZ[i]) * s->ny + CY[i]) * s->nx - CX[i];
                            s->F(s->f, i)[c] = s->solid[src] ? s->F(s->fs, OPP[i])[c] : s->F(s->fs, i)[src];
                        }
                    }
                    continue;
                }
                for (int i = 0; i < Q; ++i) s->F(s->f, i)[c] = pull(s, s->fs, i, x, y, z);
            }
    s->stream_pending = false;
    s->faces_received = 0;
}

int finish_pending(FgSim *s) {
    if (!s->stream_pending) return FG_OK;
    if (s->faces_received < s->n_internal_faces())
        return fail(s, FG_ESTATE, "halo exchange incomplete: unpack every internal face after fg_step");
    stream(s);
    return FG_OK;
}

// -------- immersed boundary --------
inline int wrap_or_skip(int v, int n, bool periodic) {
    if (v >= 0 && v < n) return v;
    if (!periodic) return -1;
    v %= n;
    return v < 0 ? v + n : v;
}

// the stencil of marker k: fn(cell, w) for every node that lies on this rank's lattice, w = product of the three 1-D weights
template <class Fn>
void stencil_nodes(const FgSim *s, int k, Fn fn) {
    const FgConfig &c = s->cfg;
    const bool px = c.bc[FG_XLO] == FG_BC_PERIODIC, py = c.bc[FG_YLO] == FG_BC_PERIODIC, pz = c.bc[FG_ZLO] == FG_BC_PERIODIC;
    double wx[4], wy[4], wz[4];
    const double X = s->mX[3 * k], Y = s->mX[3 * k + 1], Z = s->mX[3 * k + 2];
    const int i0 = s->mbase[3 * k], j0 = s->mbase[3 * k + 1], k0 = s->mbase[3 * k + 2];
    for (int a = 0; a < 4; ++a) {
        wx[a] = peskin4(X - (i0 + a));
        wy[a] = peskin4(Y - (j0 + a));
        wz[a] = peskin4(Z - (k0 + a));
    }
    for (int cz = 0; cz < 4; ++cz) {
        const int zg = wrap_or_skip(k0 + cz, s->nzg, pz);
        if (zg < 0) continue;
        const int zl = zg - s->z0;
        if (zl < 0 || zl >= s->nz) continue;
        for (int cy = 0; cy < 4; ++cy) {
            const int yy = wrap_or_skip(j0 + cy, s->ny, py);
            if (yy < 0) continue;
            for (int cx = 0; cx < 4; ++cx) {
                const int xx = wrap_or_skip(i0 + cx, s->nx, px);
                if (xx < 0) continue;
                fn(s->idx(xx, yy, zl), wx[cx] * wy[cy] * wz[cz]);
            }
        }
    }
}

void ib_forces(FgSim *s) {
    const int n = s->n_markers;
    const FgConfig &c = s->cfg;
    const bool px = c.bc[FG_XLO] == FG_BC_PERIODIC, py = c.bc[FG_YLO] == FG_BC_PERIODIC,
               pz = c.bc[FG_ZLO] == FG_BC_PERIODIC;
    std::fill(s->Fx.begin(), s->Fx.end(), 0.0);
    std::fill(s->Fy.begin(), s->Fy.end(), 0.0);
    std::fill(s->Fz.begin(), s->Fz.end(), 0.0);
    std::vector<uint8_t> touched(s->ncell, 0);
    // (a3) index map on the fp32 coordinates, integer arithmetic only afterwards
    for (int k = 0; k < n; ++k) {
        for (int d = 0; d < 3; ++d) s->mbase[3 * k + d] = int32_t(std::floor(s->mX[3 * k + d])) - 1;
        int kc = wrap_or_skip(int32_t(std::floor(s->mX[3 * k + 2])), s->nzg, pz);
        if (kc < 0) kc = std::min(std::max(int32_t(std::floor(s->mX[3 * k + 2])), 0), s->nzg - 1);
        s->mowner[k] = kc / s->nz;
    }
    // (a4..a6) unforced velocity at stencil nodes, interpolate, direct forcing
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        double wx[4], wy[4], wz[4];
        const double X = s->mX[3 * k], Y = s->mX[3 * k + 1], Z = s->mX[3 * k + 2];
        const int i0 = s->mbase[3 * k], j0 = s->mbase[3 * k + 1], k0 = s->mbase[3 * k + 2];
        for (int a = 0; a < 4; ++a) {
            wx[a] = peskin4(X - (i0 + a));
            wy[a] = peskin4(Y - (j0 + a));
            wz[a] = peskin4(Z - (k0 + a));
        }
        double us[3] = {0, 0, 0};
        for (int cz = 0; cz < 4; ++cz) {
            const int zg = wrap_or_skip(k0 + cz, s->nzg, pz);
            if (zg < 0) continue;
            const int zl = zg - s->z0;
            if (zl < 0 || zl >= s->nz) continue;
            for (int cy = 0; cy < 4; ++cy) {
                const int yy = wrap_or_skip(j0 + cy, s->ny, py);
                if (yy < 0) continue;
                for (int cx = 0; cx < 4; ++cx) {
                    const int xx = wrap_or_skip(i0 + cx, s->nx, px);
                    if (xx < 0) continue;
                    const size_t cell = s->idx(xx, yy, zl);
                    touched[cell] = 1;
                    double rho = 0, jx = 0, jy = 0, jz = 0;
                    for (int i = 0; i < Q; ++i) {
                        const double v = s->F(s->f, i)[cell];
                        rho += v; jx += CX[i] * v; jy += CY[i] * v; jz += CZ[i] * v;
                    }
                    const double w = wx[cx] * wy[cy] * wz[cz];
                    us[0] += w * jx / rho; us[1] += w * jy / rho; us[2] += w * jz / rho;
                }
            }
        }
        for (int d = 0; d < 3; ++d) {
            s->mUstar[3 * k + d] = us[d];
            s->mF[3 * k + d] = 2.0 * (double(s->mU[3 * k + d]) - us[d]);   // rho0 = 1
        }
    }
    // (a7) spread — serial over markers so the sum order is fixed
    for (int k = 0; k < n; ++k) {
        double wx[4], wy[4], wz[4];
        const double X = s->mX[3 * k], Y = s->mX[3 * k + 1], Z = s->mX[3 * k + 2];
        const int i0 = s->mbase[3 * k], j0 = s->mbase[3 * k + 1], k0 = s->mbase[3 * k + 2];
        for (int a = 0; a < 4; ++a) {
            wx[a] = peskin4(X - (i0 + a));
            wy[a] = peskin4(Y - (j0 + a));
            wz[a] = peskin4(Z - (k0 + a));
        }
        const double dV = s->mdV[k];
        for (int cz = 0; cz < 4; ++cz) {
            const int zg = wrap_or_skip(k0 + cz, s->nzg, pz);
            if (zg < 0) continue;
            const int zl = zg - s->z0;
            if (zl < 0 || zl >= s->nz) continue;
            for (int cy = 0; cy < 4; ++cy) {
                const int yy = wrap_or_skip(j0 + cy, s->ny, py);
                if (yy < 0) continue;
                for (int cx = 0; cx < 4; ++cx) {
                    const int xx = wrap_or_skip(i0 + cx, s->nx, px);
                    if (xx < 0) continue;
                    const size_t cell = s->idx(xx, yy, zl);
                    const double w = wx[cx] * wy[cy] * wz[cz] * dV;
                    s->Fx[cell] += w * s->mF[3 * k];
                    s->Fy[cell] += w * s->mF[3 * k + 1];
                    s->Fz[cell] += w * s->mF[3 * k + 2];
                }
            }
        }
    }
    // (a6, optional) multi-direct forcing, SURVEY.md A7 (4) "n_iter > 1": with the Guo half-force the fluid velocity the
    // collide works with is u = u* + F/(2 rho0), so after one pass the markers see U*_k + E_k/2, E_k = sum_x F(x) delta_h(x - X_k),
    // not U_d,k.  Every further pass adds the Jacobi correction dF_k = 2 rho0 (U_d,k - U*_k - E_k/2) to ALL markers from the
    // same force field (gather first, then spread), so the no-slip residual contracts pass by pass.
    for (int it = 1; it < std::max(1, c.ib_iterations); ++it) {
        std::vector<double> dF(3 * size_t(n), 0.0);
#pragma omp parallel for schedule(static)
        for (int k = 0; k < n; ++k) {
            double e[3] = {0, 0, 0};
            stencil_nodes(s, k, [&](size_t cell, double w) {
                e[0] += w * s->Fx[cell]; e[1] += w * s->Fy[cell]; e[2] += w * s->Fz[cell];
            });
            for (int d = 0; d < 3; ++d) dF[3 * k + d] = 2.0 * (double(s->mU[3 * k + d]) - s->mUstar[3 * k + d]) - e[d];
        }
        for (int k = 0; k < n; ++k) {          // serial: fixed sum order
            const double dV = s->mdV[k];
            stencil_nodes(s, k, [&](size_t cell, double w) {
                s->Fx[cell] += w * dV * dF[3 * k]; s->Fy[cell] += w * dV * dF[3 * k + 1]; s->Fz[cell] += w * dV * dF[3 * k + 2];
            });
            for (int d = 0; d < 3; ++d) s->mF[3 * k + d] += dF[3 * k + d];
        }
    }
    // (a8) per-link wrench ON the body = minus what the body exerts on the fluid
    std::fill(s->wrench.begin(), s->wrench.end(), 0.0);
    for (int k = 0; k < n; ++k) {
        const int l = s->mlink[k];
        if (l < 0 || l >= s->n_links) continue;
        const double dV = s->mdV[k];
        const double fx = -s->mF[3 * k] * dV, fy = -s->mF[3 * k + 1] * dV, fz = -s->mF[3 * k + 2] * dV;
        const double rx = double(s->mX[3 * k]) - s->link_origin[3 * l], ry = double(s->mX[3 * k + 1]) - s->link_origin[3 * l + 1],
                     rz = double(s->mX[3 * k + 2]) - s->link_origin[3 * l + 2];
        double *w = &s->wrench[6 * l];
        w[0] += fx; w[1] += fy; w[2] += fz;
        w[3] += ry * fz - rz * fy;
        w[4] += rz * fx - rx * fz;
        w[5] += rx * fy - ry * fx;
    }
    int bc = 0;
    for (size_t i = 0; i < s->ncell; ++i) bc += touched[i];
    s->band_cells = bc;
    s->force_dirty = true;
}

void ensure_ib_storage(FgSim *s) {
    if (s->Fx.size() != s->ncell) {
        s->Fx.assign(s->ncell, 0.0);
        s->Fy.assign(s->ncell, 0.0);
        s->Fz.assign(s->ncell, 0.0);
    }
}

void bodies_to_markers(FgSim *s) {
    // host body state -> markers rounded to fp32 once
    int n = 0, nl = 0;
    for (auto &fh : s->fish) { n += fh.n_markers(); nl += fh.n_links(); }
    s->n_markers = n; s->n_links = nl;
    s->mX.resize(3 * size_t(n)); s->mU.resize(3 * size_t(n)); s->mdV.resize(n); s->mlink.resize(n);
    s->mbase.assign(3 * size_t(n), 0); s->mowner.assign(n, 0);
    s->mF.assign(3 * size_t(n), 0.0); s->mUstar.assign(3 * size_t(n), 0.0);
    s->link_origin.assign(3 * size_t(nl), 0.0);
    s->wrench.resize(6 * size_t(nl), 0.0);
    int mo = 0, lo = 0;
    for (auto &fh : s->fish) {
        fh.emit_markers(&s->mX[3 * size_t(mo)], &s->mU[3 * size_t(mo)], &s->mdV[mo], &s->mlink[mo], lo,
                        &s->link_origin[3 * size_t(lo)]);
        mo += fh.n_markers(); lo += fh.n_links();
    }
}

}  // namespace

extern "C" {

int fg_abi_version(void) { return FG_ABI_VERSION; }
const char *fg_backend_name(void) { return "oracle-fp64"; }
const char *fg_last_error(const FgSim *sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int fg_config_default(FgConfig *cfg) {
    if (!cfg) return FG_EINVAL;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = int32_t(sizeof(FgConfig));
    cfg->nx = cfg->ny = cfg->nz = 32;
    cfg->collision = FG_BGK;
    cfg->n_ranks = 1;
    cfg->tau = 0.8;
    cfg->inlet_rho = 1.0;
    return FG_OK;
}

int fg_create(const FgConfig *cfg, FgSim **out) {
    if (!cfg || !out) return fail(nullptr, FG_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->struct_size != int32_t(sizeof(FgConfig))) return fail(nullptr, FG_EINVAL, "FgConfig.struct_size mismatch (ABI)");
    if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 1) return fail(nullptr, FG_EINVAL, "lattice dimensions must be >= 1");
    if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks || cfg->nz % cfg->n_ranks)
        return fail(nullptr, FG_EINVAL, "bad slab decomposition: need 0 <= rank < n_ranks and nz % n_ranks == 0");
    if (!(cfg->tau > 0.5)) return fail(nullptr, FG_EINVAL, "tau must be > 0.5");
    if (cfg->collision != FG_BGK && cfg->collision != FG_MRT) return fail(nullptr, FG_EINVAL, "unknown collision model");
    if (cfg->ib_iterations < 0 || cfg->ib_iterations > 16) return fail(nullptr, FG_EINVAL, "ib_iterations must be in 0 .. 16");
    for (int f = 0; f < 6; ++f) {
        const int b = cfg->bc[f];
        if (b < FG_BC_PERIODIC || b > FG_BC_OUTLET) return fail(nullptr, FG_EINVAL, "unknown boundary condition");
        if (f < FG_ZLO && b > FG_BC_WALL) return fail(nullptr, FG_EINVAL, "inlet/outlet are supported on the z faces only");
    }
    for (int a = 0; a < 3; ++a)
        if ((cfg->bc[2 * a] == FG_BC_PERIODIC) != (cfg->bc[2 * a + 1] == FG_BC_PERIODIC))
            return fail(nullptr, FG_EINVAL, "periodic boundaries must be set on both faces of an axis");
    FgSim *s = new (std::nothrow) FgSim;
    if (!s) return fail(nullptr, FG_ENOMEM, "out of memory");
    s->cfg = *cfg;
    if (s->cfg.inlet_rho == 0) s->cfg.inlet_rho = 1.0;
    s->nx = cfg->nx; s->ny = cfg->ny; s->nzg = cfg->nz; s->nz = cfg->nz / cfg->n_ranks; s->z0 = cfg->rank * s->nz;
    s->plane = size_t(s->nx) * s->ny;
    s->ncell = s->plane * (s->nz + 2);
    try {
        s->f.assign(size_t(Q) * s->ncell, 0.0);
        s->fs.assign(size_t(Q) * s->ncell, 0.0);
        s->solid.assign(s->ncell, 0);
    } catch (const std::bad_alloc &) {
        delete s;
        return fail(nullptr, FG_ENOMEM, "out of memory allocating lattices");
    }
    s->omega = 1.0 / cfg->tau;
    bool all_zero = true;
    for (int k = 0; k < Q; ++k) all_zero = all_zero && cfg->mrt_rates[k] == 0.0;
    if (all_zero) {
        // SURVEY.md A3 defaults
        const double sn = s->omega;
        const double d[Q] = {0, 1.19, 1.4, 0, 1.2, 0, 1.2, 0, 1.2, sn, 1.4, sn, 1.4, sn, sn, sn, 1.98, 1.98, 1.98};
        std::copy(d, d + Q, s->rates);
    } else {
        std::copy(cfg->mrt_rates, cfg->mrt_rates + Q, s->rates);
        s->rates[0] = s->rates[3] = s->rates[5] = s->rates[7] = 0.0;   // conserved moments
    }
    build_relaxation(s);
    equilibrium(s->cfg.inlet_rho, cfg->inlet_u[0], cfg->inlet_u[1], cfg->inlet_u[2], s->feq_in);
    set_equilibrium_everywhere(s, 1.0, 0, 0, 0);
    if (cfg->max_markers > 0) ensure_ib_storage(s);
    *out = s;
    return FG_OK;
}

int fg_destroy(FgSim *sim) {
    delete sim;
    return FG_OK;
}

int fg_reset(FgSim *s, uint64_t seed) {
    if (!s) return FG_EINVAL;
    (void)seed;
    set_equilibrium_everywhere(s, 1.0, 0, 0, 0);
    s->stream_pending = false; s->faces_received = 0; s->steps = 0; s->force_dirty = false;
    for (auto &fh : s->fish) fh.reset();
    std::fill(s->action.begin(), s->action.end(), 0.f);
    if (!s->fish.empty()) { bodies_to_markers(s); std::fill(s->wrench.begin(), s->wrench.end(), 0.0); }
    return FG_OK;
}

int fg_set_fields(FgSim *s, const float *rho, const float *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    const size_t nloc = s->plane * s->nz;
    for (int z = 0; z < s->nz; ++z)
        for (int y = 0; y < s->ny; ++y)
            for (int x = 0; x < s->nx; ++x) {
                const size_t l = (size_t(z) * s->ny + y) * s->nx + x, c = s->idx(x, y, z);
                double feq[Q];
                equilibrium(rho[l], u[l], u[nloc + l], u[2 * nloc + l], feq);
                for (int i = 0; i < Q; ++i) s->F(s->f, i)[c] = feq[i];
            }
    s->stream_pending = false;
    return FG_OK;
}

static int get_fields_impl(FgSim *s, float *rf, float *uf, double *rd, double *ud) {
    if (int rc = finish_pending(s)) return rc;
    const size_t nloc = s->plane * s->nz;
    for (int z = 0; z < s->nz; ++z)
        for (int y = 0; y < s->ny; ++y)
            for (int x = 0; x < s->nx; ++x) {
                const size_t l = (size_t(z) * s->ny + y) * s->nx + x, c = s->idx(x, y, z);
                double rho = 0, jx = 0, jy = 0, jz = 0;
                for (int i = 0; i < Q; ++i) {
                    const double v = s->F(s->f, i)[c];
                    rho += v; jx += CX[i] * v; jy += CY[i] * v; jz += CZ[i] * v;
                }
                if (rd) { rd[l] = rho; ud[l] = jx / rho; ud[nloc + l] = jy / rho; ud[2 * nloc + l] = jz / rho; }
                else { rf[l] = float(rho); uf[l] = float(jx / rho); uf[nloc + l] = float(jy / rho); uf[2 * nloc + l] = float(jz / rho); }
            }
    return FG_OK;
}
int fg_get_fields(FgSim *s, float *rho, float *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    return get_fields_impl(s, rho, u, nullptr, nullptr);
}
int fg_get_fields_f64(FgSim *s, double *rho, double *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    return get_fields_impl(s, nullptr, nullptr, rho, u);
}

int fg_set_populations(FgSim *s, const float *f19) {
    if (!s || !f19) return FG_EINVAL;
    const size_t nloc = s->plane * s->nz;
    for (int i = 0; i < Q; ++i)
        for (int z = 0; z < s->nz; ++z)
            for (int y = 0; y < s->ny; ++y)
                for (int x = 0; x < s->nx; ++x)
                    s->F(s->f, i)[s->idx(x, y, z)] = f19[size_t(i) * nloc + (size_t(z) * s->ny + y) * s->nx + x];
    s->stream_pending = false;
    return FG_OK;
}

int fg_get_populations(FgSim *s, float *f19) {
    if (!s || !f19) return FG_EINVAL;
    if (int rc = finish_pending(s)) return rc;
    const size_t nloc = s->plane * s->nz;
    for (int i = 0; i < Q; ++i)
        for (int z = 0; z < s->nz; ++z)
            for (int y = 0; y < s->ny; ++y)
                for (int x = 0; x < s->nx; ++x)
                    f19[size_t(i) * nloc + (size_t(z) * s->ny + y) * s->nx + x] = float(s->F(s->f, i)[s->idx(x, y, z)]);
    return FG_OK;
}

int fg_set_solid(FgSim *s, const uint8_t *g) {
    if (!s) return FG_EINVAL;
    std::fill(s->solid.begin(), s->solid.end(), 0);
    s->has_solid = false;
    if (!g) return FG_OK;
    const bool pz = s->cfg.bc[FG_ZLO] == FG_BC_PERIODIC;
    for (int zl = -1; zl <= s->nz; ++zl) {
        int zg = s->z0 + zl;
        if (zg < 0 || zg >= s->nzg) { if (!pz) continue; zg = (zg + s->nzg) % s->nzg; }
        for (size_t p = 0; p < s->plane; ++p) {
            const uint8_t v = g[size_t(zg) * s->plane + p] ? 1 : 0;
            s->solid[size_t(zl + 1) * s->plane + p] = v;
            s->has_solid = s->has_solid || v;
        }
    }
    return FG_OK;
}

int fg_set_markers(FgSim *s, int32_t n, const float *X, const float *U, const float *dV, const int32_t *link) {
    if (!s || n < 0 || (n > 0 && (!X || !U || !dV))) return FG_EINVAL;
    if (n > s->cfg.max_markers) return fail(s, FG_EINVAL, "more markers than FgConfig.max_markers");
    if (!s->fish.empty()) return fail(s, FG_ESTATE, "markers are generated by fish bodies on this handle");
    if (s->cfg.n_ranks > 1 && n > 0) return fail(s, FG_ENOTSUP, "oracle: immersed boundary with n_ranks > 1 is not supported");
    // validate everything before touching the state: a rejected call leaves the previous marker set in place
    int nl = n > 0 ? 1 : 0;
    if (link)
        for (int k = 0; k < n; ++k) nl = std::max(nl, link[k] + 1);
    const int maxl = std::max(s->cfg.max_links, 1);
    if (nl > maxl) return fail(s, FG_EINVAL, "link id exceeds FgConfig.max_links");
    ensure_ib_storage(s);
    s->n_markers = n;
    s->mX.assign(X, X + 3 * size_t(n));
    s->mU.assign(U, U + 3 * size_t(n));
    s->mdV.assign(dV, dV + n);
    if (link) s->mlink.assign(link, link + n); else s->mlink.assign(n, 0);
    // torque reference points live in a table of max_links entries (as in the CUDA library): links without an
    // explicit origin refer to (0,0,0), and the link count never drops below what fg_set_link_origins announced
    s->link_origin.resize(3 * size_t(maxl), 0.0);
    s->n_links = std::max(nl, s->n_origins);
    s->wrench.assign(6 * size_t(s->n_links), 0.0);
    s->mbase.assign(3 * size_t(n), 0); s->mowner.assign(n, 0);
    s->mF.assign(3 * size_t(n), 0.0); s->mUstar.assign(3 * size_t(n), 0.0);
    return FG_OK;
}

int fg_set_link_origins(FgSim *s, int32_t n_links, const double *o) {
    if (!s || n_links < 0 || (n_links && !o)) return FG_EINVAL;
    const int maxl = std::max(s->cfg.max_links, 1);
    if (n_links > maxl) return fail(s, FG_EINVAL, "more links than FgConfig.max_links");
    s->link_origin.resize(3 * size_t(maxl), 0.0);
    std::copy(o, o + 3 * size_t(n_links), s->link_origin.begin());      // entries beyond n_links keep their values
    s->n_origins = n_links;
    if (n_links > s->n_links) { s->n_links = n_links; s->wrench.resize(6 * size_t(n_links), 0.0); }
    return FG_OK;
}

int fg_get_index_map(FgSim *s, int32_t *base3, int32_t *owner) {
    if (!s || !base3 || !owner) return FG_EINVAL;
    std::copy(s->mbase.begin(), s->mbase.begin() + 3 * size_t(s->n_markers), base3);
    std::copy(s->mowner.begin(), s->mowner.begin() + s->n_markers, owner);
    return FG_OK;
}
int fg_get_marker_forces(FgSim *s, float *F3) {
    if (!s || !F3) return FG_EINVAL;
    for (size_t i = 0; i < 3 * size_t(s->n_markers); ++i) F3[i] = float(s->mF[i]);
    return FG_OK;
}
int fg_get_marker_velocities(FgSim *s, float *U3) {
    if (!s || !U3) return FG_EINVAL;
    for (size_t i = 0; i < 3 * size_t(s->n_markers); ++i) U3[i] = float(s->mUstar[i]);
    return FG_OK;
}
int fg_get_link_wrenches(FgSim *s, double *w6) {
    if (!s || !w6) return FG_EINVAL;
    std::copy(s->wrench.begin(), s->wrench.begin() + 6 * size_t(s->n_links), w6);
    return FG_OK;
}
int fg_get_force_field(FgSim *s, float *F) {
    if (!s || !F) return FG_EINVAL;
    const size_t nloc = s->plane * s->nz;
    const bool have = s->Fx.size() == s->ncell;
    for (int z = 0; z < s->nz; ++z)
        for (int y = 0; y < s->ny; ++y)
            for (int x = 0; x < s->nx; ++x) {
                const size_t l = (size_t(z) * s->ny + y) * s->nx + x, c = s->idx(x, y, z);
                F[l] = have ? float(s->Fx[c]) : 0.f;
                F[nloc + l] = have ? float(s->Fy[c]) : 0.f;
                F[2 * nloc + l] = have ? float(s->Fz[c]) : 0.f;
            }
    return FG_OK;
}

int fg_probe(FgSim *s, int32_t n, const float *X, float *out4) {
    if (!s || n < 0 || (n && (!X || !out4))) return FG_EINVAL;
    if (int rc = finish_pending(s)) return rc;
    const FgConfig &c = s->cfg;
    const bool px = c.bc[FG_XLO] == FG_BC_PERIODIC, py = c.bc[FG_YLO] == FG_BC_PERIODIC, pz = c.bc[FG_ZLO] == FG_BC_PERIODIC;
    for (int k = 0; k < n; ++k) {
        const double Xk = X[3 * k], Yk = X[3 * k + 1], Zk = X[3 * k + 2];
        const int i0 = int(std::floor(X[3 * k])) - 1, j0 = int(std::floor(X[3 * k + 1])) - 1, k0 = int(std::floor(X[3 * k + 2])) - 1;
        double acc[4] = {0, 0, 0, 0};
        for (int cz = 0; cz < 4; ++cz) {
            const int zg = wrap_or_skip(k0 + cz, s->nzg, pz);
            if (zg < 0) continue;
            const int zl = zg - s->z0;
            if (zl < 0 || zl >= s->nz) continue;
            for (int cy = 0; cy < 4; ++cy) {
                const int yy = wrap_or_skip(j0 + cy, s->ny, py);
                if (yy < 0) continue;
                for (int cx = 0; cx < 4; ++cx) {
                    const int xx = wrap_or_skip(i0 + cx, s->nx, px);
                    if (xx < 0) continue;
                    const size_t cell = s->idx(xx, yy, zl);
                    double rho = 0, jx = 0, jy = 0, jz = 0;
                    for (int i = 0; i < Q; ++i) {
                        const double v = s->F(s->f, i)[cell];
                        rho += v; jx += CX[i] * v; jy += CY[i] * v; jz += CZ[i] * v;
                    }
                    const double w = peskin4(Xk - (i0 + cx)) * peskin4(Yk - (j0 + cy)) * peskin4(Zk - (k0 + cz));
                    acc[0] += w * rho; acc[1] += w * jx / rho; acc[2] += w * jy / rho; acc[3] += w * jz / rho;
                }
            }
        }
        for (int d = 0; d < 4; ++d) out4[4 * k + d] = float(acc[d]);
    }
    return FG_OK;
}

int fg_add_fish(FgSim *s, const FgFishDesc *d, int32_t *fish_id) {
    if (!s || !d) return FG_EINVAL;
    if (s->cfg.n_ranks > 1) return fail(s, FG_ENOTSUP, "oracle: bodies with n_ranks > 1 are not supported");
    obody::Fish fh;
    std::string why;
    if (!fh.init(*d, &why)) return fail(s, FG_EINVAL, why);
    fh.forcing_passes(s->cfg.ib_iterations);
    int n = fh.n_markers(), nl = fh.n_links();
    for (auto &o : s->fish) { n += o.n_markers(); nl += o.n_links(); }
    if (n > s->cfg.max_markers || nl > s->cfg.max_links)
        return fail(s, FG_EINVAL, "fish exceeds FgConfig.max_markers / max_links");
    s->fish.push_back(fh);
    int na = 0;
    for (auto &o : s->fish) na += o.n_joints();
    s->action.assign(na, 0.f);
    ensure_ib_storage(s);
    bodies_to_markers(s);
    if (fish_id) *fish_id = int32_t(s->fish.size()) - 1;
    return FG_OK;
}

int fg_action_size(FgSim *s) { return s ? int(s->action.size()) : FG_EINVAL; }
int fg_obs_size(FgSim *s) {
    if (!s) return FG_EINVAL;
    int n = 0;
    for (auto &o : s->fish) n += o.obs_size();
    return n;
}
int fg_set_action(FgSim *s, const float *a, int32_t n) {
    if (!s || (n && !a)) return FG_EINVAL;
    if (n != int32_t(s->action.size())) return fail(s, FG_EINVAL, "action length != fg_action_size()");
    for (int i = 0; i < n; ++i) s->action[i] = std::min(1.f, std::max(-1.f, a[i]));
    return FG_OK;
}
int fg_get_obs(FgSim *s, float *obs, int32_t n) {
    if (!s || !obs) return FG_EINVAL;
    if (n != fg_obs_size(s)) return fail(s, FG_EINVAL, "obs length != fg_obs_size()");
    int o = 0;
    for (auto &fh : s->fish) { fh.write_obs(obs + o); o += fh.obs_size(); }
    return FG_OK;
}
int fg_get_markers(FgSim *s, float *X, float *U, int32_t *link, int32_t cap) {
    if (!s) return FG_EINVAL;
    const int n = std::min(cap, s->n_markers);
    if (X) std::copy(s->mX.begin(), s->mX.begin() + 3 * size_t(n), X);
    if (U) std::copy(s->mU.begin(), s->mU.begin() + 3 * size_t(n), U);
    if (link) std::copy(s->mlink.begin(), s->mlink.begin() + n, link);
    return s->n_markers;
}

int fg_step(FgSim *s, int32_t n_substeps) {
    if (!s || n_substeps < 0) return FG_EINVAL;
    if (s->cfg.n_ranks > 1 && n_substeps != 1)
        return fail(s, FG_ESTATE, "oracle: with n_ranks > 1 step one substep at a time and exchange halos in between");
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < n_substeps; ++it) {
        if (int rc = finish_pending(s)) return rc;
        if (!s->fish.empty()) {
            // A7 (1): bodies advance with the wrenches of the previous substep
            int lo = 0, ao = 0;
            for (auto &fh : s->fish) {
                fh.advance(&s->action[ao], &s->wrench[6 * size_t(lo)], &s->link_origin[3 * size_t(lo)]);
                lo += fh.n_links(); ao += fh.n_joints();
            }
            bodies_to_markers(s);
        }
        if (s->n_markers > 0) ib_forces(s); else s->force_dirty = false;
        collide(s);
        s->force_dirty = false;
        s->stream_pending = true;
        if (s->n_internal_faces() == 0) stream(s);
        ++s->steps;
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    s->last_ms = ms;
    s->last_mlups = ms > 0 ? double(s->plane) * s->nz * n_substeps / ms / 1e3 : 0;
    return FG_OK;
}

int fg_sync(FgSim *s) { return s ? FG_OK : FG_EINVAL; }

int fg_get_stats(FgSim *s, FgStats *o) {
    if (!s || !o) return FG_EINVAL;
    std::memset(o, 0, sizeof(*o));
    o->steps = s->steps;
    o->cells = int64_t(s->plane) * s->nz;
    o->last_step_ms = s->last_ms;
    o->last_mlups = s->last_mlups;
    o->n_markers = s->n_markers; o->n_links = s->n_links; o->band_cells = s->band_cells;
    return FG_OK;
}

// include/fishgym.h fg_check_finite: rest population not finite or |f_0 - 1/3| > 4 (solid cells are not fluid)
int fg_check_finite(FgSim *s, int64_t *n_bad) {
    if (!s || !n_bad) return FG_EINVAL;
    int64_t bad = 0;
    for (int z = 1; z <= s->nz; ++z)
        for (size_t c = 0; c < s->plane; ++c) {
            const size_t i = size_t(z) * s->plane + c;
            if (s->has_solid && s->solid[i]) continue;
            const double v = s->f[i] - 1.0 / 3.0;
            if (!(std::fabs(v) <= 4.0)) ++bad;
        }
    *n_bad = bad;
    return FG_OK;
}

// include/fishgym.h fg_get_solid_force (SURVEY.md A5 "momentum-exchange force on solids").  A population that the pull rule
// reflected off an obstacle (pull(): solid[src]) left cell x with momentum -c_i f and came back with +c_i f: the obstacle took
// -2 c_i f_i(x, t+1).  Same source-cell rule as pull(): wall faces and the inlet are not obstacles, the outlet clamps the plane.
inline bool reflected_off_obstacle(const FgSim *s, int i, int x, int y, int z) {
    const FgConfig &c = s->cfg;
    int sx = x - CX[i], sy = y - CY[i], sz = z - CZ[i];
    if (sx < 0) { if (c.bc[FG_XLO] != FG_BC_PERIODIC) return false; sx = s->nx - 1; }
    else if (sx >= s->nx) { if (c.bc[FG_XHI] != FG_BC_PERIODIC) return false; sx = 0; }
    if (sy < 0) { if (c.bc[FG_YLO] != FG_BC_PERIODIC) return false; sy = s->ny - 1; }
    else if (sy >= s->ny) { if (c.bc[FG_YHI] != FG_BC_PERIODIC) return false; sy = 0; }
    const int zg = s->z0 + sz;
    if (zg < 0 || zg >= s->nzg) {
        switch (c.bc[zg < 0 ? FG_ZLO : FG_ZHI]) {
            case FG_BC_PERIODIC: if (c.n_ranks <= 1) sz = zg < 0 ? s->nz - 1 : 0; break;
            case FG_BC_OUTLET: sz = z; break;
            default: return false;      // wall, inlet
        }
    }
    return s->solid[s->idx(sx, sy, sz)] != 0;
}

int fg_get_solid_force(FgSim *s, const double *origin3, double *out6) {
    if (!s || !out6) return FG_EINVAL;
    if (int rc = finish_pending(s)) return rc;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
    if (s->has_solid) {
#pragma omp parallel for schedule(static) reduction(+ : a0, a1, a2, a3, a4, a5)
        for (int z = 0; z < s->nz; ++z)
            for (int y = 0; y < s->ny; ++y)
                for (int x = 0; x < s->nx; ++x) {
                    const size_t c = s->idx(x, y, z);
                    if (s->solid[c]) continue;
                    for (int i = 1; i < Q; ++i) {
                        if (!reflected_off_obstacle(s, i, x, y, z)) continue;
                        const double m = 2.0 * s->F(s->f, i)[c];
                        const double fx = -CX[i] * m, fy = -CY[i] * m, fz = -CZ[i] * m;
                        a0 += fx; a1 += fy; a2 += fz;
                        if (origin3) {
                            const double rx = x - 0.5 * CX[i] - origin3[0], ry = y - 0.5 * CY[i] - origin3[1],
                                         rz = (s->z0 + z) - 0.5 * CZ[i] - origin3[2];
                            a3 += ry * fz - rz * fy; a4 += rz * fx - rx * fz; a5 += rx * fy - ry * fx;
                        }
                    }
                }
    }
    out6[0] = a0; out6[1] = a1; out6[2] = a2; out6[3] = a3; out6[4] = a4; out6[5] = a5;
    return FG_OK;
}

int fg_set_flags(FgSim *s, int32_t flags) {
    if (!s) return FG_EINVAL;
    s->cfg.flags = flags;
    return FG_OK;
}

// ---- host-staged halos: ship post-collision f* of the boundary plane into the neighbour's ghost plane ----
int64_t fg_halo_bytes(FgSim *s) { return s ? int64_t(5 * s->plane * sizeof(double)) : FG_EINVAL; }

int fg_halo_pack(FgSim *s, int32_t face, void *vbuf) {
    double *buf = static_cast<double *>(vbuf);
    if (!s || !buf || (face != FG_ZLO && face != FG_ZHI)) return FG_EINVAL;
    if (!s->stream_pending) return fail(s, FG_ESTATE, "fg_halo_pack: call after fg_step");
    const int zl = face == FG_ZHI ? s->nz - 1 : 0;
    const int *set = face == FG_ZHI ? ZP : ZM;   // movers leaving through this face
    for (int q = 0; q < 5; ++q) {
        const double *src = s->F(s->fs, set[q]) + s->idx(0, 0, zl);
        std::copy(src, src + s->plane, buf + q * s->plane);
    }
    return FG_OK;
}

int fg_halo_unpack(FgSim *s, int32_t face, const void *vbuf) {
    const double *buf = static_cast<const double *>(vbuf);
    if (!s || !buf || (face != FG_ZLO && face != FG_ZHI)) return FG_EINVAL;
    if (!s->stream_pending) return fail(s, FG_ESTATE, "fg_halo_unpack: call after fg_step");
    const int zl = face == FG_ZHI ? s->nz : -1;
    const int *set = face == FG_ZHI ? ZM : ZP;   // movers entering through this face
    for (int q = 0; q < 5; ++q) {
        double *dst = s->F(s->fs, set[q]) + s->idx(0, 0, zl);
        std::copy(buf + q * s->plane, buf + (q + 1) * s->plane, dst);
    }
    if (++s->faces_received >= s->n_internal_faces()) stream(s);
    return FG_OK;
}

int fg_peer_export(FgSim *s, FgPeerHandle *) { return fail(s, FG_ENOTSUP, "oracle: no device peers; use fg_halo_pack/unpack"); }
int fg_peer_connect_all(FgSim *s, const FgPeerHandle *, int32_t) {
    return fail(s, FG_ENOTSUP, "oracle: no device peers; use fg_halo_pack/unpack");
}
int fg_peer_connect(FgSim *s, const FgPeerHandle *, const FgPeerHandle *) {
    return fail(s, FG_ENOTSUP, "oracle: no device peers; use fg_halo_pack/unpack");
}

}  // extern "C"
