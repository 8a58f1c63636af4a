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
    int *cellslot;               // [(nz+2)*plane]
    int *band_cell;              // [cap]
    float *band_u;               // [3][cap]
    float *bandF;                // [3][cap]
    int *band_count;             // [1]
    int band_cap;
    uint8_t *rowflag;            // [(nz+2)*ny]
    const float *origin;         // [links][3] torque reference points
    double *wrench;              // [links][6]
    int n_links;
};

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

// (a3) marker -> grid index map (integer outputs bit-exact with the oracle) + band registration
struct IbIndexMark {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kThreads + tx;
        if (k >= p.n) return;
        const Lattice &L = p.L;
        const float X = p.X[3 * k], Y = p.X[3 * k + 1], Z = p.X[3 * k + 2];
        const int i0 = int(floorf(X)) - 1, j0 = int(floorf(Y)) - 1, k0 = int(floorf(Z)) - 1;
        p.base[3 * k] = i0; p.base[3 * k + 1] = j0; p.base[3 * k + 2] = k0;
        int kc = wrap_or_skip(k0 + 1, L.nzg, p.per_z != 0);
        if (kc < 0) kc = k0 + 1 < 0 ? 0 : L.nzg - 1;
        p.owner[k] = kc / L.nz;
        for (int c = 0; c < 4; ++c)
            for (int b = 0; b < 4; ++b)
                for (int a = 0; a < 4; ++a) {
                    const long long cell = stencil_cell(p, i0, j0, k0, a, b, c);
                    if (cell < 0) continue;
                    if (p.cellslot[cell] != 0) continue;
                    if (atomic_cas_i(&p.cellslot[cell], 0, -1) == 0) {
                        const int pos = atomic_add_i(p.band_count, 1);
                        if (pos < p.band_cap) {
                            p.band_cell[pos] = int(cell);
                            p.cellslot[cell] = pos + 1;
                            p.rowflag[cell / L.nx] = 1;
                        }
                    }
                }
    }
};

// (a4) unforced velocity of the band cells from the populations arriving at time t (parity aware)
template <int PARITY>
struct IbBandMoments {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int pos = bx * kThreads + tx;
        const int cnt = *p.band_count < p.band_cap ? *p.band_count : p.band_cap;
        if (pos >= cnt) return;
        const Lattice &L = p.L;
        const long long idx = p.band_cell[pos];
        const int x = int(idx % L.nx), y = int((idx / L.nx) % L.ny), zz = int(idx / L.plane);
        const Nbr nb = make_nbr(L, x, y, zz);
        float h[Q];
        if (L.solid && L.solid[idx]) {
            FG_UNROLL
            for (int i = 0; i < Q; ++i) h[i] = L.f[i * L.slot + idx];   // obstacles keep their initial state (as in the oracle)
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
};

// (a5-a7) interpolate U*, direct forcing, spread.  One thread per marker in this first version.
struct IbInterpSpread {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int k = bx * kThreads + tx;
        if (k >= p.n) return;
        const float X = p.X[3 * k], Y = p.X[3 * k + 1], Z = p.X[3 * k + 2];
        const int i0 = p.base[3 * k], j0 = p.base[3 * k + 1], k0 = p.base[3 * k + 2];
        float wx[4], wy[4], wz[4];
        FG_UNROLL
        for (int a = 0; a < 4; ++a) {
            wx[a] = peskin4(X - float(i0 + a));
            wy[a] = peskin4(Y - float(j0 + a));
            wz[a] = peskin4(Z - float(k0 + a));
        }
        float us0 = 0.f, us1 = 0.f, us2 = 0.f;
        for (int c = 0; c < 4; ++c)
            for (int b = 0; b < 4; ++b)
                for (int a = 0; a < 4; ++a) {
                    const long long cell = stencil_cell(p, i0, j0, k0, a, b, c);
                    if (cell < 0) continue;
                    const int s = p.cellslot[cell] - 1;
                    if (s < 0) continue;
                    const float w = wx[a] * wy[b] * wz[c];
                    us0 += w * p.band_u[s]; us1 += w * p.band_u[p.band_cap + s]; us2 += w * p.band_u[2 * p.band_cap + s];
                }
        const float f0 = 2.0f * (p.U[3 * k] - us0), f1 = 2.0f * (p.U[3 * k + 1] - us1), f2 = 2.0f * (p.U[3 * k + 2] - us2);
        p.Ustar[3 * k] = us0; p.Ustar[3 * k + 1] = us1; p.Ustar[3 * k + 2] = us2;
        p.Fm[3 * k] = f0; p.Fm[3 * k + 1] = f1; p.Fm[3 * k + 2] = f2;
        const float dV = p.dV[k];
        for (int c = 0; c < 4; ++c)
            for (int b = 0; b < 4; ++b)
                for (int a = 0; a < 4; ++a) {
                    const long long cell = stencil_cell(p, i0, j0, k0, a, b, c);
                    if (cell < 0) continue;
                    const int s = p.cellslot[cell] - 1;
                    if (s < 0) continue;
                    const float w = wx[a] * wy[b] * wz[c] * dV;
                    atomic_add_f(&p.bandF[s], w * f0);
                    atomic_add_f(&p.bandF[p.band_cap + s], w * f1);
                    atomic_add_f(&p.bandF[2 * p.band_cap + s], w * f2);
                }
        // (a8) hydrodynamic wrench ON the link = minus what the markers exert on the fluid
        const int l = p.link[k];
        if (l >= 0 && l < p.n_links) {
            const double fx = -double(f0) * dV, fy = -double(f1) * dV, fz = -double(f2) * dV;
            const double rx = double(X) - double(p.origin[3 * l]), ry = double(Y) - double(p.origin[3 * l + 1]),
                         rz = double(Z) - double(p.origin[3 * l + 2]);
            double *w = p.wrench + 6 * l;
            atomic_add_d(w + 0, fx); atomic_add_d(w + 1, fy); atomic_add_d(w + 2, fz);
            atomic_add_d(w + 3, ry * fz - rz * fy);
            atomic_add_d(w + 4, rz * fx - rx * fz);
            atomic_add_d(w + 5, rx * fy - ry * fx);
        }
    }
};

// forget the band of the previous step
struct IbClearBand {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 4;
    FG_HD static void run(const IbParams &p, int bx, int, int, int tx) {
        const int pos = bx * kThreads + tx;
        const int cnt = *p.band_count < p.band_cap ? *p.band_count : p.band_cap;
        if (pos >= cnt) return;
        const long long idx = p.band_cell[pos];
        p.cellslot[idx] = 0;
        p.rowflag[idx / p.L.nx] = 0;
    }
};

// ---------------------------------------------------------------- host-side state
template <class Dev>
class IbState {
public:
    bool ready() const { return cap_ > 0; }
    int n_markers() const { return n_; }
    int n_links() const { return nl_; }
    int band_cells() const { return band_cells_; }
    const double *wrench_ptr() const { return h_wrench_.data(); }
    const double *origin_ptr() const { return h_origin_.data(); }
    void clear_wrenches() { std::fill(h_wrench_.begin(), h_wrench_.end(), 0.0); }

    int create(Dev &dev, const FgConfig &cfg, const Lattice &L, std::string &err) {
        cap_ = cfg.max_markers;
        maxl_ = std::max(cfg.max_links, 1);
        const long long cells = L.slot;
        band_cap_ = int(std::min<long long>(64ll * cap_, (long long)L.plane * L.nz));
        per_[0] = cfg.bc[FG_XLO] == FG_BC_PERIODIC; per_[1] = cfg.bc[FG_YLO] == FG_BC_PERIODIC; per_[2] = cfg.bc[FG_ZLO] == FG_BC_PERIODIC;
        auto A = [&](size_t bytes) { void *p = dev.alloc(bytes, err); if (p && !dev.zero(p, bytes)) { err = dev.err; p = nullptr; } return p; };
        dX_ = (float *)A(sizeof(float) * 3 * cap_); dU_ = (float *)A(sizeof(float) * 3 * cap_);
        ddV_ = (float *)A(sizeof(float) * cap_); dlink_ = (int *)A(sizeof(int) * cap_);
        dbase_ = (int *)A(sizeof(int) * 3 * cap_); downer_ = (int *)A(sizeof(int) * cap_);
        dF_ = (float *)A(sizeof(float) * 3 * cap_); dUs_ = (float *)A(sizeof(float) * 3 * cap_);
        cellslot_ = (int *)A(sizeof(int) * cells);
        band_cell_ = (int *)A(sizeof(int) * band_cap_);
        band_u_ = (float *)A(sizeof(float) * 3 * band_cap_);
        bandF_ = (float *)A(sizeof(float) * 3 * band_cap_);
        band_count_ = (int *)A(sizeof(int));
        rowflag_ = (uint8_t *)A(size_t(L.nz + 2) * L.ny);
        dorigin_ = (float *)A(sizeof(float) * 3 * maxl_);
        dwrench_ = (double *)A(sizeof(double) * 6 * maxl_);
        if (!dX_ || !dU_ || !ddV_ || !dlink_ || !dbase_ || !downer_ || !dF_ || !dUs_ || !cellslot_ || !band_cell_ || !band_u_ ||
            !bandF_ || !band_count_ || !rowflag_ || !dorigin_ || !dwrench_)
            return FG_ENOMEM;
        h_wrench_.assign(6 * size_t(maxl_), 0.0);
        h_origin_.assign(3 * size_t(maxl_), 0.0);
        return FG_OK;
    }

    void destroy(Dev &dev) {
        void *ps[] = {dX_, dU_, ddV_, dlink_, dbase_, downer_, dF_, dUs_, cellslot_, band_cell_, band_u_, bandF_, band_count_, rowflag_,
                      dorigin_, dwrench_};
        for (void *p : ps) dev.free(p);
        cap_ = 0;
    }

    int set_markers(Dev &dev, int n, const float *X, const float *U, const float *dV, const int32_t *link, std::string &err) {
        if (n > cap_) { err = "more markers than FgConfig.max_markers"; return FG_EINVAL; }
        int nl = 0;
        std::vector<int32_t> zero;
        if (!link) { zero.assign(n, 0); link = zero.data(); }
        for (int k = 0; k < n; ++k) nl = std::max(nl, link[k] + 1);
        if (nl > maxl_) { err = "link id exceeds FgConfig.max_links"; return FG_EINVAL; }
        bool ok = true;
        if (n > 0)
            ok = dev.h2d(dX_, X, sizeof(float) * 3 * n) && dev.h2d(dU_, U, sizeof(float) * 3 * n) &&
                 dev.h2d(ddV_, dV, sizeof(float) * n) && dev.h2d(dlink_, link, sizeof(int) * n);
        if (!ok) { err = dev.err; return FG_ECUDA; }
        n_ = n;
        nl_ = std::max(nl, nl_origins_);
        forces_valid_ = false;
        return FG_OK;
    }

    int set_link_origins(int n, const double *o, std::string &err) {
        if (n > maxl_) { err = "more links than FgConfig.max_links"; return FG_EINVAL; }
        for (int i = 0; i < 3 * n; ++i) h_origin_[i] = o[i];
        nl_origins_ = n;
        nl_ = std::max(nl_, n);
        origins_dirty_ = true;
        return FG_OK;
    }

    IbParams params(const Lattice &L, const Collision &C) const {
        IbParams p{};
        p.L = L; p.C = C; p.n = n_;
        p.per_x = per_[0]; p.per_y = per_[1]; p.per_z = per_[2];
        p.X = dX_; p.U = dU_; p.dV = ddV_; p.link = dlink_;
        p.base = dbase_; p.owner = downer_; p.Fm = dF_; p.Ustar = dUs_;
        p.cellslot = cellslot_; p.band_cell = band_cell_; p.band_u = band_u_; p.bandF = bandF_;
        p.band_count = band_count_; p.band_cap = band_cap_; p.rowflag = rowflag_;
        p.origin = dorigin_; p.wrench = dwrench_; p.n_links = nl_;
        return p;
    }

    // SURVEY.md A7 (2)-(5),(7) on the device; the collide that follows reads force_view()
    int compute_forces(Dev &dev, const Lattice &L, const Collision &C, int parity, std::string &err) {
        const IbParams p = params(L, C);
        const int bound = int(std::min<long long>(band_cap_, 64ll * std::max(n_prev_, 1)));
        bool ok = true;
        if (band_live_) ok = dev.template launch<IbClearBand>(Dim3x((bound + 127) / 128), p);
        ok = ok && dev.zero(band_count_, sizeof(int)) && dev.zero(dwrench_, sizeof(double) * 6 * maxl_);
        if (origins_dirty_) {
            std::vector<float> o(3 * size_t(maxl_));
            for (size_t i = 0; i < o.size(); ++i) o[i] = float(h_origin_[i]);
            ok = ok && dev.h2d(dorigin_, o.data(), sizeof(float) * o.size());
            origins_dirty_ = false;
        }
        const int nb = (n_ + 127) / 128;
        ok = ok && dev.template launch<IbIndexMark>(Dim3x(nb), p);
        const int bound2 = int(std::min<long long>(band_cap_, 64ll * n_));
        ok = ok && (parity == 0 ? dev.template launch<IbBandMoments<0>>(Dim3x((bound2 + 127) / 128), p)
                                : dev.template launch<IbBandMoments<1>>(Dim3x((bound2 + 127) / 128), p));
        ok = ok && dev.template launch<IbInterpSpread>(Dim3x(nb), p);
        if (!ok) { err = dev.err; return FG_ECUDA; }
        band_live_ = true; n_prev_ = n_; forces_valid_ = true; wrench_fetched_ = false;
        return FG_OK;
    }

    ForceField force_view() const { return ForceField{cellslot_, bandF_, band_cap_, rowflag_}; }
    int after_collide(Dev &, std::string &) { return FG_OK; }   // the band stays readable until the next compute_forces

    int fetch_wrenches(Dev &dev, std::string &err) {
        if (!forces_valid_ || wrench_fetched_) return FG_OK;
        if (!dev.sync() || !dev.d2h(h_wrench_.data(), dwrench_, sizeof(double) * 6 * maxl_)) { err = dev.err; return FG_ECUDA; }
        int cnt = 0;
        if (!dev.d2h(&cnt, band_count_, sizeof(int))) { err = dev.err; return FG_ECUDA; }
        if (cnt > band_cap_) { err = "IB band overflow"; return FG_ENOMEM; }
        band_cells_ = cnt;
        wrench_fetched_ = true;
        return FG_OK;
    }

    // read-outs for tests and observations
    int get_index_map(Dev &dev, int32_t *base3, int32_t *owner, std::string &err) {
        if (n_ == 0) return FG_OK;
        if (!dev.sync() || !dev.d2h(base3, dbase_, sizeof(int) * 3 * n_) || !dev.d2h(owner, downer_, sizeof(int) * n_)) { err = dev.err; return FG_ECUDA; }
        return FG_OK;
    }
    int get_marker_array(Dev &dev, bool forces, float *out, std::string &err) {
        if (n_ == 0) return FG_OK;
        if (!dev.sync() || !dev.d2h(out, forces ? dF_ : dUs_, sizeof(float) * 3 * n_)) { err = dev.err; return FG_ECUDA; }
        return FG_OK;
    }
    int get_force_field(Dev &dev, const Lattice &L, float *F, std::string &err) {
        const size_t nloc = size_t(L.plane) * L.nz;
        std::fill(F, F + 3 * nloc, 0.f);
        if (!band_live_) return FG_OK;
        int cnt = 0;
        if (!dev.sync() || !dev.d2h(&cnt, band_count_, sizeof(int))) { err = dev.err; return FG_ECUDA; }
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
        const int n = std::min(cap, n_);
        bool ok = true;
        if (n > 0 && X) ok = ok && dev.d2h(X, dX_, sizeof(float) * 3 * n);
        if (n > 0 && U) ok = ok && dev.d2h(U, dU_, sizeof(float) * 3 * n);
        if (n > 0 && link) ok = ok && dev.d2h(link, dlink_, sizeof(int) * n);
        if (!ok) { err = dev.err; return FG_ECUDA; }
        return n_;
    }

private:
    static Dim3 Dim3x(int x) { Dim3 d; d.x = std::max(x, 1); return d; }
    int cap_ = 0, maxl_ = 1, n_ = 0, nl_ = 0, nl_origins_ = 0, n_prev_ = 0, band_cap_ = 0, band_cells_ = 0;
    int per_[3] = {1, 1, 1};
    bool band_live_ = false, forces_valid_ = false, wrench_fetched_ = true, origins_dirty_ = true;
    float *dX_ = nullptr, *dU_ = nullptr, *ddV_ = nullptr, *dF_ = nullptr, *dUs_ = nullptr, *band_u_ = nullptr, *bandF_ = nullptr,
          *dorigin_ = nullptr;
    int *dlink_ = nullptr, *dbase_ = nullptr, *downer_ = nullptr, *cellslot_ = nullptr, *band_cell_ = nullptr, *band_count_ = nullptr;
    uint8_t *rowflag_ = nullptr;
    double *dwrench_ = nullptr;
    std::vector<double> h_wrench_, h_origin_;
};

}  // namespace fg
