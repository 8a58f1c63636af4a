// sim.hpp — host runtime of the coupled step: handle state, step orchestration, z-slab halos.
//
// SURVEY.md §3b call stack (env.step -> fg_step -> per-substep launches); nothing in /root/reference to
// mirror (README.md only).  Templated on a device policy:
//   CudaDev (dev_cuda.cuh)   the product — CUDA streams/events, sm_100a kernels, CUDA-IPC peers
//   HostDev (tests/emu)      test infrastructure — runs the same kernel bodies in a CPU loop
#pragma once
#include "../../include/fishgym.h"
#include "lbm_core.cuh"
#include "ib_core.cuh"
#include "body.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace fg {

// what one rank publishes to its z-neighbours (fits FgPeerHandle.bytes)
struct PeerBlob {
    uint64_t magic;
    int32_t pid, device;
    uint64_t f_ptr, flag_ptr, x_ptr;   // valid inside the publishing process (x: IB exchange buffer, may be 0)
    unsigned char f_ipc[64], flag_ipc[64], x_ipc[64];
    int32_t nx, ny, nz;
    int32_t rank, n_ranks;
    uint64_t x_bytes;
};
static_assert(sizeof(PeerBlob) <= sizeof(FgPeerHandle), "PeerBlob must fit the ABI blob");
constexpr uint64_t kPeerMagic = 0x4647504545523034ull + kPopBytes;   // "FGPEER04" + bytes per population (f32 and f16 builds do not mix)

template <class Dev>
class SimT {
public:
    std::string err;
    FgConfig cfg{};
    Dev dev;

    // ------------------------------------------------------------------ lifecycle
    int create(const FgConfig &c) {
        cfg = c;
        if (cfg.inlet_rho == 0) cfg.inlet_rho = 1.0;
        nzl_ = cfg.nz / cfg.n_ranks;
        if (nzl_ < 2) return fail(FG_EINVAL, "each z-slab needs at least 2 planes");
        if (!dev.init(cfg.device, err)) return FG_ECUDA;
        L_ = Lattice{};
        L_.nx = cfg.nx; L_.ny = cfg.ny; L_.nz = nzl_;
        const long long plane = (long long)cfg.nx * cfg.ny;
        if (plane >= (1ll << 30) || plane * (nzl_ + 2) >= (1ll << 31))
            return fail(FG_EINVAL, "slab too large for 32-bit cell indices");
        L_.plane = int(plane);
        L_.slot = plane * (nzl_ + 2);
        L_.z0 = cfg.rank * nzl_; L_.nzg = cfg.nz;
        L_.wall_x = cfg.bc[FG_XLO] == FG_BC_WALL; L_.wall_y = cfg.bc[FG_YLO] == FG_BC_WALL;
        L_.bc_zlo = cfg.bc[FG_ZLO]; L_.bc_zhi = cfg.bc[FG_ZHI];
        L_.solid = nullptr;
        L_.f = static_cast<pop_t *>(dev.alloc(size_t(Q) * L_.slot * sizeof(pop_t), err));
        if (!L_.f) return FG_ENOMEM;
        flags_ = static_cast<int *>(dev.alloc(4 * sizeof(int), err));
        if (!flags_) return FG_ENOMEM;
        if (!dev.zero(flags_, 4 * sizeof(int))) return cuda_fail();
        pair_ctr_ = static_cast<int *>(dev.alloc(sizeof(int) * size_t(L_.nz + 3), err));
        if (!pair_ctr_) return FG_ENOMEM;
        if (!dev.zero(pair_ctr_, sizeof(int) * size_t(L_.nz + 3))) return cuda_fail();
        setup_collision();
        if (cfg.max_markers > 0) {
            if (int rc = ib_.create(dev, cfg, L_, err)) return rc;
        }
        return reset(0);
    }

    void destroy() {
        dev.enter();
        ib_.destroy(dev);
        dev.close_peers();
        dev.free(L_.f); L_.f = nullptr;
        dev.free(flags_); flags_ = nullptr;
        dev.free(pair_ctr_); pair_ctr_ = nullptr;
        dev.free(solid_); solid_ = nullptr;
        dev.free(stage_[0]); dev.free(stage_[1]); stage_[0] = stage_[1] = nullptr;
        dev.free(probe_d_); dev.free_host(probe_h_); probe_d_ = probe_h_ = nullptr; probe_cap_ = 0;
        dev.free(finite_d_); finite_d_ = nullptr;
        dev.free(solid_force_d_); solid_force_d_ = nullptr;
        dev.shutdown();
    }

    int reset(uint64_t) {
        InitParams p{L_, nullptr, nullptr};
        if (!dev.template launch<InitEquilibrium>(grid_planes(L_.nz + 2), p)) return cuda_fail();
        parity_ = 0; steps_ = 0;
        pending_faces_ = 0;     // the ghost planes were rebuilt: no halo of an earlier step is owed any more
        for (auto &f : fish_) f.reset();
        std::fill(action_.begin(), action_.end(), 0.f);
        if (!fish_.empty()) {
            if (int rc = bodies_to_markers()) return rc;
            ib_.clear_wrenches();
        }
        if (!dev.sync()) return cuda_fail();
        return FG_OK;
    }

    // ------------------------------------------------------------------ fluid state
    int set_fields(const float *rho, const float *u) {
        const size_t n = size_t(L_.plane) * L_.nz;
        float *d = static_cast<float *>(dev.alloc(4 * n * sizeof(float), err));
        if (!d) return FG_ENOMEM;
        bool ok = dev.h2d(d, rho, n * sizeof(float)) && dev.h2d(d + n, u, 3 * n * sizeof(float));
        InitParams p{L_, d, d + n};
        ok = ok && dev.template launch<InitEquilibrium>(grid_planes(L_.nz + 2), p) && dev.sync();
        dev.free(d);
        if (!ok) return cuda_fail();
        parity_ = 0;
        pending_faces_ = 0;
        return FG_OK;
    }

    // Read-outs gather the populations ARRIVING at each cell; after an even step the boundary planes pull them from the
    // ghost planes, i.e. from what the z-neighbours pushed after THEIR last step.  Host-staged slabs must have unpacked
    // every face (as in the oracle); peered slabs wait for the neighbours' flags on the device — a neighbour may lag by a
    // whole step (another process, another GPU), and without the wait the read-out raced with its push (found in round 2
    // by running the two-process test on ONE device, where time slicing makes the lag the rule).
    int halos_ready_for_readout() {
        if (cfg.n_ranks <= 1 || parity_ == 0) return FG_OK;
        if (!peers_) {
            if (pending_faces_ > 0) return fail(FG_ESTATE, "halo exchange incomplete: unpack every internal face after fg_step");
            return FG_OK;
        }
        if (!dev.wait_flags(flags_, has_lo_peer(), has_hi_peer()) || !dev.sync()) return cuda_fail();
        return check_peer_timeout();
    }

    int get_moments(std::vector<float> &mom) {
        if (int rc = halos_ready_for_readout()) return rc;
        const size_t n = size_t(L_.plane) * L_.nz;
        float *d = static_cast<float *>(dev.alloc(4 * n * sizeof(float), err));
        if (!d) return FG_ENOMEM;
        GatherParams p{L_, C_, nullptr, d};
        bool ok = parity_ == 0 ? dev.template launch<GatherArriving<0>>(grid_planes(L_.nz), p)
                               : dev.template launch<GatherArriving<1>>(grid_planes(L_.nz), p);
        mom.resize(4 * n);
        ok = ok && dev.sync() && dev.d2h(mom.data(), d, 4 * n * sizeof(float));
        dev.free(d);
        return ok ? FG_OK : cuda_fail();
    }

    int get_fields(float *rho, float *u) {
        std::vector<float> mom;
        if (int rc = get_moments(mom)) return rc;
        const size_t n = size_t(L_.plane) * L_.nz;
        for (size_t i = 0; i < n; ++i) rho[i] = 1.0f + mom[i];
        std::memcpy(u, mom.data() + n, 3 * n * sizeof(float));
        return FG_OK;
    }
    int get_fields_f64(double *rho, double *u) {
        std::vector<float> mom;
        if (int rc = get_moments(mom)) return rc;
        const size_t n = size_t(L_.plane) * L_.nz;
        for (size_t i = 0; i < n; ++i) rho[i] = 1.0 + double(mom[i]);   // rho-1 travels in fp32, the 1 is added in fp64
        for (size_t i = 0; i < 3 * n; ++i) u[i] = double(mom[n + i]);
        return FG_OK;
    }

    int set_populations(const float *f19) {
        const size_t n = size_t(L_.plane) * L_.nz;
        std::vector<pop_t> h(n);
        for (int i = 0; i < Q; ++i) {
            for (size_t k = 0; k < n; ++k) pop_st(&h[k], float(double(f19[i * n + k]) - WD[i]));
            if (!dev.h2d(L_.f + i * L_.slot + L_.plane, h.data(), n * sizeof(pop_t))) return cuda_fail();
        }
        parity_ = 0;
        pending_faces_ = 0;
        return FG_OK;
    }

    int get_populations(float *f19) {
        if (int rc = halos_ready_for_readout()) return rc;
        const size_t n = size_t(L_.plane) * L_.nz;
        float *d = static_cast<float *>(dev.alloc(size_t(Q) * n * sizeof(float), err));
        if (!d) return FG_ENOMEM;
        GatherParams p{L_, C_, d, nullptr};
        bool ok = parity_ == 0 ? dev.template launch<GatherArriving<0>>(grid_planes(L_.nz), p)
                               : dev.template launch<GatherArriving<1>>(grid_planes(L_.nz), p);
        ok = ok && dev.sync() && dev.d2h(f19, d, size_t(Q) * n * sizeof(float));
        dev.free(d);
        if (!ok) return cuda_fail();
        for (int i = 0; i < Q; ++i)
            for (size_t k = 0; k < n; ++k) f19[i * n + k] = float(double(f19[i * n + k]) + WD[i]);
        return FG_OK;
    }

    // After an even step the streaming of that step is still pending (the next odd step performs it while reading, with
    // the bounce-back links of the mask it sees THEN).  Whatever changes the mask must therefore first complete the
    // streaming with the old one: gather the arriving populations and store them back in natural layout (parity 0).
    int finish_pending_streaming() {
        if (parity_ == 0) return FG_OK;
        if (cfg.n_ranks > 1 && !peers_ && pending_faces_ > 0)
            return fail(FG_ESTATE, "halo exchange incomplete: unpack every internal face after fg_step");
        // on peered slabs the ghost planes must hold the neighbours' halos of the last step (all ranks call this in step)
        if (peers_ && !dev.wait_flags(flags_, has_lo_peer(), has_hi_peer())) return cuda_fail();
        const size_t n = size_t(L_.plane) * L_.nz;
        float *d = static_cast<float *>(dev.alloc(size_t(Q) * n * sizeof(float), err));
        if (!d) return FG_ENOMEM;
        GatherParams p{L_, C_, d, nullptr};
        const bool ok = dev.template launch<GatherArriving<1>>(grid_planes(L_.nz), p) &&
                        dev.template launch<ScatterNatural>(grid_planes(L_.nz), p) && dev.sync();
        dev.free(d);
        if (!ok) return cuda_fail();
        parity_ = 0;
        return FG_OK;
    }

    int set_solid(const uint8_t *g) {
        dev.graph_clear();
        if (int rc = finish_pending_streaming()) return rc;
        if (!g) {
            dev.free(solid_); solid_ = nullptr; L_.solid = nullptr;
            return FG_OK;
        }
        const size_t tot = size_t(L_.plane) * (L_.nz + 2);
        std::vector<uint8_t> h(tot, 0);
        const bool pz = cfg.bc[FG_ZLO] == FG_BC_PERIODIC;
        bool any = false;
        for (int zl = -1; zl <= L_.nz; ++zl) {
            int zg = L_.z0 + zl;
            if (zg < 0 || zg >= L_.nzg) {
                const int bc = zg < 0 ? cfg.bc[FG_ZLO] : cfg.bc[FG_ZHI];
                if (pz) zg = (zg + L_.nzg) % L_.nzg;
                else if (bc == FG_BC_OUTLET) zg = zg < 0 ? 0 : L_.nzg - 1;   // zero-gradient: the pull source is clamped (oracle pull())
                else continue;
            }
            for (int p = 0; p < L_.plane; ++p) {
                const uint8_t v = g[size_t(zg) * L_.plane + p] ? 1 : 0;
                h[size_t(zl + 1) * L_.plane + p] = v;
                any = any || v;
            }
        }
        if (!any) return set_solid(nullptr);
        if (!solid_) solid_ = static_cast<uint8_t *>(dev.alloc(tot, err));
        if (!solid_) return FG_ENOMEM;
        if (!dev.h2d(solid_, h.data(), tot)) return cuda_fail();
        L_.solid = solid_;
        return FG_OK;
    }

    // ------------------------------------------------------------------ immersed boundary / bodies
    int set_markers(int n, const float *X, const float *U, const float *dV, const int32_t *link) {
        if (n > cfg.max_markers) return fail(FG_EINVAL, "more markers than FgConfig.max_markers");
        if (!fish_.empty()) return fail(FG_ESTATE, "markers are generated by fish bodies on this handle");
        if (cfg.n_ranks > 1 && !ib_.exchange_on()) {
            // without fg_peer_connect_all there is no marker exchange: every 4-wide stencil must lie inside this rank's planes
            for (int k = 0; k < n; ++k) {
                const int k0 = int(std::floor(X[3 * k + 2])) - 1;
                if (k0 < L_.z0 || k0 + 3 >= L_.z0 + L_.nz)
                    return fail(FG_ENOTSUP, "a marker stencil crosses a z-slab face: connect all ranks with fg_peer_connect_all first");
            }
        }
        if (n > 0 && !ib_.ready()) return fail(FG_ESTATE, "FgConfig.max_markers was 0 at create");
        if (!ib_.ready()) return FG_OK;
        return ib_.set_markers(dev, n, X, U, dV, link, nullptr, 0, err);
    }
    int set_link_origins(int n, const double *o) {
        if (!ib_.ready()) return fail(FG_ESTATE, "FgConfig.max_markers was 0 at create");
        return ib_.set_link_origins(dev, n, o, err);
    }
    IbState<Dev> &ib() { return ib_; }
    int get_force_field(float *F) {
        if (!ib_.ready()) {
            std::fill(F, F + 3 * size_t(L_.plane) * L_.nz, 0.f);
            return FG_OK;
        }
        return ib_.get_force_field(dev, L_, F, err);
    }

    int probe(int n, const float *X, float *out4) {
        if (n == 0) return FG_OK;
        if (int rc = halos_ready_for_readout()) return rc;
        // persistent, grow-only buffers: the call sits in the per-step loop of an env with velocity probes, and a
        // cudaMalloc / cudaFree pair per call would synchronise the device every step
        if (n > probe_cap_) {
            dev.free(probe_d_); dev.free_host(probe_h_);
            probe_cap_ = 0;
            const int cap = std::max(64, n);
            probe_d_ = static_cast<float *>(dev.alloc(sizeof(float) * 7 * size_t(cap), err));
            probe_h_ = static_cast<float *>(dev.alloc_host(sizeof(float) * 7 * size_t(cap), err));
            if (!probe_d_ || !probe_h_) return FG_ENOMEM;
            probe_cap_ = cap;
        }
        float *d = probe_d_;
        const size_t N = size_t(n);
        std::memcpy(probe_h_, X, sizeof(float) * 3 * N);
        ProbeParams p{L_, C_, n, cfg.bc[FG_XLO] == FG_BC_PERIODIC, cfg.bc[FG_YLO] == FG_BC_PERIODIC, cfg.bc[FG_ZLO] == FG_BC_PERIODIC, d, d + 3 * N};
        Dim3 g{(n + kMarkersPerCta - 1) / kMarkersPerCta, 1, 1};
        bool ok = dev.h2d_async(d, probe_h_, sizeof(float) * 3 * N) && dev.zero(d + 3 * N, sizeof(float) * 4 * N);
        ok = ok && (parity_ == 0 ? dev.template launch<ProbeMoments<0>>(g, p) : dev.template launch<ProbeMoments<1>>(g, p));
        ok = ok && dev.d2h_async(probe_h_ + 3 * N, d + 3 * N, sizeof(float) * 4 * N) && dev.sync();
        if (!ok) return cuda_fail();
        std::memcpy(out4, probe_h_ + 3 * N, sizeof(float) * 4 * N);
        return FG_OK;
    }

    // fg_check_finite: cells whose rest population is not finite / out of range (lbm_core.cuh FiniteCheck)
    int check_finite(int64_t *n_bad) {
        if (!finite_d_) finite_d_ = static_cast<unsigned long long *>(dev.alloc(sizeof(unsigned long long), err));
        if (!finite_d_) return FG_ENOMEM;
        FiniteParams p{L_, finite_d_};
        unsigned long long h = 0;
        const bool ok = dev.zero(finite_d_, sizeof(unsigned long long)) && dev.template launch<FiniteCheck>(grid_planes(L_.nz), p) &&
                        dev.sync() && dev.d2h(&h, finite_d_, sizeof(h));
        if (!ok) return cuda_fail();
        *n_bad = int64_t(h);
        return FG_OK;
    }

    // include/fishgym.h fg_get_solid_force: momentum exchange over the bounce-back links of the obstacle cells (lbm_core.cuh SolidForce)
    int get_solid_force(const double *origin3, double *out6) {
        std::fill(out6, out6 + 6, 0.0);
        if (!L_.solid) return FG_OK;
        if (int rc = halos_ready_for_readout()) return rc;
        if (!solid_force_d_) solid_force_d_ = static_cast<double *>(dev.alloc(6 * sizeof(double), err));
        if (!solid_force_d_) return FG_ENOMEM;
        SolidForceParams p{L_, C_, solid_force_d_, origin3 ? origin3[0] : 0.0, origin3 ? origin3[1] : 0.0, origin3 ? origin3[2] : 0.0, origin3 ? 1 : 0};
        bool ok = dev.zero(solid_force_d_, 6 * sizeof(double));
        ok = ok && (parity_ == 0 ? dev.template launch<SolidForce<0>>(grid_planes(L_.nz), p) : dev.template launch<SolidForce<1>>(grid_planes(L_.nz), p));
        ok = ok && dev.sync() && dev.d2h(out6, solid_force_d_, 6 * sizeof(double));
        return ok ? FG_OK : cuda_fail();
    }

    int add_fish(const FgFishDesc &d, int32_t *id) {
        if (cfg.n_ranks > 1 && !ib_.exchange_on())
            return fail(FG_ENOTSUP, "bodies on z-slabs: connect all ranks with fg_peer_connect_all first (and add the same fish on every rank)");
        if (!ib_.ready()) return fail(FG_ESTATE, "FgConfig.max_markers was 0 at create");
        Fish f;
        if (!f.init(d, err)) return FG_EINVAL;
        f.set_forcing_passes(ib_.iterations());
        int n = f.n_markers(), nl = f.n_links();
        for (auto &o : fish_) { n += o.n_markers(); nl += o.n_links(); }
        if (n > cfg.max_markers || nl > cfg.max_links) return fail(FG_EINVAL, "fish exceeds FgConfig.max_markers / max_links");
        fish_.push_back(f);
        int na = 0;
        for (auto &o : fish_) na += o.n_joints();
        action_.assign(na, 0.f);
        if (int rc = bodies_to_markers()) return rc;
        if (id) *id = int(fish_.size()) - 1;
        return FG_OK;
    }
    int action_size() const { return int(action_.size()); }
    int obs_size() const {
        int n = 0;
        for (auto &f : fish_) n += f.obs_size();
        return n;
    }
    int set_action(const float *a, int n) {
        if (n != action_size()) return fail(FG_EINVAL, "action length != fg_action_size()");
        for (int i = 0; i < n; ++i) action_[i] = std::min(1.f, std::max(-1.f, a[i]));
        return FG_OK;
    }
    int get_obs(float *o, int n) {
        if (n != obs_size()) return fail(FG_EINVAL, "obs length != fg_obs_size()");
        int k = 0;
        for (auto &f : fish_) { f.write_obs(o + k); k += f.obs_size(); }
        return FG_OK;
    }

    // ------------------------------------------------------------------ stepping
    int step(int n) {
        if (n < 0) return fail(FG_EINVAL, "n_substeps < 0");
        const bool ranks = cfg.n_ranks > 1;
        if (ranks && !peers_ && n != 1)
            return fail(FG_ESTATE, "z-slabs without device peers: step one substep at a time and exchange halos (fg_halo_pack/unpack)");
        if (ranks && !peers_ && pending_faces_ > 0)
            return fail(FG_ESTATE, "halo exchange incomplete: unpack every internal face after fg_step");
        const bool prof = (cfg.flags & FG_FLAG_PROFILE) != 0;
        const bool graphs = !prof && !(cfg.flags & FG_FLAG_NO_GRAPHS);
        Range step_range(dev, "fg_step");
        dev.marks_reset();
        dev.tic();
        for (int it = 0; it < n; ++it) {
            if (!fish_.empty()) {
                // A7 (1): bodies advance on the host with the wrenches of the previous substep
                if (int rc = advance_bodies()) return rc;
            }
            // The device work of one substep is a fixed sequence for a given (parity, IB counter, staging buffer,
            // marker count, plane split): captured once into a CUDA graph and replayed afterwards (the host code below
            // still runs on replay to advance its state, the device policy just does not enqueue).
            if (!fish_.empty()) emit_bodies();      // host only: marker coordinates of this substep (and their z-range)
            const bool ib_on = ib_.ready() && ib_.n_markers() > 0;
            const bool overlap = ranks && peers_ && !(cfg.flags & FG_FLAG_NO_OVERLAP) && L_.nz >= 4;
            // planes the halo push does not wait for: all of them, or the interior ones when boundary planes go first
            const int lo = overlap ? 2 : 1, hi = overlap ? L_.nz : L_.nz + 1;
            // Plane split: only planes [near_a_, near_b_) hold cells the IB kernels read or write (IbState::near_planes).
            // The collide of the other ("far") planes needs neither the force nor the marker upload, so it runs on a
            // low-priority branch BESIDE the IB kernels instead of after them.
            split_ = false; near_a_ = lo; near_b_ = hi;
            // Halo branch (peered slabs with bodies): when no stencil reaches a boundary plane, the boundary planes need no IB
            // force, so boundary planes -> halo push -> signal run on their own high-priority stream (4, thin rows on 5) BESIDE
            // the IB kernels instead of between them and the near-plane collide.  Default on slabs below 8 M cells, where that
            // chain is as long as the interior collide beside it; FG_HALO_BRANCH=1 forces it on every slab, FG_NO_HALO_BRANCH=1
            // turns it off.  Two GPUs, one 256x128x128 channel + sphere each, 2000 steps (gpu pass b10): serial order 71 650
            // MLUPS (0.92 of two single GPUs); branch 75 200 with the one-cell kernels and 77 400 - 78 000 with the two-cell
            // ones (0.99 - 1.00), which such a slab therefore runs while the branch is active (small_peered_slab()).
            // Bit-identical either way (tests/test_slabs.py toggles it under every stream-order policy of the emulation).
            bool halo_branch = false;
            if (ib_on && !prof && !(cfg.flags & FG_FLAG_NO_SPLIT)) {
                int a, b;
                if (ib_.near_planes(a, b)) {
                    halo_branch = overlap && ib_.boundary_planes_free(L_.nz) && (halo_branch_ == 1 || (halo_branch_ < 0 && small_slab()));
                    a = std::min(std::max(a, lo), hi); b = std::max(std::min(b, hi), a);
                    // worth it when the far planes are at least a quarter of the slab and ~25 us of work (1 M cells)
                    const int far = (a - lo) + (hi - b);
                    if (far * 4 >= hi - lo && (long long)far * L_.plane >= (cfg.split_min_cells > 0 ? cfg.split_min_cells : (1 << 20))) { split_ = true; near_a_ = a; near_b_ = b; }
                }
            }
            halo_branch_now_ = halo_branch;     // part of the substep's graph key: it changes the captured sequence
            // Peered z-slabs WITHOUT bodies: nothing waits for a force, so every interior plane is "far" — the interior
            // collide goes to the low-priority branch and the high-priority chain (wait for the neighbours' flags ->
            // boundary planes -> halo push -> signal) runs BESIDE it instead of in front of it.  In round 1 that chain
            // (~10 us of latency-bound launches) sat serially before the interior kernel in the substep's graph and was
            // the whole weak-scaling loss of small slabs.  Interior planes touch no location of
            // the boundary planes' update (the AA update of plane p owns plane p-1's -z slots, plane p+1's +z slots
            // and its own rest), so the two branches are independent.
            if (!ib_on && overlap && !prof && !(cfg.flags & FG_FLAG_NO_SPLIT) &&
                (long long)(hi - lo) * L_.plane >= (cfg.split_min_cells > 0 ? cfg.split_min_cells : (1 << 20))) {
                split_ = true; near_a_ = near_b_ = lo;
            }
            // Two substeps in one wavefront launch (StreamCollidePair): opt-in, measured slower than two plain launches
            // (lbm_core.cuh); no bodies, one rank.
            if (!prof && (cfg.flags & FG_FLAG_FUSED_PAIRS) && parity_ == 0 && it + 1 < n && !ib_on && !ranks && !L_.solid && L_.nz >= 8) {
                struct Scope {
                    Dev &d; bool on; bool done = false;
                    ~Scope() { if (on && !done) d.graph_abort(); }
                } scope{dev, graphs && dev.graph_begin(GraphKey{{0x50414952ull, 0, 0, 0}})};
                int done[2][2];
                if (!launch_pair(1, L_.nz + 1, 0, 0, done) || !launch_faces()) return cuda_fail();
                parity_ = 1;
                // the planes at the range ends (their z-neighbours are ghost planes, set by the face operations above)
                if (!launch_collide_except(1, L_.nz + 1, ForceField{}, done) || !launch_faces()) return cuda_fail();
                parity_ = 0;
                steps_ += 2; pair_substeps_ += 2;
                ++it;
                if (scope.on) {
                    scope.done = true;
                    if (!dev.graph_end()) return cuda_fail();
                }
                continue;
            }
            // A substep with a far branch is launched kernel by kernel: inside a captured graph the two priorities had no
            // effect on B200 (r1 pass 8/9: 0.137 ms/step as a graph with either the stream priorities, per-node priority
            // attributes + cudaGraphInstantiateFlagUseNodePriority, or the far branch outside the graph; 0.1125 ms direct),
            // and with the IB chain hidden beside the far collide its launch latencies no longer matter.
            struct Scope {
                Dev &d; bool on; bool done = false;
                ~Scope() { if (on && !done) d.graph_abort(); }
            } scope{dev, graphs && !split_ && dev.graph_begin(substep_key())};
            // without the boundary-first order the far branch contains the boundary planes, which do read what the
            // neighbours pushed in the previous step
            if (ranks && peers_ && !overlap && !dev.wait_flags(flags_, has_lo_peer(), has_hi_peer())) return cuda_fail();
            if (split_) {
                ++split_substeps_;
                if (!dev.fork_to(2) || !launch_collide(lo, hi, ForceField{}, 1, near_a_, near_b_) || !dev.switch_to(0)) return cuda_fail();
            }
            if (!fish_.empty()) {
                if (int rc = upload_bodies()) return rc;
            }
            // neighbours must have delivered the halos of the previous step before anything reads ghost planes
            // (the IB band moments do, at odd parity) or boundary planes
            if (ranks && peers_ && overlap && !dev.wait_flags(flags_, has_lo_peer(), has_hi_peer())) return cuda_fail();
            // the branch starts here (behind the neighbour wait), but its kernels are SUBMITTED after the IB kernels: streams of
            // equal priority are served in submission order, and the IB kernels head the longer chain
            if (halo_branch && (!dev.fork_to(4) || !dev.switch_to(0))) return cuda_fail();
            if (halo_branch && halo_first_ && (!dev.switch_to(4) || !launch_boundary_planes(ForceField{}) || !launch_faces() || !dev.switch_to(0))) return cuda_fail();
            ForceField F{};
            if (ib_on) {
                if (prof) dev.mark(1);
                ib_.set_fused((cfg.flags & FG_FLAG_FUSED_IB) != 0);
                ib_.set_tile_spread((cfg.flags & FG_FLAG_IB_TILE_SPREAD) != 0);
                {
                    Range r(dev, "fg:immersed_boundary");
                    if (int rc = ib_.compute_forces(dev, L_, C_, parity_, err)) return rc;
                }
                if (prof) dev.mark(1);
                F = ib_.force_view();
            }
            if (halo_branch && !halo_first_ && (!dev.switch_to(4) || !launch_boundary_planes(ForceField{}) || !launch_faces() || !dev.switch_to(0))) return cuda_fail();
            if (overlap && !halo_branch) {
                // boundary planes first, push halos over NVLink, then the interior hides the exchange
                if (!launch_boundary_planes(F) || !launch_faces()) return cuda_fail();
            }
            if (!launch_collide(near_a_, near_b_, F)) return cuda_fail();
            if (split_ && !dev.join_from(2)) return cuda_fail();
            if (halo_branch && !dev.join_from(4)) return cuda_fail();
            if (!overlap && !launch_faces()) return cuda_fail();
            parity_ ^= 1;
            ++steps_;
            if (scope.on) {
                scope.done = true;
                if (!dev.graph_end()) return cuda_fail();
            }
        }
        // The call returns once what it exposes synchronously is there: the link wrenches (event recorded right after the
        // IB kernels of the last substep) and, with bodies, the host-integrated observation.  The collide of the last
        // substep may still be running; every later call is ordered behind it on the handle's stream, fg_sync() waits for
        // it, and the timing of the call is read when fg_get_stats() finds it complete.  A caller that feeds markers /
        // actions step by step thereby overlaps its own work and the next upload with the fluid kernel.
        dev.toc_record();
        timing_pending_ = true; timed_substeps_ = n;
        const bool wait_all = prof || (cfg.flags & FG_FLAG_SYNC_STEP);
        if (wait_all) {
            if (!dev.sync()) return cuda_fail();
            finish_timing(true);
        }
        if (prof) {
            collide_ms_ = dev.marks_elapsed(0); ib_ms_ = dev.marks_elapsed(1);
            last_collide_launches_ = collide_launches_; last_collide_cells_ = collide_cells_;
        }
        collide_launches_ = 0; collide_cells_ = 0;
        if (int rc = check_peer_timeout()) return rc;
        if (ib_.ready() && ib_.n_markers() > 0) {
            if (int rc = ib_.fetch_wrenches(dev, err)) return rc;
        }
        if (ranks && !peers_) pending_faces_ = int(internal_lo()) + int(internal_hi());
        return FG_OK;
    }

    // A7 (1): wait for the wrenches of the previous substep's IB pass, integrate every fish one substep on the host
    int advance_bodies() {
        if (int rc = ib_.fetch_wrenches(dev, err)) return rc;
        int lo = 0, ao = 0;
        for (auto &f : fish_) {
            f.advance(&action_[ao], ib_.wrench_ptr() + 6 * lo, ib_.origin_ptr() + 3 * lo);
            lo += f.n_links(); ao += f.n_joints();
        }
        return FG_OK;
    }

    // the neighbour-wait kernels raise a pinned host word when they give up (ranks out of step): seen here without a sync,
    // possibly one call late — the state is lost either way
    int check_peer_timeout() {
        if (!peers_) return FG_OK;
        const int t = dev.peer_timeout();
        if (t == 1) return fail(FG_EPEER, "timed out waiting for a z-neighbour's halo (ranks out of step?)");
        if (t == 2) return fail(FG_EPEER, "timed out waiting for another rank's marker / wrench exchange (ranks out of step?)");
        return FG_OK;
    }

    // device time of the latest fg_step call, once its events have completed (wait: block until they have)
    void finish_timing(bool wait) {
        if (!timing_pending_) return;
        bool ok = false;
        const double ms = dev.toc_elapsed(wait, ok);
        if (!ok) return;
        timing_pending_ = false;
        last_ms_ = ms;
        last_mlups_ = ms > 0 ? double(L_.plane) * L_.nz * timed_substeps_ / ms / 1e3 : 0;
    }

    int get_stats(FgStats *o) {
        finish_timing(false);
        std::memset(o, 0, sizeof(*o));
        o->steps = steps_;
        o->cells = int64_t(L_.plane) * L_.nz;
        o->last_step_ms = last_ms_;
        o->last_mlups = last_mlups_;
        o->kernel_launches = dev.launches;
        o->n_markers = ib_.n_markers(); o->n_links = ib_.n_links(); o->band_cells = ib_.band_cells();
        o->parity = parity_;
        o->collide_ms = collide_ms_; o->collide_launches = last_collide_launches_; o->ib_ms = ib_ms_;
        o->collide_cells = last_collide_cells_;
        o->split_substeps = split_substeps_;
        o->pair_substeps = pair_substeps_;
        o->graph_launches = dev.graph_launches;
        return FG_OK;
    }

    // ------------------------------------------------------------------ halos
    bool internal_lo() const { return cfg.n_ranks > 1 && (cfg.rank > 0 || cfg.bc[FG_ZLO] == FG_BC_PERIODIC); }
    bool internal_hi() const { return cfg.n_ranks > 1 && (cfg.rank < cfg.n_ranks - 1 || cfg.bc[FG_ZHI] == FG_BC_PERIODIC); }
    bool has_lo_peer() const { return peers_ && internal_lo(); }
    bool has_hi_peer() const { return peers_ && internal_hi(); }
    int64_t halo_bytes() const { return int64_t(5) * L_.plane * sizeof(pop_t); }

    // the last step's parity decides which plane / slots carry the populations that crossed the face
    int halo_pack(int face, void *buf) {
        if (face != FG_ZLO && face != FG_ZHI) return fail(FG_EINVAL, "face must be FG_ZLO or FG_ZHI");
        if (steps_ == 0) return fail(FG_ESTATE, "fg_halo_pack: call after fg_step");
        const bool hi = face == FG_ZHI;
        const int done = parity_ ^ 1;
        pop_t *out = static_cast<pop_t *>(buf);
        for (int q = 0; q < 5; ++q) {
            const int slot = done == 0 ? (hi ? ZMT[q] : ZPT[q]) : (hi ? ZPT[q] : ZMT[q]);
            const int zz = done == 0 ? (hi ? L_.nz : 1) : (hi ? L_.nz + 1 : 0);
            if (!dev.d2h(out + size_t(q) * L_.plane, L_.f + slot * L_.slot + (long long)zz * L_.plane, L_.plane * sizeof(pop_t)))
                return cuda_fail();
        }
        return FG_OK;
    }

    int halo_unpack(int face, const void *buf) {
        if (face != FG_ZLO && face != FG_ZHI) return fail(FG_EINVAL, "face must be FG_ZLO or FG_ZHI");
        if (pending_faces_ <= 0) return fail(FG_ESTATE, "fg_halo_unpack: call after fg_step");
        const bool hi = face == FG_ZHI;
        pop_t *&st = stage_[hi];
        if (!st) st = static_cast<pop_t *>(dev.alloc(size_t(5) * L_.plane * sizeof(pop_t), err));
        if (!st) return FG_ENOMEM;
        if (!dev.h2d(st, buf, size_t(5) * L_.plane * sizeof(pop_t))) return cuda_fail();
        const int done = parity_ ^ 1;
        HaloParams p{};
        p.L = L_; p.C = C_; p.parity_done = done;
        p.op[1].mode = BC_WALL;
        FaceOp &op = p.op[0];
        op.mode = BC_PEER;
        op.hi = hi ? 0 : 1;                 // what arrives at my low face left the sender through ITS high face
        op.src = st; op.src_slot = L_.plane; op.src_off = 0; op.src_by_index = 1;
        op.dst = L_.f;
        // even: into my ghost plane; odd: into my boundary plane
        const int dzz = done == 0 ? (hi ? L_.nz + 1 : 0) : (hi ? L_.nz : 1);
        op.dst_off = (long long)dzz * L_.plane;
        // the sender's boundary plane is my ghost plane
        op.sender_solid = L_.solid ? L_.solid + (long long)(hi ? L_.nz + 1 : 0) * L_.plane : nullptr;
        op.receiver_solid = L_.solid ? L_.solid + (long long)(hi ? L_.nz : 1) * L_.plane : nullptr;
        Dim3 g{(L_.plane + ZFaceOp::kThreads - 1) / ZFaceOp::kThreads, 5, 1};
        if (!dev.template launch<ZFaceOp>(g, p) || !dev.sync()) return cuda_fail();
        --pending_faces_;
        return FG_OK;
    }

    int peer_export(FgPeerHandle *out) {
        PeerBlob b{};
        b.magic = kPeerMagic;
        b.nx = L_.nx; b.ny = L_.ny; b.nz = L_.nz;
        b.rank = cfg.rank; b.n_ranks = cfg.n_ranks;
        b.x_bytes = ib_.ready() ? ib_.xbuf_bytes() : 0;
        if (!dev.export_peer(L_.f, flags_, ib_.ready() ? ib_.xbuf() : nullptr, b, err)) return FG_EPEER;
        std::memset(out, 0, sizeof(*out));
        std::memcpy(out->bytes, &b, sizeof(b));
        return FG_OK;
    }

    int peer_connect(const FgPeerHandle *lo, const FgPeerHandle *hi) {
        if (cfg.n_ranks < 2) return fail(FG_ESTATE, "fg_peer_connect needs n_ranks > 1");
        if (internal_lo() != (lo != nullptr) || internal_hi() != (hi != nullptr))
            return fail(FG_EINVAL, "pass a handle for exactly the internal faces of this slab");
        const FgPeerHandle *hs[2] = {lo, hi};
        for (int s = 0; s < 2; ++s) {
            peer_f_[s] = nullptr; peer_flags_[s] = nullptr;
            if (!hs[s]) continue;
            PeerBlob b;
            std::memcpy(&b, hs[s]->bytes, sizeof(b));
            if (b.magic != kPeerMagic) return fail(FG_EPEER, "peer handle: bad magic");
            if (b.nx != L_.nx || b.ny != L_.ny || b.nz != L_.nz) return fail(FG_EPEER, "peer handle: slab geometry differs");
            void *x = nullptr;
            if (!dev.open_peer(b, &peer_f_[s], &peer_flags_[s], &x, err)) return FG_EPEER;
        }
        peers_ = true;
        pending_faces_ = 0;
        dev.graph_clear();
        return FG_OK;
    }

    // handles of ALL ranks (indexed by rank): halos as in peer_connect, plus the IB exchange between slabs
    int peer_connect_all(const FgPeerHandle *hs, int n) {
        if (cfg.n_ranks < 2) return fail(FG_ESTATE, "fg_peer_connect_all needs n_ranks > 1");
        if (n != cfg.n_ranks) return fail(FG_EINVAL, "pass one handle per rank");
        if (n > kMaxRanks) return fail(FG_ENOTSUP, "at most 8 ranks");
        // the marker message on the device was packed WITHOUT the culled layout of the exchange (no global-id column):
        // reading it with the other layout would silently mis-assign link origins and drop markers
        if (ib_.ready() && !ib_.exchange_on() && ib_.n_local() > 0)
            return fail(FG_ESTATE, "fg_peer_connect_all: call before fg_set_markers / fg_add_fish (the marker set must be sent again after connecting)");
        const int lo = (cfg.rank - 1 + n) % n, hi = (cfg.rank + 1) % n;
        if (int rc = peer_connect(internal_lo() ? hs + lo : nullptr, internal_hi() ? hs + hi : nullptr)) return rc;
        if (!ib_.ready()) return FG_OK;
        if (L_.nz < 4) return fail(FG_EINVAL, "bodies across slabs need at least 4 planes per slab");
        void *all[kMaxRanks] = {};
        for (int r = 0; r < n; ++r) {
            PeerBlob b;
            std::memcpy(&b, hs[r].bytes, sizeof(b));
            if (b.magic != kPeerMagic || b.rank != r || b.n_ranks != n) return fail(FG_EPEER, "peer handle: wrong rank order");
            if (b.x_bytes != ib_.xbuf_bytes()) return fail(FG_EPEER, "peer handle: IB capacity (max_markers / max_links) differs between ranks");
            if (r == cfg.rank) { all[r] = ib_.xbuf(); continue; }
            pop_t *f = nullptr; int *fl = nullptr;
            if (!dev.open_peer(b, &f, &fl, &all[r], err)) return FG_EPEER;
            if (!all[r]) return fail(FG_EPEER, "peer handle: rank has no IB exchange buffer");
        }
        ib_.enable_exchange(cfg.rank, n, internal_lo() ? all[lo] : nullptr, internal_hi() ? all[hi] : nullptr, all);
        dev.graph_clear();
        return FG_OK;
    }

    // everything that selects kernels or changes their arguments from one substep to the next
    GraphKey substep_key() const {
        GraphKey k{};
        k[0] = uint64_t(parity_) | (fish_.empty() ? 0u : 2u) | (halo_branch_now_ ? 4u : 0u);      // (substeps with a plane split are not captured)
        if (ib_.ready()) {
            uint64_t w[3];
            ib_.graph_key(w);
            k[1] = w[0]; k[2] = w[1]; k[3] = w[2];
        }
        return k;
    }
    int fail(int code, const std::string &m) { err = m; return code; }
    int cuda_fail() { err = dev.err; return FG_ECUDA; }

private:
    // NVTX range (CUDA policy; a no-op in the host emulation) around one phase of a substep
    struct Range {
        Dev &d;
        Range(Dev &dev_, const char *name) : d(dev_) { d.range_push(name); }
        ~Range() { d.range_pop(); }
    };
    Dim3 grid_planes(int planes) const {
        return Dim3{(L_.nx + 127) / 128, L_.ny, planes};
    }

    void setup_collision() {
        C_ = Collision{};
        C_.omega = float(1.0 / cfg.tau);
        bool all_zero = true;
        for (int k = 0; k < Q; ++k) all_zero = all_zero && cfg.mrt_rates[k] == 0.0;
        const double sn = 1.0 / cfg.tau;
        const double dflt[Q] = {0, 1.19, 1.4, 0, 1.2, 0, 1.2, 0, 1.2, sn, 1.4, sn, 1.4, sn, sn, sn, 1.98, 1.98, 1.98};
        for (int k = 0; k < Q; ++k) C_.rate[k] = float(all_zero ? dflt[k] : cfg.mrt_rates[k]);
        C_.rate[0] = C_.rate[3] = C_.rate[5] = C_.rate[7] = 0.f;
        for (int d = 0; d < 3; ++d) C_.g[d] = float(cfg.body_force[d]);
        for (int f = 0; f < 6; ++f)
            for (int i = 0; i < Q; ++i)
                C_.wallterm[f][i] = float(6.0 * WD[i] * (CXT[i] * cfg.wall_u[f][0] + CYT[i] * cfg.wall_u[f][1] + CZT[i] * cfg.wall_u[f][2]));
        const double r = cfg.inlet_rho, ux = cfg.inlet_u[0], uy = cfg.inlet_u[1], uz = cfg.inlet_u[2];
        const double uu = ux * ux + uy * uy + uz * uz;
        for (int i = 0; i < Q; ++i) {
            const double cu = CXT[i] * ux + CYT[i] * uy + CZT[i] * uz;
            C_.heq_in[i] = float(WD[i] * ((r - 1.0) + r * (3.0 * cu + 4.5 * cu * cu - 1.5 * uu)));
        }
    }

    template <int PARITY, int MODE>
    bool launch_collide_pm(const StepParams &p, Dim3 g) {
        // 8 CTAs/SM (64 registers) on SMALL z-slabs, 9 (56 registers) otherwise (lbm_core.cuh StreamCollide): the ~4 us per
        // step that 9 cost the high-priority chain on slabs of 4.2 M cells outweigh its 2 % faster kernel only below ~8 M
        // cells per rank (512^3 and the school on 8 GPUs were measured with 9: 318 252 / 272 528 MLUPS)
        if (small_peered_slab())
            return cfg.collision == FG_MRT ? dev.template launch<StreamCollide<PARITY, true, MODE, 8>>(g, p)
                                           : dev.template launch<StreamCollide<PARITY, false, MODE, 8>>(g, p);
        return cfg.collision == FG_MRT ? dev.template launch<StreamCollide<PARITY, true, MODE>>(g, p)
                                       : dev.template launch<StreamCollide<PARITY, false, MODE>>(g, p);
    }
    // opt-in experiment: even step with 4 / 2 cells per thread and 16- / 8-byte accesses (lbm_core.cuh StreamCollideEvenVec)
    int even_vec_width() const {
        if (cfg.flags & FG_FLAG_EVEN_SCALAR) return 0;
        if ((cfg.flags & FG_FLAG_EVEN_VEC4) && L_.nx % 4 == 0) return 4;
        if ((cfg.flags & FG_FLAG_EVEN_VEC2) && L_.nx % 2 == 0) return 2;
        // default (round 2, gpu pass 2): two cells per thread with 64-bit accesses wherever a row fills whole CTAs that way —
        // +2.6 ... +3.2 % MLUPS on 256- and 512-wide lattices (fewer memory instructions per byte; the 4-cell form gains
        // less: 125 registers leave 4 CTAs per SM).  Narrower rows would leave half of each CTA idle and keep the scalar kernel.
        if (small_peered_slab()) return 0;
#if defined(FG_POP16)
        // the 16-bit build is issue-bound: four cells per thread (64-bit accesses) where rows fill whole CTAs that way
        // (512^3, 100 steps, gpu pass b14: 61 922 against 58 617 MLUPS with two cells)
        if (kVecDefault != 0 && L_.nx % (4 * kCollideThreads) == 0) return 4;
#endif
        if (L_.nx % (2 * kCollideThreads) == 0 || narrow_rows_log2() > 0) return kVecDefault;
        return 0;
    }
    // Peered z-slabs below 8 M cells keep the one-cell kernels (at 8 CTAs/SM, launch_collide_pm) in substeps WITHOUT the halo
    // branch (no bodies, or a body at a slab face): there the high-priority chain (neighbour flags -> IB kernels -> boundary
    // planes -> halo push -> near planes) is as long as the interior collide beside it, and each of its ~12 dependent launches
    // waits for a resident interior CTA to retire — 3.5 us for a one-cell CTA, 5.3 us for a two-cell one.  Two GPUs,
    // 256x128x128 channel + sphere per GPU, 2000 steps, serial chain (gpu pass b5): one-cell kernels 71 636 MLUPS, two-cell even /
    // odd / both 69 062 / 69 993 / 67 574, one-cell at 9 CTAs/SM 68 957 — although the two-cell kernel itself is 5 % faster per
    // launch (54.4 against 57.4 us).  With the halo branch the chain is four launches shorter and the order flips (pass b10:
    // 75 200 one-cell, 77 400 two-cell).  On 512^3 slabs the chain is 1 % of the step.
    bool small_slab() const { return (long long)L_.plane * L_.nz < (8ll << 20); }
    bool small_peered_slab() const { return peers_ && slab_occ8_ && small_slab() && !halo_branch_now_; }
    // rows of 128 or 64 cells: a CTA of the two-cell kernels (256 cells) takes 2 or 4 consecutive rows of the launch, so the
    // 256x128x128 channel of BASELINE.json configs[1] runs the two-cell kernels as well; log2(rows per CTA), 0 otherwise
    int narrow_rows_log2() const {
        if (kCollideThreads != 128) return 0;
        return L_.nx == 128 ? 1 : (L_.nx == 64 ? 2 : 0);
    }
    // launch geometry of a two-cell kernel over `rows` rows and `planes` planes (and the row mapping it needs in StepParams)
    Dim3 vec2_grid(StepParams &p, int rows, int planes) const {
        const int rl = narrow_rows_log2();
        p.rows = rows;
        if (rl > 0) {
            p.rows_per_cta_log2 = rl; p.nx_log2 = L_.nx == 128 ? 7 : 6;
            return Dim3{1, (rows + (1 << rl) - 1) >> rl, planes};
        }
        return Dim3{(L_.nx / 2 + kCollideThreads - 1) / kCollideThreads, rows, planes};
    }
    // what the library picks by itself where rows are wide enough: two cells per thread.  The opt-in 16-bit-storage build
    // gains most from it (it is issue-bound: 512^3, 100 steps, gpu pass b1: 50 919 MLUPS with one cell per thread in both
    // steps, 55 872 / 53 565 with two cells in the even / odd step only, 59 473 with both = +17 %)
#if defined(FG_POP16_VEC_DEFAULT)
    static constexpr int kVecDefault = FG_POP16_VEC_DEFAULT;
#else
    static constexpr int kVecDefault = 2;
#endif
    // bulk odd step (rows without a blocked link, periodic x) with two cells per thread: the default wherever a row fills
    // whole CTAs that way, like the even step (round 2, gpu pass 7: +2.6 % MLUPS at 20 steps, +3.9 % sustained over 300 steps
    // on 512^3 — 19 % fewer instructions, 6 CTAs of 80 registers); FG_FLAG_ODD_VEC2 forces it for any even nx, FG_FLAG_ODD_SCALAR
    // the one-cell kernel
    bool odd_vec2() const {
        if ((cfg.flags & FG_FLAG_ODD_SCALAR) || L_.solid || L_.nx % 2 != 0 || L_.nx < 4) return false;
        if (cfg.flags & FG_FLAG_ODD_VEC2) return true;
        return kVecDefault != 0 && !small_peered_slab() && (L_.nx % (2 * kCollideThreads) == 0 || narrow_rows_log2() > 0);
    }
    template <bool XW, bool NARROW>
    bool launch_odd_vec2_t(const StepParams &p, Dim3 g) {
        return cfg.collision == FG_MRT ? dev.template launch<StreamCollideOddVec2<true, XW, NARROW>>(g, p)
                                       : dev.template launch<StreamCollideOddVec2<false, XW, NARROW>>(g, p);
    }
    bool launch_odd_vec2(const StepParams &p, Dim3 g, bool xwall = false) {
        if (p.rows_per_cta_log2 > 0) return xwall ? launch_odd_vec2_t<true, true>(p, g) : launch_odd_vec2_t<false, true>(p, g);
        return xwall ? launch_odd_vec2_t<true, false>(p, g) : launch_odd_vec2_t<false, false>(p, g);
    }
    bool launch_even_vec(const StepParams &p, Dim3 g, int vec) {
        const bool mrt = cfg.collision == FG_MRT;
        if (vec == 4) return mrt ? dev.template launch<StreamCollideEvenVec<true, 4>>(g, p) : dev.template launch<StreamCollideEvenVec<false, 4>>(g, p);
        if (p.rows_per_cta_log2 > 0)
            return mrt ? dev.template launch<StreamCollideEvenVec<true, 2, true>>(g, p) : dev.template launch<StreamCollideEvenVec<false, 2, true>>(g, p);
        return mrt ? dev.template launch<StreamCollideEvenVec<true, 2>>(g, p) : dev.template launch<StreamCollideEvenVec<false, 2>>(g, p);
    }
    // hole: planes [hole_b, hole_e) inside [zb, ze) are left out (zstride 1 only)
    bool launch_rows(int mode, int zb, int ze, int y0, int ystride, int rows, const ForceField &F, int zstride = 1, int hole_b = 0, int hole_e = 0,
                     bool timed = false) {
        const int hole = hole_e > hole_b ? hole_e - hole_b : 0;
        const int planes = (ze - zb - hole + zstride - 1) / zstride;
        if (planes <= 0 || rows <= 0) return true;
        const bool down = parity_ == 1 && !(cfg.flags & FG_FLAG_NO_SWEEP_FLIP);
        StepParams p{L_, C_, F, zb, zstride, hole ? hole_b : 0x7fffffff, hole, down ? planes - 1 : -1, y0, ystride, 0, 0, rows, {}};
        for (int s = 0; s < Q; ++s)
            for (int d = 0; d < 3; ++d) p.kz[s][d] = (long long)kPopBytes * (s * L_.slot + (long long)(d - 1) * L_.plane);
        const Dim3 g{(L_.nx + kCollideThreads - 1) / kCollideThreads, rows, planes};
        const bool prof = timed && (cfg.flags & FG_FLAG_PROFILE);
        if (prof) {   // FG_FLAG_PROFILE: event pair around the bulk launch alone, and the cells it updates
            dev.mark(0);
            ++collide_launches_;
            collide_cells_ += int64_t(L_.nx) * rows * planes;
        }
        bool ok;
        if (parity_ == 0) {
            // the even step is purely local: only obstacles need the checked variant
            const int vec = L_.solid ? 0 : even_vec_width();
            if (vec == 2) { const Dim3 g2 = vec2_grid(p, rows, planes); ok = launch_even_vec(p, g2, 2); }
            else if (vec) ok = launch_even_vec(p, Dim3{(L_.nx / vec + kCollideThreads - 1) / kCollideThreads, rows, planes}, vec);
            else ok = mode == CHECK_ALL && L_.solid ? launch_collide_pm<0, CHECK_ALL>(p, g) : launch_collide_pm<0, CHECK_NONE>(p, g);
        } else {
            switch (mode) {
                case CHECK_ALL: ok = launch_collide_pm<1, CHECK_ALL>(p, g); break;
                case CHECK_XEDGE: ok = launch_collide_pm<1, CHECK_XEDGE>(p, g); break;
                case CHECK_XWARP:
                    // rows between x walls (the tank, the school): the two-cell kernel with the wall selects in its row-end warps
                    if (odd_vec2()) { const Dim3 g2 = vec2_grid(p, rows, planes); ok = launch_odd_vec2(p, g2, true); }
                    else ok = launch_collide_pm<1, CHECK_XWARP>(p, g);
                    break;
                default:
                    if (odd_vec2()) { const Dim3 g2 = vec2_grid(p, rows, planes); ok = launch_odd_vec2(p, g2); }
                    else ok = launch_collide_pm<1, CHECK_NONE>(p, g);
                    break;
            }
        }
        if (prof) dev.mark(0);
        return ok;
    }
    // Partition planes [zz_begin, zz_end) (minus the hole) so that boundary code only runs where a link can be blocked.
    // zstride > 1 covers planes zz_begin, zz_begin + zstride, ... (the two boundary planes of a slab in one launch)
    bool launch_collide(int zz_begin, int zz_end, const ForceField &F, int zstride = 1, int hole_b = 0, int hole_e = 0) {
        Range r(dev, "fg:stream_collide");
        if (hole_e <= hole_b || hole_e <= zz_begin || hole_b >= zz_end) hole_b = hole_e = 0;
        if (hole_e > hole_b) {           // a hole touching an end of the range just shortens the range
            if (hole_b <= zz_begin) { zz_begin = hole_e; hole_b = hole_e = 0; }
            else if (hole_e >= zz_end) { zz_end = hole_b; hole_b = hole_e = 0; }
        }
        if (zz_end <= zz_begin) return true;
        const int ny = L_.ny;
        if (L_.solid || parity_ == 0) return launch_rows(CHECK_ALL, zz_begin, zz_end, 0, 1, ny, F, zstride, hole_b, hole_e, true);
        const bool zlo_wall = L_.bc_zlo == BC_WALL && L_.z0 == 0, zhi_wall = L_.bc_zhi == BC_WALL && L_.z0 + L_.nz == L_.nzg;
        if (hole_e > hole_b && (zlo_wall || zhi_wall))   // wall planes are peeled off the ends below: keep that logic hole-free
            return launch_collide(zz_begin, hole_b, F) && launch_collide(hole_e, zz_end, F);
        int zb = zz_begin, ze = zz_end;
        bool ok = true;
        // The launches below touch disjoint cells, so the thin checked ones run on a forked stream (parallel
        // graph branches) while the bulk runs on the current one.
        const int base = dev.current();
        const bool thin = (zlo_wall && zb <= 1 && 1 < ze) || (zhi_wall && zb <= L_.nz && L_.nz < ze) || (L_.wall_y && ny >= 2);
        if (thin) ok = dev.fork_to(base + 1);
        if (zlo_wall && zb <= 1 && 1 < ze) { ok = ok && launch_rows(CHECK_ALL, 1, 2, 0, 1, ny, F); zb = 2; }
        if (zhi_wall && zb <= L_.nz && L_.nz < ze) { ok = ok && launch_rows(CHECK_ALL, L_.nz, L_.nz + 1, 0, 1, ny, F); ze = L_.nz; }
        const int bulk = L_.wall_x ? ((cfg.flags & FG_FLAG_NO_XWARP) ? CHECK_XEDGE : CHECK_XWARP) : CHECK_NONE;
        if (L_.wall_y && ny >= 2) {
            ok = ok && launch_rows(CHECK_ALL, zb, ze, 0, ny - 1, 2, F, zstride, hole_b, hole_e);
            if (thin) ok = ok && dev.switch_to(base);
            ok = ok && launch_rows(bulk, zb, ze, 1, 1, ny - 2, F, zstride, hole_b, hole_e, true);
        } else {
            if (thin) ok = ok && dev.switch_to(base);
            ok = ok && launch_rows(L_.wall_y ? CHECK_ALL : bulk, zb, ze, 0, 1, ny, F, zstride, hole_b, hole_e, true);
        }
        if (thin) ok = dev.join_from(base + 1) && ok;
        return ok;
    }

    // the two boundary planes of a slab in one launch (plane stride nz - 1), unless one of them is a z-wall plane
    bool launch_boundary_planes(const ForceField &F) {
        const bool zwall = (L_.bc_zlo == BC_WALL && L_.z0 == 0) || (L_.bc_zhi == BC_WALL && L_.z0 + L_.nz == L_.nzg);
        if (zwall) return launch_collide(1, 2, F) && launch_collide(L_.nz, L_.nz + 1, F);
        return launch_collide(1, L_.nz + 1, F, L_.nz - 1);
    }

    // planes [zb, ze) that are in neither of the (ascending, disjoint) ranges done[0], done[1]
    bool launch_collide_except(int zb, int ze, const ForceField &F, const int done[2][2]) {
        int at = zb;
        for (int i = 0; i < 2; ++i) {
            if (done[i][1] <= done[i][0]) continue;
            if (!launch_collide(at, std::min(done[i][0], ze), F)) return false;
            at = std::max(at, done[i][1]);
        }
        return launch_collide(at, ze, F);
    }
    // EVEN step on planes [zb, ze) minus the hole and, `lag` planes behind it in the same launch, the following ODD step
    // on the planes whose z-neighbours belong to the launch (lbm_core.cuh StreamCollidePair).  Returns the planes that
    // took both steps in done[2][2] = up to two ranges [b, e).
    bool launch_pair(int zb, int ze, int hole_b, int hole_e, int done[2][2]) {
        if (hole_e <= hole_b || hole_e <= zb || hole_b >= ze) hole_b = hole_e = 0;
        if (hole_e > hole_b) {
            if (hole_b <= zb) { zb = hole_e; hole_b = hole_e = 0; }
            else if (hole_e >= ze) { ze = hole_b; hole_b = hole_e = 0; }
        }
        const int hole = hole_e - hole_b, planes = ze - zb - hole;
        done[0][0] = done[0][1] = done[1][0] = done[1][1] = 0;
        if (planes <= 0) return true;
        PairParams p{};
        p.s = StepParams{L_, C_, ForceField{}, zb, 1, hole ? hole_b : 0x7fffffff, hole, -1, 0, 1, 0, 0, 0, {}};
        for (int s = 0; s < Q; ++s)
            for (int d = 0; d < 3; ++d) p.s.kz[s][d] = (long long)kPopBytes * (s * L_.slot + (long long)(d - 1) * L_.plane);
        p.planes = planes; p.rows = L_.ny; p.xblocks = (L_.nx + 127) / 128;
        p.lag = std::min(pair_lag(), planes);
        // a zero-gradient outlet copies from the boundary plane after the even step what the odd step of the plane
        // next to it would already have overwritten: that plane waits, too
        const bool out_lo = L_.bc_zlo == BC_OUTLET && L_.z0 == 0 && zb == 1, out_hi = L_.bc_zhi == BC_OUTLET && L_.z0 + L_.nz == L_.nzg && ze == L_.nz + 1;
        p.odd_lo = out_lo ? 2 : 1; p.odd_hi = planes - (out_hi ? 2 : 1);
        p.odd_skip_v = hole ? hole_b - zb : -1;
        p.odd_y_lo = L_.wall_y ? 1 : 0; p.odd_y_hi = L_.wall_y ? L_.ny - 1 : L_.ny;
        p.ticket = pair_ctr_; p.done = pair_ctr_ + 1;
        if (hole) { done[0][0] = zb + p.odd_lo; done[0][1] = hole_b - 1; done[1][0] = hole_e + 1; done[1][1] = ze - (out_hi ? 2 : 1); }
        else { done[0][0] = zb + p.odd_lo; done[0][1] = ze - (out_hi ? 2 : 1); }
        for (int i = 0; i < 2; ++i) if (done[i][1] < done[i][0]) done[i][0] = done[i][1] = 0;
        const Dim3 g{p.xblocks, p.rows, planes};
        const bool mrt = cfg.collision == FG_MRT;
        if (!dev.zero_on_current(pair_ctr_, sizeof(int) * size_t(planes + 1))) return false;
        if (L_.wall_x) return mrt ? dev.template launch_ticketed<StreamCollidePair<true, CHECK_XEDGE>>(g, p)
                                  : dev.template launch_ticketed<StreamCollidePair<false, CHECK_XEDGE>>(g, p);
        return mrt ? dev.template launch_ticketed<StreamCollidePair<true, CHECK_NONE>>(g, p)
                   : dev.template launch_ticketed<StreamCollidePair<false, CHECK_NONE>>(g, p);
    }
    // the odd step runs this many planes behind the even step: far enough that it never waits, near enough that the
    // planes in between (lag x 76 B x nx x ny) stay in L2
    int pair_lag() const {
        if (cfg.pair_lag > 0) return cfg.pair_lag;
        const long long plane_bytes = 76ll * L_.plane;
        return int(std::max<long long>(2, std::min<long long>(8, (24ll << 20) / plane_bytes)));
    }

    // z-face plane ops after the step of parity `parity_`: one launch covers both faces
    bool launch_faces() {
        Range r(dev, "fg:z_faces");
        HaloParams p{};
        p.L = L_; p.C = C_; p.parity_done = parity_;
        bool any = false;
        for (int s = 0; s < 2; ++s) {
            const bool hi = s == 1;
            FaceOp &op = p.op[s];
            op.mode = BC_WALL;
            op.hi = hi;
            const bool at_global = hi ? (cfg.rank == cfg.n_ranks - 1) : (cfg.rank == 0);
            const int bc = hi ? cfg.bc[FG_ZHI] : cfg.bc[FG_ZLO];
            const bool internal = hi ? internal_hi() : internal_lo();
            pop_t *dst = nullptr;
            if (cfg.n_ranks == 1) {
                if (bc == FG_BC_PERIODIC) dst = L_.f;            // wrap onto myself
                else if (bc == FG_BC_INLET || bc == FG_BC_OUTLET) op.mode = bc;
            } else if (internal) {
                if (peers_) dst = peer_f_[s];                    // neighbour's lattice over NVLink
            } else if (at_global && (bc == FG_BC_INLET || bc == FG_BC_OUTLET)) {
                op.mode = bc;
            }
            if (dst) {
                op.mode = BC_PEER;
                op.src = L_.f; op.src_slot = L_.slot; op.src_by_index = 0;
                op.dst = dst;
                if (parity_ == 0) {   // boundary plane -> neighbour ghost plane
                    op.src_off = (long long)(hi ? L_.nz : 1) * L_.plane;
                    op.dst_off = (long long)(hi ? 0 : L_.nz + 1) * L_.plane;
                } else {              // my ghost plane -> neighbour boundary plane
                    op.src_off = (long long)(hi ? L_.nz + 1 : 0) * L_.plane;
                    op.dst_off = (long long)(hi ? 1 : L_.nz) * L_.plane;
                }
                op.sender_solid = L_.solid ? L_.solid + (long long)(hi ? L_.nz : 1) * L_.plane : nullptr;
                // my ghost-plane flags mirror the neighbour's boundary plane (fg_set_solid fills them from the global field)
                op.receiver_solid = L_.solid ? L_.solid + (long long)(hi ? L_.nz + 1 : 0) * L_.plane : nullptr;
            }
            any = any || op.mode != BC_WALL;
        }
        if (any) {
            // plain copies of whole planes (periodic wrap, halo push; no x / y walls, no obstacles): four cells per thread
            bool copies = !L_.wall_x && !L_.wall_y && !L_.solid && L_.plane % 4 == 0 && L_.slot % 4 == 0 && face_copy4_;
            for (int s = 0; s < 2; ++s) copies = copies && (p.op[s].mode == BC_WALL || (p.op[s].mode == BC_PEER && !p.op[s].src_by_index));
            if (copies) {
                Dim3 g{int((L_.plane / 4 + ZFaceCopy4::kThreads - 1) / ZFaceCopy4::kThreads), 5, 2};
                if (!dev.template launch<ZFaceCopy4>(g, p)) return false;
            } else {
                Dim3 g{(L_.plane + ZFaceOp::kThreads - 1) / ZFaceOp::kThreads, 5, 2};
                if (!dev.template launch<ZFaceOp>(g, p)) return false;
            }
        }
        if (peers_) {
            // publish "my halos of tick+1 are in your memory" to both neighbours
            if (!dev.signal_flags(flags_, has_lo_peer() ? peer_flags_[0] + 1 : nullptr, has_hi_peer() ? peer_flags_[1] + 0 : nullptr))
                return false;
        }
        return true;
    }

    // host part: marker coordinates of all bodies (and the planes their stencils touch)
    void emit_bodies() {
        int n = 0, nl = 0;
        for (auto &f : fish_) { n += f.n_markers(); nl += f.n_links(); }
        mX_.resize(3 * size_t(n)); mU_.resize(3 * size_t(n)); mdV_.resize(n); mlink_.resize(n);
        morigin_.resize(3 * size_t(nl));
        int mo = 0, lo = 0;
        for (auto &f : fish_) {
            f.emit_markers(&mX_[3 * size_t(mo)], &mU_[3 * size_t(mo)], &mdV_[mo], &mlink_[mo], lo, &morigin_[3 * size_t(lo)]);
            mo += f.n_markers(); lo += f.n_links();
        }
        ib_.update_range(n, mX_.data());
    }
    // device part: the pinned marker message goes up (inside the substep's graph)
    int upload_bodies() {
        return ib_.set_markers(dev, int(mdV_.size()), mX_.data(), mU_.data(), mdV_.data(), mlink_.data(), morigin_.data(),
                               int(morigin_.size() / 3), err, /*range_known=*/true);
    }
    int bodies_to_markers() {
        emit_bodies();
        return upload_bodies();
    }

    Lattice L_{};
    Collision C_{};
    int nzl_ = 0;
    int parity_ = 0;
    int64_t steps_ = 0;
    double last_ms_ = 0, last_mlups_ = 0, collide_ms_ = 0, ib_ms_ = 0;
    bool timing_pending_ = false;
    int timed_substeps_ = 0;
    int64_t collide_launches_ = 0, last_collide_launches_ = 0, collide_cells_ = 0, last_collide_cells_ = 0;
    int64_t split_substeps_ = 0, pair_substeps_ = 0;
    bool slab_occ8_ = std::getenv("FG_SLAB_OCC9") == nullptr;     // A/B switch for launch_collide_pm's small-slab rule
    // boundary planes + push beside the IB kernels: 1 forced on (FG_HALO_BRANCH), 0 off (FG_NO_HALO_BRANCH), -1 the default (small slabs)
    int halo_branch_ = std::getenv("FG_NO_HALO_BRANCH") ? 0 : (std::getenv("FG_HALO_BRANCH") ? 1 : -1);
    bool face_copy4_ = std::getenv("FG_NO_FACE_COPY4") == nullptr;      // A/B switch: ZFaceCopy4 for plain plane copies
    bool halo_first_ = std::getenv("FG_HALO_FIRST") != nullptr;        // A/B: submit the branch before the IB kernels
    bool halo_branch_now_ = false;
    int *pair_ctr_ = nullptr;      // [1 + nz + 2] ticket + per-plane completion counters of StreamCollidePair
    bool split_ = false;           // this substep: far planes collide beside the IB kernels
    int near_a_ = 1, near_b_ = 1;  // planes [near_a_, near_b_) wait for the IB force
    uint8_t *solid_ = nullptr;
    int *flags_ = nullptr;         // [0] written by my z-low neighbour, [1] by my z-high neighbour, [2] my own halo
                                   // counter (device-resident, never reset: orders pushes between neighbours), [3] timeout
    pop_t *stage_[2] = {nullptr, nullptr};
    bool peers_ = false;
    pop_t *peer_f_[2] = {nullptr, nullptr};
    int *peer_flags_[2] = {nullptr, nullptr};
    int pending_faces_ = 0;
    float *probe_d_ = nullptr, *probe_h_ = nullptr;   // fg_probe: device [7 cap] / pinned host [7 cap], grow-only
    int probe_cap_ = 0;
    unsigned long long *finite_d_ = nullptr;           // fg_check_finite counter
    double *solid_force_d_ = nullptr;                  // fg_get_solid_force accumulator [6]
    IbState<Dev> ib_;
    std::vector<Fish> fish_;
    std::vector<float> action_;
    std::vector<float> mX_, mU_, mdV_;
    std::vector<int32_t> mlink_;
    std::vector<double> morigin_;
};

}  // namespace fg
