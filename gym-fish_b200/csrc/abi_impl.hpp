// abi_impl.hpp — the extern "C" entry points of include/fishgym.h over SimT<FG_DEV>.
// Included once by fishgym_cuda.cu (FG_DEV = CudaDev) and once by tests/emu (FG_DEV = HostDev).
// Argument meaning and error behaviour follow include/fishgym.h; no exception leaves this file.
#pragma once
#include "sim.hpp"

#include <new>

#ifndef FG_DEV
#error "define FG_DEV (device policy) and FG_BACKEND_NAME before including abi_impl.hpp"
#endif

struct FgSim {
    fg::SimT<FG_DEV> sim;
};

namespace {
thread_local std::string g_create_error;

int create_fail(int code, const std::string &m) {
    g_create_error = m;
    return code;
}

int validate(const FgConfig *cfg) {
    if (cfg->struct_size != int32_t(sizeof(FgConfig))) return create_fail(FG_EINVAL, "FgConfig.struct_size mismatch (ABI)");
    if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 1) return create_fail(FG_EINVAL, "lattice dimensions must be >= 1");
    if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks || cfg->nz % cfg->n_ranks)
        return create_fail(FG_EINVAL, "bad slab decomposition: need 0 <= rank < n_ranks and nz % n_ranks == 0");
    if (!(cfg->tau > 0.5)) return create_fail(FG_EINVAL, "tau must be > 0.5");
    if (cfg->collision != FG_BGK && cfg->collision != FG_MRT) return create_fail(FG_EINVAL, "unknown collision model");
    if (cfg->ib_iterations < 0 || cfg->ib_iterations > 16) return create_fail(FG_EINVAL, "ib_iterations must be in 0 .. 16");
    for (int f = 0; f < 6; ++f) {
        const int b = cfg->bc[f];
        if (b < FG_BC_PERIODIC || b > FG_BC_OUTLET) return create_fail(FG_EINVAL, "unknown boundary condition");
        if (f < FG_ZLO && b > FG_BC_WALL) return create_fail(FG_EINVAL, "inlet/outlet are supported on the z faces only");
    }
    for (int a = 0; a < 3; ++a)
        if ((cfg->bc[2 * a] == FG_BC_PERIODIC) != (cfg->bc[2 * a + 1] == FG_BC_PERIODIC))
            return create_fail(FG_EINVAL, "periodic boundaries must be set on both faces of an axis");
    if (cfg->max_markers < 0 || cfg->max_links < 0) return create_fail(FG_EINVAL, "max_markers / max_links must be >= 0");
    return FG_OK;
}
}  // namespace

// every entry point makes the handle's device current ONCE; the device policy does not repeat it per operation
#define FG_TRY try { s->sim.dev.enter();
#define FG_CATCH(s)                                                                   \
    }                                                                                 \
    catch (const std::bad_alloc &) { (s)->sim.err = "host out of memory"; return FG_ENOMEM; } \
    catch (const std::exception &e) { (s)->sim.err = e.what(); return FG_EINVAL; }

extern "C" {

int fg_abi_version(void) { return FG_ABI_VERSION; }
const char *fg_backend_name(void) { return FG_BACKEND_NAME; }
const char *fg_last_error(const FgSim *s) { return s ? s->sim.err.c_str() : g_create_error.c_str(); }

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
    if (!cfg || !out) return create_fail(FG_EINVAL, "null argument");
    *out = nullptr;
    if (int rc = validate(cfg)) return rc;
    FgSim *s = new (std::nothrow) FgSim;
    if (!s) return create_fail(FG_ENOMEM, "host out of memory");
    int rc;
    try {
        rc = s->sim.create(*cfg);
    } catch (const std::exception &e) {
        s->sim.err = e.what();
        rc = FG_ENOMEM;
    }
    if (rc != FG_OK) {
        g_create_error = s->sim.err;
        s->sim.destroy();
        delete s;
        return rc;
    }
    *out = s;
    return FG_OK;
}

int fg_destroy(FgSim *s) {
    if (!s) return FG_OK;
    s->sim.destroy();
    delete s;
    return FG_OK;
}

int fg_reset(FgSim *s, uint64_t seed) {
    if (!s) return FG_EINVAL;
    FG_TRY return s->sim.reset(seed); FG_CATCH(s)
}
int fg_set_fields(FgSim *s, const float *rho, const float *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    FG_TRY return s->sim.set_fields(rho, u); FG_CATCH(s)
}
int fg_get_fields(FgSim *s, float *rho, float *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    FG_TRY return s->sim.get_fields(rho, u); FG_CATCH(s)
}
int fg_get_fields_f64(FgSim *s, double *rho, double *u) {
    if (!s || !rho || !u) return FG_EINVAL;
    FG_TRY return s->sim.get_fields_f64(rho, u); FG_CATCH(s)
}
int fg_set_populations(FgSim *s, const float *f) {
    if (!s || !f) return FG_EINVAL;
    FG_TRY return s->sim.set_populations(f); FG_CATCH(s)
}
int fg_get_populations(FgSim *s, float *f) {
    if (!s || !f) return FG_EINVAL;
    FG_TRY return s->sim.get_populations(f); FG_CATCH(s)
}
int fg_set_solid(FgSim *s, const uint8_t *g) {
    if (!s) return FG_EINVAL;
    FG_TRY return s->sim.set_solid(g); FG_CATCH(s)
}

int fg_set_markers(FgSim *s, int32_t n, const float *X, const float *U, const float *dV, const int32_t *link) {
    if (!s || n < 0 || (n > 0 && (!X || !U || !dV))) return FG_EINVAL;
    FG_TRY return s->sim.set_markers(n, X, U, dV, link); FG_CATCH(s)
}
int fg_set_link_origins(FgSim *s, int32_t n, const double *o) {
    if (!s || n < 0 || (n && !o)) return FG_EINVAL;
    FG_TRY return s->sim.set_link_origins(n, o); FG_CATCH(s)
}
int fg_get_index_map(FgSim *s, int32_t *base3, int32_t *owner) {
    if (!s || !base3 || !owner) return FG_EINVAL;
    FG_TRY return s->sim.ib().get_index_map(s->sim.dev, base3, owner, s->sim.err); FG_CATCH(s)
}
int fg_get_marker_forces(FgSim *s, float *F3) {
    if (!s || !F3) return FG_EINVAL;
    FG_TRY return s->sim.ib().get_marker_array(s->sim.dev, true, F3, s->sim.err); FG_CATCH(s)
}
int fg_get_marker_velocities(FgSim *s, float *U3) {
    if (!s || !U3) return FG_EINVAL;
    FG_TRY return s->sim.ib().get_marker_array(s->sim.dev, false, U3, s->sim.err); FG_CATCH(s)
}
int fg_get_link_wrenches(FgSim *s, double *w6) {
    if (!s || !w6) return FG_EINVAL;
    FG_TRY
    if (!s->sim.ib().ready()) return FG_OK;
    if (!s->sim.ib().forces_valid()) {       // markers were (re)set and no step has run since: nothing to report yet
        std::fill(w6, w6 + 6 * size_t(s->sim.ib().n_links()), 0.0);
        return FG_OK;
    }
    if (int rc = s->sim.ib().fetch_wrenches(s->sim.dev, s->sim.err)) return rc;
    std::memcpy(w6, s->sim.ib().wrench_ptr(), sizeof(double) * 6 * s->sim.ib().n_links());
    return FG_OK;
    FG_CATCH(s)
}
int fg_get_force_field(FgSim *s, float *F) {
    if (!s || !F) return FG_EINVAL;
    FG_TRY return s->sim.get_force_field(F); FG_CATCH(s)
}

int fg_probe(FgSim *s, int32_t n, const float *X, float *out4) {
    if (!s || n < 0 || (n && (!X || !out4))) return FG_EINVAL;
    FG_TRY return s->sim.probe(n, X, out4); FG_CATCH(s)
}

int fg_add_fish(FgSim *s, const FgFishDesc *d, int32_t *id) {
    if (!s || !d) return FG_EINVAL;
    FG_TRY return s->sim.add_fish(*d, id); FG_CATCH(s)
}
int fg_set_action(FgSim *s, const float *a, int32_t n) {
    if (!s || (n && !a)) return FG_EINVAL;
    FG_TRY return s->sim.set_action(a, n); FG_CATCH(s)
}
int fg_get_obs(FgSim *s, float *o, int32_t n) {
    if (!s || !o) return FG_EINVAL;
    FG_TRY return s->sim.get_obs(o, n); FG_CATCH(s)
}
int fg_obs_size(FgSim *s) { return s ? s->sim.obs_size() : FG_EINVAL; }
int fg_action_size(FgSim *s) { return s ? s->sim.action_size() : FG_EINVAL; }
int fg_get_markers(FgSim *s, float *X, float *U, int32_t *link, int32_t cap) {
    if (!s) return FG_EINVAL;
    FG_TRY
    if (!s->sim.ib().ready()) return 0;
    return s->sim.ib().get_markers(s->sim.dev, X, U, link, cap, s->sim.err);
    FG_CATCH(s)
}

int fg_step(FgSim *s, int32_t n) {
    if (!s) return FG_EINVAL;
    FG_TRY return s->sim.step(n); FG_CATCH(s)
}
int fg_sync(FgSim *s) {
    if (!s) return FG_EINVAL;
    s->sim.dev.enter();
    if (!s->sim.dev.sync()) return s->sim.cuda_fail();
    return s->sim.check_peer_timeout();
}
int fg_get_stats(FgSim *s, FgStats *o) {
    if (!s || !o) return FG_EINVAL;
    FG_TRY return s->sim.get_stats(o); FG_CATCH(s)
}

int fg_check_finite(FgSim *s, int64_t *n_bad) {
    if (!s || !n_bad) return FG_EINVAL;
    FG_TRY return s->sim.check_finite(n_bad); FG_CATCH(s)
}

int fg_get_solid_force(FgSim *s, const double *origin3, double *out6) {
    if (!s || !out6) return FG_EINVAL;
    FG_TRY return s->sim.get_solid_force(origin3, out6); FG_CATCH(s)
}

int fg_set_flags(FgSim *s, int32_t flags) {
    if (!s) return FG_EINVAL;
    s->sim.cfg.flags = flags;
    s->sim.dev.enter();
    s->sim.dev.graph_clear();
    return FG_OK;
}

int64_t fg_halo_bytes(FgSim *s) { return s ? s->sim.halo_bytes() : FG_EINVAL; }
int fg_halo_pack(FgSim *s, int32_t face, void *buf) {
    if (!s || !buf) return FG_EINVAL;
    FG_TRY return s->sim.halo_pack(face, buf); FG_CATCH(s)
}
int fg_halo_unpack(FgSim *s, int32_t face, const void *buf) {
    if (!s || !buf) return FG_EINVAL;
    FG_TRY return s->sim.halo_unpack(face, buf); FG_CATCH(s)
}
int fg_peer_export(FgSim *s, FgPeerHandle *out) {
    if (!s || !out) return FG_EINVAL;
    FG_TRY return s->sim.peer_export(out); FG_CATCH(s)
}
int fg_peer_connect_all(FgSim *s, const FgPeerHandle *handles, int32_t n) {
    if (!s || !handles) return FG_EINVAL;
    FG_TRY return s->sim.peer_connect_all(handles, n); FG_CATCH(s)
}
int fg_peer_connect(FgSim *s, const FgPeerHandle *lo, const FgPeerHandle *hi) {
    if (!s) return FG_EINVAL;
    FG_TRY return s->sim.peer_connect(lo, hi); FG_CATCH(s)
}

}  // extern "C"
