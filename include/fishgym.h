/* fishgym.h — the C ABI of the coupled fluid step (drop-in boundary).
 *
 * Reference interface replaced: none exists.  /root/reference/README.md:2-3 names the path
 * ("physical articulated underwater agent interaction with fluid", "coupled interation between
 * agents and fluid") and README.md:14 only promises "control through Python interface"; the
 * shape of this boundary is the one BASELINE.json:5 (`north_star`) prescribes: a Gym-style Python
 * env over a thin C ABI into a CUDA library, no PyTorch on the sim path.
 *
 * Two shared libraries export exactly these symbols:
 *   gym-fish_b200/csrc/libfishgym_cuda.so   the product: CUDA sm_100a, fp32 populations
 *   oracle/libfishgym_oracle.so             the checker: OpenMP C++, fp64 (tests/bench baseline only)
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns FG_OK (0) or a negative FG_E* code;
 *     fg_last_error() gives the message; no exception or abort crosses the boundary.
 *   - the library owns everything behind the opaque FgSim*; the caller owns every buffer it passes
 *     and the library never keeps a caller pointer past the call.
 *   - host arrays are C-contiguous: scalars [nz][ny][nx] (x fastest, z slowest = swim axis),
 *     vectors [3][nz][ny][nx], populations [19][nz][ny][nx].  With z-slab decomposition
 *     (n_ranks > 1) nz in array shapes is the LOCAL slab height nz/n_ranks.
 *   - lattice: D3Q19 in d'Humieres ordering; lattice units dx = dt = 1.
 *   - one FgSim is single-caller; distinct handles are independent.
 */
#ifndef FISHGYM_H
#define FISHGYM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_ABI_VERSION 8
#define FG_Q 19

/* error codes */
#define FG_OK        0
#define FG_EINVAL   -1   /* bad argument / config */
#define FG_ENOMEM   -2   /* host or device allocation failed */
#define FG_ECUDA    -3   /* CUDA runtime error (message has the cudaError string) */
#define FG_ESTATE   -4   /* call not valid in the current state */
#define FG_ENOTSUP  -5   /* feature not supported by this backend */
#define FG_EPEER    -6   /* peer (multi-GPU) setup or exchange failed */

/* collision models */
#define FG_BGK 0
#define FG_MRT 1

/* face ids and boundary conditions (inlet/outlet on z faces only: z is the swim/flow axis) */
#define FG_XLO 0
#define FG_XHI 1
#define FG_YLO 2
#define FG_YHI 3
#define FG_ZLO 4
#define FG_ZHI 5
#define FG_BC_PERIODIC 0
#define FG_BC_WALL     1   /* half-way bounce-back, optional tangential wall velocity */
#define FG_BC_INLET    2   /* equilibrium at (inlet_rho, inlet_u) pulled from outside the face */
#define FG_BC_OUTLET   3   /* zero-gradient: pull source clamped to the last plane */

/* FgConfig.flags */
#define FG_FLAG_NO_OVERLAP 1   /* multi-GPU: halo after the full-slab kernel (overlap off) */
#define FG_FLAG_PROFILE    2   /* bracket every bulk stream-collide launch with events -> FgStats.collide_ms/_cells (no graphs, no split) */
#define FG_FLAG_FUSED_IB   8   /* run the IB phases as ONE cooperative kernel with grid barriers (measured slower on B200: r1) */
#define FG_FLAG_NO_GRAPHS  4   /* launch every kernel directly instead of replaying per-substep CUDA graphs */
#define FG_FLAG_NO_SWEEP_FLIP 32 /* sweep the planes upwards in every step (default: odd steps downwards, for L2 reuse between steps) */
#define FG_FLAG_FUSED_PAIRS 64 /* even step + following odd step as ONE L2-resident wavefront launch (no bodies, one rank); halves DRAM traffic but measured slower on B200 (r1; so were a launch-level wavefront of plane chunks and a persistent-CTA form in r2, both removed: profiles/r2_summary.md) */
#define FG_FLAG_NO_XWARP 128   /* x walls: predicated wall selects in every thread (default: only the two warps at the row ends run the wall code) */
#define FG_FLAG_SYNC_STEP 256  /* fg_step returns only when all its device work has finished (default: when wrenches / obs are there) */
#define FG_FLAG_IB_TILE_SPREAD 512 /* A/B: force spreading staged per CTA in a shared-memory table keyed by band cell, one global reduction per touched cell (default: one red.add per stencil node); pays only for spatially ordered marker lists (profiles/r2_summary.md) */
#define FG_FLAG_EVEN_VEC4 1024 /* even steps take 4 cells per thread with 128-bit loads / stores (needs nx % 4 == 0, no obstacles); measured +1.3 ... +2.4 %, less than the 2-cell form (profiles/r2_summary.md) */
#define FG_FLAG_EVEN_VEC2 2048 /* ... 2 cells per thread with 64-bit accesses for any even nx (the DEFAULT when nx is a multiple of 256 or 128 / 64: +2.6 ... +3.2 %) */
#define FG_FLAG_EVEN_SCALAR 8192 /* even steps with the scalar one-cell-per-thread kernel everywhere (A/B against the default) */
#define FG_FLAG_ODD_VEC2 16384 /* bulk odd steps (x periodic or between walls, no obstacles) with 2 cells per thread, 64-bit accesses for the 9 slots with c_x = 0, for any even nx (the DEFAULT when nx is a multiple of 256 or 128 / 64: +2.6 ... +3.9 %) */
#define FG_FLAG_ODD_SCALAR 32768 /* bulk odd steps with the one-cell-per-thread kernel everywhere (A/B against the default) */
#define FG_FLAG_NO_SPLIT  16   /* collide all planes after the IB kernels (default: planes away from the bodies run beside them) */

typedef struct FgConfig {
    int32_t struct_size;      /* = sizeof(FgConfig); checked by fg_create */
    int32_t nx, ny, nz;       /* GLOBAL lattice size; nz % n_ranks == 0 */
    int32_t collision;        /* FG_BGK | FG_MRT */
    int32_t bc[6];            /* per face, FG_BC_* */
    int32_t n_ranks, rank;    /* z-slab decomposition: this handle owns slab `rank` */
    int32_t device;           /* CUDA device ordinal (ignored by the oracle) */
    int32_t max_markers;      /* capacity for IB markers (0 = no immersed boundary) */
    int32_t max_links;        /* capacity for rigid links markers can belong to */
    int32_t flags;            /* FG_FLAG_* */
    int32_t split_min_cells;  /* plane split only when at least this many cells lie in far planes; 0 => 1<<20 (~25 us of work) */
    int32_t pair_lag;         /* FG_FLAG_FUSED_PAIRS: planes between the even and the odd wavefront; 0 => chosen from the plane size */
    int32_t ib_iterations;    /* direct-forcing passes per substep (multi-direct forcing): 0 or 1 => one pass F_k = 2 rho0 (U_d - U*_k);
                               * n > 1 => n - 1 Jacobi corrections dF_k = 2 rho0 (U_d - U*_k) - sum_x F(x) delta_h(x - X_k), F_k += dF_k, spread dF_k
                               * (the no-slip residual at the markers shrinks with every pass); at most 16.  With bodies across z-slab faces (fg_peer_connect_all)
                               * every pass exchanges the partial sums of face-crossing markers with the z-neighbours, like U*. */
    double  tau;              /* relaxation time; nu = (tau - 1/2)/3 */
    double  mrt_rates[19];    /* MRT relaxation rates per moment; all zero => SURVEY.md A3 defaults */
    double  wall_u[6][3];     /* wall velocity per face (used where bc == WALL) */
    double  inlet_u[3];
    double  inlet_rho;        /* 0 => 1.0 */
    double  body_force[3];    /* uniform Guo body force density */
    double  reserved_d[4];
} FgConfig;

typedef struct FgStats {
    int64_t steps;            /* fluid steps taken since create/reset */
    int64_t cells;            /* local lattice cells (nx*ny*nz_local) */
    double  last_step_ms;     /* device time (CUDA events) of the latest fg_step call whose work has finished (fg_sync first to be sure) */
    double  last_mlups;       /* cells * n_substeps / last_step_ms / 1e3 */
    int64_t kernel_launches;  /* kernels of this library launched since create */
    int32_t n_markers, n_links;
    int32_t band_cells;       /* cells inside marker stencils in the last step */
    int32_t parity;           /* AA-pattern parity of the next step (0 even / 1 odd); oracle: 0 */
    double  collide_ms;       /* FG_FLAG_PROFILE: summed device time of the bulk stream-collide launches of the last fg_step */
    int64_t collide_launches; /* ... how many launches that covers */
    double  ib_ms;            /* FG_FLAG_PROFILE: summed device time of the immersed-boundary kernels of the last fg_step */
    int64_t collide_cells;    /* ... and how many cell updates those launches did (thin wall-row launches are not in either) */
    int64_t split_substeps;   /* substeps since create whose far-plane collide ran beside the IB kernels (plane split) */
    int64_t pair_substeps;    /* substeps since create that ran as half of a fused even+odd pair (StreamCollidePair) */
    int64_t graph_launches;   /* CUDA graph launches since create (substeps replayed or captured-and-launched as one graph; direct launches are not in here) */
} FgStats;

/* articulated swimmer description: a planar chain of n_links ellipsoid links, yawing joints */
typedef struct FgFishDesc {
    int32_t n_links;          /* >= 1; n_joints = n_links - 1 */
    int32_t markers_per_link; /* target count; actual chosen by the surface sampler */
    double  link_len[8];      /* full length along the body axis (lattice units) */
    double  link_rad[8];      /* half-width (y and lateral radius) */
    double  root_pos[3];      /* position of the head link centre, lattice coords (x,y,z) */
    double  heading;          /* yaw of the body axis about +y, radians; 0 => nose towards -z */
    double  density_ratio;    /* body/fluid density (mass and inertia of the ellipsoids) */
    double  joint_gain;       /* PD stiffness: joint rate = gain*(target - angle) clipped to joint_rate_max */
    double  joint_limit;      /* |joint angle| <= limit, radians */
    double  joint_rate_max;   /* radians per lattice step */
    int32_t free_root;        /* 1: root integrates hydrodynamic wrench (planar x-z + yaw); 0: pinned */
    int32_t reserved;
} FgFishDesc;

typedef struct FgSim FgSim;

/* opaque blob a rank publishes so that other ranks can map its lattice, flags and IB exchange buffer (CUDA IPC) */
typedef struct FgPeerHandle { unsigned char bytes[320]; } FgPeerHandle;

int         fg_abi_version(void);
const char *fg_backend_name(void);                 /* "cuda-sm100a" | "oracle-fp64" */
const char *fg_last_error(const FgSim *sim);       /* sim may be NULL: last create error */

/* ---- lifecycle ---- */
int fg_config_default(FgConfig *cfg);              /* fills struct_size, BGK, tau 0.8, periodic, 1 rank */
int fg_create(const FgConfig *cfg, FgSim **out);
int fg_destroy(FgSim *sim);
int fg_reset(FgSim *sim, uint64_t seed);           /* f <- f_eq(1,0); bodies to initial pose; step counter 0 */

/* ---- fluid state (local slab, natural layout = populations ARRIVING at the cell at time t) ---- */
int fg_set_fields(FgSim *sim, const float *rho, const float *u);      /* f <- f_eq(rho,u) */
int fg_get_fields(FgSim *sim, float *rho, float *u);                  /* rho = sum f, u = sum c f / rho (no force) */
int fg_get_fields_f64(FgSim *sim, double *rho, double *u);            /* same, double out (oracle: exact) */
int fg_set_populations(FgSim *sim, const float *f19);
int fg_get_populations(FgSim *sim, float *f19);
int fg_set_solid(FgSim *sim, const uint8_t *solid_global);            /* [nz_global][ny][nx], 1 = solid; NULL clears */

/* ---- immersed boundary: Lagrangian markers (prescribed, or generated by fish bodies) ----
 * The read-outs below describe the immersed-boundary pass of the LAST step.  Between fg_set_markers and the next fg_step
 * there is none for the new marker set: index map, marker forces / velocities and link wrenches read as zero. */
int fg_set_markers(FgSim *sim, int32_t n, const float *X, const float *U,
                   const float *dV, const int32_t *link_id);          /* X,U: [n][3]; dV,link_id: [n] */
int fg_set_link_origins(FgSim *sim, int32_t n_links, const double *origin3); /* torque reference points */
int fg_get_index_map(FgSim *sim, int32_t *base3, int32_t *owner);     /* base: [n][3] = floor(X)-1; owner rank: [n] */
int fg_get_marker_forces(FgSim *sim, float *F3);                      /* force density on fluid at markers, last step */
int fg_get_marker_velocities(FgSim *sim, float *Ustar3);              /* interpolated unforced fluid velocity, last step */
int fg_get_link_wrenches(FgSim *sim, double *w6);                     /* [n_links][6]: hydrodynamic force, torque ON the link */
int fg_get_force_field(FgSim *sim, float *F);                         /* [3][nz][ny][nx] Eulerian IB force of the last step */
/* fluid state at arbitrary points (velocity-probe observations): out[n][4] = (rho, u) of the populations arriving at
 * time t, interpolated with the same 4-point delta as the markers; nodes outside the domain / slab contribute 0 */
int fg_probe(FgSim *sim, int32_t n, const float *X, float *out4);

/* ---- articulated bodies, integrated on the host every substep ---- */
int fg_add_fish(FgSim *sim, const FgFishDesc *desc, int32_t *fish_id);
int fg_set_action(FgSim *sim, const float *a, int32_t n);             /* joint targets in [-1,1], all fish concatenated */
int fg_get_obs(FgSim *sim, float *obs, int32_t n);                    /* per fish: pos3, heading, vel3, yaw rate, joints, joint rates */
int fg_obs_size(FgSim *sim);
int fg_action_size(FgSim *sim);
int fg_get_markers(FgSim *sim, float *X, float *U, int32_t *link_id, int32_t cap); /* returns count or <0 */

/* ---- stepping ---- */
/* fg_step returns when the outputs it exposes synchronously are there — link wrenches, marker forces' inputs, observations;
 * the stream-collide of the last substep may still be running.  Every later call is ordered behind it, fg_sync() waits. */
int fg_step(FgSim *sim, int32_t n_substeps);
int fg_sync(FgSim *sim);
int fg_get_stats(FgSim *sim, FgStats *out);
/* divergence guard for the FLUID (the observation only covers the bodies): counts the cells of the local slab whose rest
 * population is not finite or has left |f_0 - 1/3| <= 4 — one 4-byte read per cell (~85 us at 512^3 on a B200).  A NaN
 * anywhere in a cell reaches its rest population with the next collision, so a blow-up shows here one step later at most. */
int fg_check_finite(FgSim *sim, int64_t *n_bad_cells);
/* Momentum-exchange force of the fluid ON the obstacle cells (fg_set_solid), SURVEY.md A5: every bounce-back link of the last step
 * hands the obstacle the momentum of the population it reflected, out6[0..2] = sum over fluid cells x and directions i whose
 * source x - c_i is an obstacle of -2 c_i f_i(x, t+1)  (f_i(x, t+1) = the reflected population now arriving at x), out6[3..5] =
 * the torque of those link forces about origin3 (lattice coordinates x, y, global z; applied at the link midpoints x - c_i/2;
 * NULL: zeros).  Domain walls (FG_BC_WALL) are not obstacles.  With z-slabs every rank returns the share of its own fluid
 * cells: sum over ranks.  A read-out (one pass over the slab, synchronous); before the first step it reflects the initial state. */
int fg_get_solid_force(FgSim *sim, const double *origin3, double *out6);
int fg_set_flags(FgSim *sim, int32_t flags);        /* change FgConfig.flags (FG_FLAG_PROFILE, FG_FLAG_NO_GRAPHS, FG_FLAG_NO_OVERLAP) */

/* ---- z-slab halos ----
 * host-staged form (works on both backends; used by the CPU gloo tests):
 *   after every fg_step(sim, 1) each rank packs the 5 populations crossing each internal face,
 *   ships them to the neighbour, and the neighbour unpacks.  Messages are opaque bytes of the
 *   backend's own population type (fp32 on CUDA, fp64 in the oracle): fg_halo_bytes() per face.
 * device form (CUDA backend): publish/connect CUDA-IPC handles once; fg_step then pushes halos
 *   over NVLink with peer stores and neighbour flags, no host involvement. */
int64_t fg_halo_bytes(FgSim *sim);                                    /* bytes per face message: 5*nx*ny*sizeof(pop) */
int fg_halo_pack(FgSim *sim, int32_t face /*FG_ZLO|FG_ZHI*/, void *buf);
int fg_halo_unpack(FgSim *sim, int32_t face, const void *buf);
int fg_peer_export(FgSim *sim, FgPeerHandle *out);
int fg_peer_connect(FgSim *sim, const FgPeerHandle *zlo_neighbour, const FgPeerHandle *zhi_neighbour);
/* all ranks' handles, indexed by rank: z-neighbours are derived from them, and immersed bodies may then cross slab
 * faces (partial marker velocities go to the face neighbour, link wrenches are all-gathered, both by peer stores).
 * Every rank must hold the same (replicated) marker list / fish. */
int fg_peer_connect_all(FgSim *sim, const FgPeerHandle *handles, int32_t n_handles);

#ifdef __cplusplus
}
#endif
#endif /* FISHGYM_H */
