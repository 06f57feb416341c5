/*
 * flecsph_b200.h — C ABI of the B200-native FleCSPH per-step particle hot path.
 *
 * This is the drop-in boundary: one `fsph_system`
 * replaces one `body_system<double,D>` of the reference
 * (mpisph/bodies_system.h:38-379).  All particle state lives on the device as
 * fp64 / int64 SoA columns; the host sees it only through fsph_set_column /
 * fsph_get_column (or the AoS marshalling calls) at the points where the
 * reference touches `std::vector<body>` (read_bodies, write_bodies,
 * getLocalbodies, analysis).  Every entry point below cites the reference
 * interface it replaces.
 *
 * Error convention: every call returns 0 on success, non-zero on failure
 * (fsph_last_error gives the text).  The reference aborts (assert /
 * mpi_assert -> MPI_Abort, mpisph/utils.h:34-38); the C++ facade on top of
 * this ABI turns a non-zero code into the same abort.
 *
 * The same ABI is implemented three times, which is what makes the parity
 * tests read like one another:
 *   flecsph_b200/csrc  -> libflecsph_b200.so   the product (sm_100a CUDA, no CPU path)
 *   oracle/            -> oracle_port.py       numpy restatement with the same method names (test infrastructure)
 *   oracle/_ref        -> libflecsph_ref_Dd.so the unmodified reference headers behind
 *                                              this ABI (test infrastructure)
 */
#ifndef FLECSPH_B200_H
#define FLECSPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsph_system fsph_system;

/* sph_kernel keyword, same order as include/params.h:124-133 */
enum fsph_kernel {
  FSPH_KERNEL_CUBIC_SPLINE = 0,
  FSPH_KERNEL_QUINTIC_SPLINE = 1,
  FSPH_KERNEL_WENDLAND_C2 = 2,
  FSPH_KERNEL_WENDLAND_C4 = 3,
  FSPH_KERNEL_WENDLAND_C6 = 4,
  FSPH_KERNEL_GAUSSIAN = 5,
  FSPH_KERNEL_SUPER_GAUSSIAN = 6,
  FSPH_KERNEL_SINC = 7
};
/* eos_type keyword, include/params.h:135-141 */
enum fsph_eos { FSPH_EOS_IDEAL = 0, FSPH_EOS_POLYTROPIC = 1, FSPH_EOS_WD = 2, FSPH_EOS_PPT = 3, FSPH_EOS_NONE = 4 };
/* sph_viscosity keyword, include/params.h:145-148 */
enum fsph_viscosity { FSPH_VISC_CONSTANT = 0, FSPH_VISC_CULLEN = 1 };

/* POD snapshot of the `param::` globals the hot path reads
 * (include/params.h:160-628) plus the `physics::` time globals
 * (include/physics/default_physics.h:31-40). */
typedef struct fsph_params {
  int32_t dim;                    /* gdimension (app/drivers/include/user.h:36) */
  int32_t sph_kernel;             /* params.h:227 */
  int32_t eos_type;               /* params.h:468 */
  int32_t sph_viscosity;          /* params.h:495 */
  double sph_eta;                 /* params.h:212 */
  double sph_sinc_index;          /* params.h:232 */
  double poly_gamma;              /* params.h:478 */
  double sph_viscosity_alpha;     /* params.h:501 */
  double sph_viscosity_beta;      /* params.h:506 */
  double sph_viscosity_epsilon;   /* params.h:511 */
  double timestep_cfl_factor;     /* params.h:183 */
  double initial_dt;              /* params.h:178 */
  double out_screen_dt;           /* params.h:339 */
  double out_scalar_dt;           /* params.h:350 */
  double out_h5data_dt;           /* params.h:366 */
  double fmm_macangle;            /* params.h:542 */
  double gravitational_constant;  /* params.h:628 */
  double box_length, box_width, box_height; /* params.h:258-266 */
  int32_t adaptive_timestep;      /* params.h:188 */
  int32_t sph_variable_h;         /* params.h:244 */
  int32_t sph_update_uniform_h;   /* params.h:238 */
  int32_t evolve_internal_energy; /* params.h:584 */
  int32_t thermokinetic_formulation; /* params.h:591 */
  int32_t enable_fmm;             /* params.h:537 */
  int32_t relaxation_steps;       /* params.h:563 */
  int32_t pressure_updates_number;/* params.h:194 */
  int32_t periodic_boundary_x, periodic_boundary_y, periodic_boundary_z; /* params.h:292-300 */
  int32_t reserved_;
  double sph_viscosity_alpha_max;   /* params.h:516  (Cullen-Dehnen viscosity) */
  double sph_viscosity_l;           /* params.h:523 */
  double sph_viscosity_delta;       /* params.h:529 */
  double poly_gamma2;               /* params.h:483  (piecewise polytrope) */
  double ppt_density_thr;           /* params.h:488 */
  double relaxation_beta;           /* params.h:568  (drag of the relaxation phase) */
  double relaxation_gamma;          /* params.h:572 */
  double relaxation_repulsion_radius; /* params.h:576 */
  double relaxation_repulsion_gamma;  /* params.h:580 */
  /* external forces and potentials (include/physics/eforce.h): the keyword external_force::select parses
   * (eforce.h:358-433; "none", or a comma-separated list of "spherical wall", "airfoil", "spherical density support",
   * "gravity", "orbit", "walls:xyz", "poison"; an empty string means "none") and the parameters its terms read */
  char external_force_type[64];       /* params.h:597 */
  char density_profile[32];           /* params.h:671 ("constant", "parabolic", "mesa"; "from file" is refused) */
  double sphere_radius;               /* params.h:270 */
  double extforce_wall_powerindex;    /* params.h:608 */
  double extforce_wall_steepness;     /* params.h:613 */
  double gravity_acceleration_constant; /* params.h:623 */
  double zero_potential_poison_value; /* params.h:603 */
  double mesa_rim_width;              /* params.h:618 */
  double rho_initial, pressure_initial; /* params.h:682,687 */
  double orbital_separation, mass_neutron_star, mass_white_dwarf; /* params.h:637-647 */
  double airfoil_size, airfoil_thickness, airfoil_camber;         /* params.h:761-769 */
  double airfoil_anchor_x, airfoil_anchor_y, airfoil_attack_angle;/* params.h:773-781 */
} fsph_params;

/* Fill `p` with the reference's defaults (include/params.h) for `dim`. */
void fsph_default_params(fsph_params * p, int32_t dim);

/* Particle columns (fields of body_u, include/physics/body.h:221-251, and of
 * entity, include/tree_topology/tree_types.h:376-381).  Vector fields are
 * stored as `dim` separate contiguous columns; fsph_set/get_column move them
 * as an [n][dim] row-major host array. */
enum fsph_column {
  FSPH_COL_POS = 0,     /* coordinates_      [n][dim] f64 */
  FSPH_COL_VEL,         /* velocity_         [n][dim] f64 */
  FSPH_COL_VELHALF,     /* velocityhalf_     [n][dim] f64 */
  FSPH_COL_ACC,         /* acceleration_     [n][dim] f64 */
  FSPH_COL_GACC,        /* g_acceleration_   [n][dim] f64 */
  FSPH_COL_MASS,        /* mass_             [n] f64 */
  FSPH_COL_H,           /* radius_ = full kernel support radius [n] f64 */
  FSPH_COL_RHO,         /* density_          */
  FSPH_COL_P,           /* pressure_         */
  FSPH_COL_U,           /* internalenergy_   */
  FSPH_COL_E,           /* totalenergy_      */
  FSPH_COL_CS,          /* soundspeed_       */
  FSPH_COL_DUDT,        /* dudt_             */
  FSPH_COL_DEDT,        /* dedt_             */
  FSPH_COL_DT,          /* dt_               */
  FSPH_COL_ALPHA,       /* alpha_            */
  FSPH_COL_VSIG,        /* signalspeed_      */
  FSPH_COL_GPOT,        /* g_potential_      */
  FSPH_COL_ADIABATIC,   /* adiabatic_        */
  FSPH_COL_ID,          /* id_               [n] i64 */
  FSPH_COL_KEY,         /* key_              [n] u64 (read-only: written by update_iteration) */
  FSPH_COL_TYPE,        /* type_ (0 NORMAL, 1 WALL) [n] i32 */
  FSPH_COL_NNEIGH,      /* neighbour count of the last sweep [n] i32 (read-only) */
  FSPH_COL_YE,          /* electronfraction_ (white-dwarf EOS) [n] f64 */
  FSPH_COL_DIVV,        /* divergenceV_      (Cullen-Dehnen viscosity) */
  FSPH_COL_DDIVVDT,     /* dDivVdt_          */
  FSPH_COL_TRIGGER,     /* trigger_          (outputs of viscosity::compute_alpha) */
  FSPH_COL_XI,          /* xi_               */
  FSPH_COL_TRACESS,     /* traceSS_          */
  FSPH_COL_GRADV,       /* gradv_ (first component of the velocity-gradient matrix) */
  FSPH_COL_COUNT_
};

/* apply_in_smoothinglength(f) callables (mpisph/bodies_system.h:304) the
 * in-scope drivers pass (app/drivers/hydro/main_driver.cc:118-233). */
enum fsph_sweep_op {
  FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED = 0, /* physics::compute_density_pressure_soundspeed, default_physics.h:309 */
  FSPH_SWEEP_ACCELERATION = 1,                /* physics::compute_acceleration, default_physics.h:348 */
  FSPH_SWEEP_DUDT = 2,                        /* physics::compute_dudt, default_physics.h:421 */
  FSPH_SWEEP_DEDT = 3,                        /* physics::compute_dedt, default_physics.h:502 */
  FSPH_SWEEP_NEIGHBORS_ONLY = 4,              /* empty functor: neighbour search only (app/drivers/tree/main_driver.cc:103) */
  FSPH_SWEEP_COMPUTE_ALPHA = 5,               /* viscosity::compute_alpha, viscosity.h:310 (Cullen-Dehnen) */
  FSPH_SWEEP_SHORT_RANGE_REPULSION = 6        /* physics::add_short_range_repulsion, default_physics.h:718 */
};

/* apply_all(f) callables (mpisph/bodies_system.h:318). */
enum fsph_apply_op {
  FSPH_APPLY_EOS_INIT = 0,                 /* eos::init, eos/eos.h:157 */
  FSPH_APPLY_SET_TOTAL_ENERGY,             /* physics::set_total_energy, default_physics.h:327 */
  FSPH_APPLY_SAVE_VELOCITYHALF,            /* integration::save_velocityhalf, integration.h:52 */
  FSPH_APPLY_LEAPFROG_KICK_V,              /* integration::leapfrog_kick_v, integration.h:65 */
  FSPH_APPLY_LEAPFROG_KICK_U,              /* integration::leapfrog_kick_u, integration.h:80 */
  FSPH_APPLY_LEAPFROG_KICK_E,              /* integration::leapfrog_kick_e, integration.h:94 */
  FSPH_APPLY_LEAPFROG_DRIFT,               /* integration::leapfrog_drift, integration.h:106 */
  FSPH_APPLY_RECOMPUTE_P_CS,               /* physics::recompute_pressure_soundspeed, default_physics.h:153 */
  FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC, /* physics::recompute_pressure_soundspeed_thermokinetic, default_physics.h:169 */
  FSPH_APPLY_COMPUTE_DT,                   /* physics::compute_dt, default_physics.h:602 */
  FSPH_APPLY_CHECK_NANS,                   /* physics::check_nans, default_physics.h:918 */
  FSPH_APPLY_CHECK_NEGATIVITY,             /* physics::check_negativity, default_physics.h:962 */
  FSPH_APPLY_INITIALIZE_ALPHA,             /* viscosity::initialize_alpha, viscosity.h:172 */
  FSPH_APPLY_ADD_DRAG_ACCELERATION,        /* physics::add_drag_acceleration, default_physics.h:702 */
  FSPH_APPLY_ADD_DRAG_DEDT,                /* physics::add_drag_dedt, default_physics.h:573 */
  FSPH_APPLY_ADD_DRAG_DUDT                 /* physics::add_drag_dudt, default_physics.h:587 */
};

/* get_all(f) callables (mpisph/bodies_system.h:335). */
enum fsph_reduce_op {
  FSPH_GET_SET_ADAPTIVE_TIMESTEP = 0,      /* physics::set_adaptive_timestep, default_physics.h:761; out[0] = new dt */
  FSPH_GET_COMPUTE_SMOOTHINGLENGTH,        /* physics::compute_smoothinglength, default_physics.h:825 */
  FSPH_GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH,/* physics::compute_average_smoothinglength, default_physics.h:879 */
  FSPH_GET_TOTAL_MASS,                     /* analysis::compute_total_mass, analysis.h:69; out[0] */
  FSPH_GET_TOTAL_ENERGY,                   /* analysis::compute_total_energy, analysis.h:113; out[0] */
  FSPH_GET_TOTAL_MOMENTUM,                 /* analysis::compute_total_momentum, analysis.h:56; out[0..dim) */
  FSPH_GET_TOTAL_KINETIC_ENERGY,           /* analysis::compute_total_kinetic_energy, analysis.h:131; out[0] */
  FSPH_GET_TOTAL_INTERNAL_ENERGY,          /* analysis::compute_total_internal_energy, analysis.h:147; out[0] */
  FSPH_GET_TOTAL_GRAVITATIONAL_ENERGY,     /* analysis::compute_total_gravitational_energy, analysis.h:160; out[0] */
  FSPH_GET_TOTAL_ANG_MOM,                  /* analysis::compute_total_ang_mom, analysis.h:176; out[0] (2-D) / out[0..3) (3-D) */
  FSPH_GET_SMOOTHINGLENGTH_STATS,          /* diagnostic::compute_smoothinglength_stats, diagnostic.h:71; out = {min, max, sum} */
  FSPH_GET_VELOCITY_STATS,                 /* diagnostic::compute_velocity_stats, diagnostic.h:93; out = {min, max, sum} of |v| */
  FSPH_GET_NEIGHBORS_STATS                 /* diagnostic::compute_neighbors_stats, diagnostic.h:46; out = {min, max, sum} of the counts */
};

/* --- life cycle ---------------------------------------------------------- */
/* body_system<double,D>::body_system() (bodies_system.h:45).  `device` is the
 * CUDA ordinal.  Returns NULL on failure (no CUDA device => failure; there is
 * no CPU path in the product library). */
fsph_system * fsph_create(const fsph_params * p, int32_t device);
void fsph_destroy(fsph_system * s);
const char * fsph_last_error(const fsph_system * s);
/* "cuda-sm100a" for the product, "oracle-port" / "reference" for the checkers. */
const char * fsph_backend(void);

/* --- particle storage: read_bodies / getLocalbodies (bodies_system.h:83,345) */
int32_t fsph_resize(fsph_system * s, int64_t n);
int64_t fsph_nbodies(const fsph_system * s);                      /* bodies_system.h:349 */
int32_t fsph_set_column(fsph_system * s, int32_t col, const void * host);
int32_t fsph_get_column(fsph_system * s, int32_t col, void * host);
/* One contiguous component of a column ([n] values; comp = 0 for scalar columns): the zero-staging path the
 * C++ facade uses for SoA host buffers (pinned memory makes it a straight DMA). */
int32_t fsph_set_component(fsph_system * s, int32_t col, int32_t comp, const void * host);
int32_t fsph_get_component(fsph_system * s, int32_t col, int32_t comp, void * host);
/* AoS marshalling of the reference's 368-byte `body` records: offsets[col]
 * is offsetof(body, field) obtained in the C++ facade, -1 to skip a column. */
int32_t fsph_upload_aos(fsph_system * s, const void * bodies, int64_t n, int64_t stride, const int64_t * offsets);
int32_t fsph_download_aos(fsph_system * s, void * bodies, int64_t stride, const int64_t * offsets);

/* --- time globals: physics::dt / totaltime / iteration (default_physics.h:31-40) */
int32_t fsph_set_time(fsph_system * s, double dt, double totaltime, int64_t iteration);
int32_t fsph_get_time(const fsph_system * s, double * dt, double * totaltime, int64_t * iteration);
int32_t fsph_advance_time(fsph_system * s);                       /* physics::advance_time, default_physics.h:853 */
/* All eight time globals at once, in the order of default_physics.h:32-39: dt, dt_saved, totaltime, totaltime_next,
 * t_h5data_output, t_screen_output, t_scalar_output (st[0..7)) and iteration.  The C++ facade mirrors the host
 * globals into the system before a device call and back after set_adaptive_timestep. */
int32_t fsph_set_time_state(fsph_system * s, const double * st, int64_t iteration);
int32_t fsph_get_time_state(const fsph_system * s, double * st, int64_t * iteration);

/* --- the hot path --------------------------------------------------------- */
int32_t fsph_update_iteration(fsph_system * s);                   /* bodies_system.h:152: range, keys, sort, tree */
int32_t fsph_reset_ghosts(fsph_system * s);                       /* bodies_system.h:273 */
int32_t fsph_sweep(fsph_system * s, int32_t sweep_op);            /* apply_in_smoothinglength, bodies_system.h:304 */
int32_t fsph_apply(fsph_system * s, int32_t apply_op);            /* apply_all, bodies_system.h:318 */
int32_t fsph_reduce(fsph_system * s, int32_t reduce_op, double * out, int32_t nout); /* get_all, bodies_system.h:335 */
int32_t fsph_gravitation_fmm(fsph_system * s);                    /* bodies_system.h:282 */
int32_t fsph_get_range(fsph_system * s, double * lo, double * hi);/* getRange, bodies_system.h:131 */
double fsph_get_smoothinglength(fsph_system * s);                 /* getSmoothinglength, bodies_system.h:112 */

/* One whole non-initial iteration of the hydro / newtonian driver loop
 * (app/drivers/hydro/main_driver.cc:175-277, newtonian/main_driver.cc:184-296)
 * issued as one call so the element-wise passes can be fused between the
 * sweeps; result identical to issuing the individual calls in driver order.
 * `initial` != 0 runs the first-iteration branch (hydro/main_driver.cc:102-173). */
int32_t fsph_step(fsph_system * s, int32_t initial);

/* fsph_step with its results streamed to the host while the step is still running
 * (the host-resident body array of the reference, bodies_system.h:352-358, wants
 * every field back after an iteration): component `comp` of column `col` of the
 * bodies this rank owns goes to the pinned host buffer `host` on a side stream as
 * soon as the iteration has written it for the last time -- ids, masses and
 * positions right after update_iteration, accelerations after the acceleration
 * sweep (and gravity), velocities after the second kick, the rest at the end.
 * Returns when every copy has arrived.  Same results as fsph_step followed by
 * fsph_get_component for each entry. */
typedef struct fsph_stream_out {
  int32_t col, comp;
  void * host;
} fsph_stream_out;
int32_t fsph_step_streamed(fsph_system * s, int32_t initial, const fsph_stream_out * out, int32_t nout);

/* The same with the iteration's INPUTS coming from host buffers as well, pipelined against the neighbouring iterations:
 * the reference's loop keeps the bodies in host memory (bodies_system.h:352-358), so a drop-in that is handed host
 * buffers moves every field to the device before an iteration and back after it.  Component `comp` of column `col` of
 * the owned bodies ([n] values, as fsph_set_component) is copied from `host` on an upload stream; each copy waits (on
 * the device, with events) for the kernels of the previous iteration and for the previous iteration's device-to-host
 * copy of the same column or into the same host buffer -- so with the buffers of iteration k handed back as the inputs
 * of iteration k+1 the uploads of k+1 run while the tail of k's results is still on its way down (PCIe is full duplex),
 * and every byte still goes host -> device -> host every iteration.  The call returns when the iteration's kernels are
 * done, NOT when its results have arrived: call fsph_stream_wait before reading the host buffers (the next
 * fsph_step_pipelined orders itself after the copies on the device; any other call on the system waits for them).
 * Same results as fsph_set_component for every input, fsph_step, fsph_get_component for every output. */
typedef struct fsph_stream_in {
  int32_t col, comp;
  const void * host;
} fsph_stream_in;
int32_t fsph_step_pipelined(fsph_system * s, int32_t initial, const fsph_stream_in * in, int32_t nin, const fsph_stream_out * out,
  int32_t nout);
int32_t fsph_stream_wait(fsph_system * s);

/* --- introspection used by the parity tests ------------------------------- */
/* Octree node table, sorted by key: tree_topology::cofm_ (tree_topology.h:2103)
 * with node fields of cofm_u (tree_types.h:51-141). */
int64_t fsph_tree_num_nodes(fsph_system * s);
int32_t fsph_tree_get_nodes(fsph_system * s, uint64_t * key, double * mass, double * com /*[nn][dim]*/,
  double * radius, double * lap, double * bmin /*[nn][dim]*/, double * bmax /*[nn][dim]*/, int64_t * sub_entities);
/* Root node of the tree: tree()->root_node()->coordinates() / mass() (analysis.h:396, mpisph/test/tree3d.cc:85). */
int32_t fsph_tree_root(fsph_system * s, double * com /*[dim]*/, double * mass);
/* Depth of every particle's leaf cell in the tree (in storage order). */
int32_t fsph_tree_get_leaf_depth(fsph_system * s, int32_t * depth);
/* Neighbour lists of the last sweep in CSR form, neighbour *ids* sorted
 * ascending per particle; rows in storage (key-sorted) order.
 * offsets has n+1 entries; call with ids == NULL to get only offsets. */
int32_t fsph_neighbor_lists(fsph_system * s, int64_t * offsets, int64_t * ids);

/* --- initial data on the device -----------------------------
 * The reference's generators are host-serial (`assert(size == 1)`, app/id_generators/sedov/main.cc:182) and write
 * H5part files; at 64-128 M bodies that is minutes of host time before the first step.  These calls build the same
 * bodies in the system's columns directly:
 *   fsph_generate_lattice  particle_lattice::generator_lattice_3d + in_domain_3d (include/lattice.h:224-293,107-133):
 *                          lattice_type 0 rectangular / 1 HCP / 2 FCC, domain_type 0 box / 1 sphere; positions in
 *                          generation order, ids = id_offset + index.  slab_k of slab_K selects one contiguous group of
 *                          z layers (one per rank of a multi-GPU run).  Resizes the system; *n_out = bodies made.
 *   fsph_count_within      bodies with |x - c| < radius (blast region, sedov/main.cc:283-295).
 *   fsph_id_fill           f = {mass, h, rho, u, u_blast_add, blast_radius, gamma, dt, radial_velocity, rho_eos, pressure}:
 *                          uniform mass / h / rho / dt, u (+ u_blast_add inside the blast radius, sedov/main.cc:297-304),
 *                          P = rho_eos u (gamma - 1) -- the generators compute it with the renormalised profile density,
 *                          which noh then replaces by rho_initial in the stored column (noh/main.cc:251-261,290) -- or the
 *                          constant `pressure` when that is not NaN (collapse/main.cc:186), v = radial_velocity r / |r|
 *                          (noh/main.cc:272-276). */
int32_t fsph_generate_lattice(fsph_system * s, int32_t lattice_type, int32_t domain_type, const double * bmin, const double * bmax,
  double sep, int32_t slab_k, int32_t slab_K, int64_t id_offset, int64_t * n_out);
int32_t fsph_count_within(fsph_system * s, const double * center, double radius, int64_t * out);
int32_t fsph_id_fill(fsph_system * s, const double * f11);

/* Device micro-benchmarks (roofline denominators measured on the GPU the run uses; SURVEY.md 8d asks for the fp64
 * one because MEASURED_PEAKS.json has none): sustained DFMA rate in TFLOP/s and 128-bit copy bandwidth in GB/s
 * (read + write bytes).  < 0 on failure.  No reference counterpart (the reference has no GPU path). */
double fsph_measure_fp64_tflops(int32_t device, int32_t reps);
double fsph_measure_copy_gbs(int32_t device, int64_t bytes, int32_t reps);

/* Per-phase device times of the last fsph_step in milliseconds (CUDA events);
 * names is a ';'-separated list, returns the number of phases. */
int32_t fsph_phase_times(fsph_system * s, double * ms, int32_t cap, const char ** names);

/* --- several ranks: Morton-key domain decomposition ------------------------
 * The reference distributes a body_system over MPI ranks (MPI_Init in
 * app/drivers/hydro/main.cc:38; MPI_Comm_rank/size in bodies_system.h:47-52):
 * psort.h:227-322 (sample sort + MPI_Alltoallv), tree_topology.h:1549-1744
 * (share_nodes_) and :1276-1542 (remote branch requests).  Here a system joins
 * a communicator once, after fsph_create; from then on fsph_update_iteration,
 * fsph_reset_ghosts, fsph_reduce and fsph_step are collective calls (every rank
 * must issue them in the same order) and each rank holds the bodies of one
 * contiguous range of the global key order.  fsph_nbodies / fsph_get_column see
 * the owned bodies only.  Results equal those of one rank.
 *   nccl : one process per GPU; rank 0 calls fsph_comm_unique_id and ships the
 *          128 bytes to the other ranks (any side channel), every rank then calls
 *          fsph_comm_init_nccl.  libnccl.so.2 is resolved at run time.
 *   local: several systems of one process, one host thread each, on one or more
 *          GPUs (used by the single-GPU test-suite). */
int32_t fsph_comm_unique_id(char * out128);
int32_t fsph_comm_init_nccl(fsph_system * s, const char * unique_id, int32_t rank, int32_t nranks);
void * fsph_local_group_create(int32_t nranks);
void fsph_local_group_destroy(void * group);
int32_t fsph_comm_init_local(fsph_system * s, void * group, int32_t rank);
int32_t fsph_comm_rank(const fsph_system * s);
int32_t fsph_comm_size(const fsph_system * s);
int64_t fsph_nghosts(const fsph_system * s);  /* remote copies currently held (after update_iteration) */
/* Host-side planning used by fsph_update_iteration on several ranks (pure functions, no device):
 * weighted splitters from `nsamples` regular key samples per rank (psort.h:258-301), and the move of a splitter
 * down to the first key of the <= 128-body work cell (tree_topology.h:221-235) that straddles it, given the keys
 * held by all ranks just below / at-or-above it. */
void fsph_dd_choose_splitters(int32_t nranks, int32_t nsamples, const uint64_t * sample_keys, const int64_t * counts,
  uint64_t * splitters /*[nranks-1]*/);
uint64_t fsph_dd_align_splitter(int32_t dim, const uint64_t * below, int32_t nbelow, const uint64_t * above, int32_t nabove,
  uint64_t splitter);

/* Number of device kernels this library has launched in the calling process (0 for the CPU checkers). */
uint64_t fsph_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLECSPH_B200_H */
