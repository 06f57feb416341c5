// fsph_capi.cu -- the C ABI of include/flecsph_b200.h over the device path.  No CPU fallback: without a CUDA
// device fsph_create fails.  Each entry point cites the reference interface it replaces in the header.
#include <cstring>
#include <stdexcept>
#include <algorithm>
#include "fsph_internal.h"

using namespace fsph;

struct fsph_system : public System {};

static thread_local std::string g_create_error;

namespace {

const int VEC_FIELD[5] = {F_POS, F_VEL, F_VELH, F_ACC, F_GACC};

int scalar_field_any(int col) {
  switch(col) {
    case FSPH_COL_MASS: return F_MASS;
    case FSPH_COL_H: return F_H;
    case FSPH_COL_RHO: return F_RHO;
    case FSPH_COL_P: return F_P;
    case FSPH_COL_U: return F_U;
    case FSPH_COL_E: return F_E;
    case FSPH_COL_CS: return F_CS;
    case FSPH_COL_DUDT: return F_DUDT;
    case FSPH_COL_DEDT: return F_DEDT;
    case FSPH_COL_DT: return F_DT;
    case FSPH_COL_ALPHA: return F_ALPHA;
    case FSPH_COL_VSIG: return F_VSIG;
    case FSPH_COL_GPOT: return F_GPOT;
    case FSPH_COL_ADIABATIC: return F_ADIA;
    case FSPH_COL_YE: return F_YE;
    case FSPH_COL_DIVV: return F_DIVV;
    case FSPH_COL_DDIVVDT: return F_DDIVVDT;
    case FSPH_COL_TRIGGER: return F_TRIGGER;
    case FSPH_COL_XI: return F_XI;
    case FSPH_COL_TRACESS: return F_TRACESS;
    case FSPH_COL_GRADV: return F_GRADV;
    default: return -1;
  }
}

// F64 index of a scalar column of this system; -1 for non-scalar columns and for optional columns it did not allocate
int scalar_field(const System & s, int col) {
  const int f = scalar_field_any(col);
  return f >= 0 && s.col_used[f] ? f : -1;
}

void finish_phases(System & s) {
  if(!s.timing) return;
  cudaStreamSynchronize(s.stream);
  s.phase_ms.clear();
  s.phase_names.clear();
  // merge phases with the same name (e.g. the two leapfrog kicks)
  std::vector<std::string> names;
  for(auto & p : s.phases) {
    float ms = 0;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
    auto it = std::find(names.begin(), names.end(), p.name);
    if(it == names.end()) {
      names.push_back(p.name);
      s.phase_ms.push_back(ms);
    }
    else s.phase_ms[it - names.begin()] += ms;
  }
  for(size_t i = 0; i < names.size(); ++i) s.phase_names += (i ? ";" : "") + names[i];
  s.phases.clear();
}

void invalidate(System & s) {
  s.tree_valid = false;
  s.search_valid = false;
}

void do_update_iteration(System & s) {
  // body_system::update_iteration, bodies_system.h:152-229 (non-periodic)
  if(s.par.periodic_boundary_x || s.par.periodic_boundary_y || s.par.periodic_boundary_z) {
    PhaseScope p(s, "boundary");
    periodic_walls(s);
  }
  if(s.comm) {
    // several ranks: range all-reduce, sample sort + all-to-all, halo of remote bodies and tree nodes (fsph_dist.cu)
    PhaseScope p(s, "decomposition");
    dist_update_iteration(s);
    return;
  }
  {
    PhaseScope p(s, "range_keys");
    compute_range(s);
    compute_keys(s);
  }
  {
    PhaseScope p(s, "sort");
    sort_bodies(s);
  }
  {
    PhaseScope p(s, "tree");
    build_tree(s);
  }
}

void do_reset_ghosts(System & s) {
  // bodies_system.h:273 -> tree_topology::reset_ghosts (tree_topology.h:153-157): clean + build_tree with the
  // current keys.  The tree is a pure function of (keys, positions, h, mass): rebuilt only if those changed.
  if(!s.tree_valid) throw std::runtime_error("reset_ghosts before update_iteration");
  if(s.comm) {
    PhaseScope p(s, "halo_refresh");
    dist_refresh_ghosts(s);
    return;
  }
  if(s.tree_dirty) {
    PhaseScope p(s, "tree");
    build_tree(s);
  }
}

void do_sweep(System & s, int op) {
  if(s.tree_valid && s.tree_dirty && !s.comm) {
    // positions / h / masses changed through the ABI since the tree was built (outside the driver's order of calls):
    // rebuild it with the current keys, as reset_ghosts would, before searching on it
    PhaseScope p(s, "tree");
    build_tree(s);
  }
  if(!s.search_valid) {
    PhaseScope p(s, "search");
    if(!s.tree_valid) throw std::runtime_error("sweep before update_iteration");
    neighbor_search(s);
  }
  const char * nm = (op == FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED || op == 100) ? "density" : (op == FSPH_SWEEP_ACCELERATION || op == 101) ? "acceleration"
                  : op == FSPH_SWEEP_DUDT ? "dudt" : op == FSPH_SWEEP_DEDT ? "dedt" : op == FSPH_SWEEP_COMPUTE_ALPHA ? "viscosity_alpha"
                  : op == FSPH_SWEEP_SHORT_RANGE_REPULSION ? "repulsion" : "neighbors";
  PhaseScope p(s, nm);
  // with an external potential the internal energy is recovered from e by an element-wise pass (the potential's terms
  // stay out of the sweep kernels); without one the density sweep does it in its epilogue
  if((op == FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED || op == 100) && s.ef.n > 0 && s.par.evolve_internal_energy && s.par.thermokinetic_formulation)
    apply(s, 103);
  sweep(s, op);
  // the last statement of compute_acceleration: acc_a += external_force::acceleration(particle), default_physics.h:404
  if((op == FSPH_SWEEP_ACCELERATION || op == 101) && s.ef.nacc > 0) apply(s, 102);
}

// device address and element size of one component of a C-ABI column
static const void * column_source(System & s, int col, int comp, size_t & esz) {
  esz = 8;
  if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
    if(comp < 0 || comp >= s.dim) throw std::runtime_error("bad component");
    return s.col[VEC_FIELD[col] + comp].p;
  }
  if(scalar_field(s, col) >= 0) return s.col[scalar_field(s, col)].p;
  if(col == FSPH_COL_ID) return s.col_id.p;
  if(col == FSPH_COL_KEY) return s.col_key.p;
  if(col == FSPH_COL_TYPE) { esz = 4; return s.col_type.p; }
  if(col == FSPH_COL_NNEIGH) { esz = 4; return s.col_nneigh.p; }
  throw std::runtime_error("bad column");
}

// fsph_step_streamed: the listed columns (all the remaining ones when `cols` is null) are final from here on; their
// owned range goes to the host on the side stream, ordered after everything queued on the main stream so far
static void stream_flush(System & s, const int * cols, int ncols) {
  if(s.stream_out.empty()) return;
  bool any = false;
  for(auto & e : s.stream_out) {
    if(e.done) continue;
    bool take = cols == nullptr;
    for(int k = 0; k < ncols && !take; ++k) take = e.col == cols[k];
    if(!take) continue;
    if(!any) {
      FSPH_CUDA_OK(cudaEventRecord(s.side_event, s.stream));
      FSPH_CUDA_OK(cudaStreamWaitEvent(s.side_stream, s.side_event, 0));
      any = true;
    }
    size_t esz;
    const char * src = static_cast<const char *>(column_source(s, e.col, e.comp, esz));
    const int64_t first = s.has_ghosts ? s.own_begin : 0, cnt = s.n_owned_or_n();
    FSPH_CUDA_OK(cudaMemcpyAsync(e.host, src + (size_t)first * esz, (size_t)cnt * esz, cudaMemcpyDeviceToHost, s.side_stream));
    if(e.ev) FSPH_CUDA_OK(cudaEventRecord(e.ev, s.side_stream));
    e.done = true;
  }
}

void do_step(System & s, int initial) {
  // app/drivers/hydro/main_driver.cc:102-277 (+ newtonian/main_driver.cc:218-222)
  const fsph_params & P = s.par;
  auto gravity = [&]() {
    // newtonian/main_driver.cc:218-222: right after compute_acceleration (which zeroed g_acc / g_pot)
    if(P.enable_fmm) {
      PhaseScope p(s, "fmm");
      gravitation_fmm(s);
    }
  };
  auto A = [&](int op) { PhaseScope p(s, "elementwise"); apply(s, op); };
  // the element-wise functors after the last sweep of the iteration are issued as one fused pass
  const int recompute_op = P.thermokinetic_formulation ? FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC : FSPH_APPLY_RECOMPUTE_P_CS;
  const int energy_sweep = P.thermokinetic_formulation ? FSPH_SWEEP_DEDT : FSPH_SWEEP_DUDT;
  const int kick_energy = P.thermokinetic_formulation ? FSPH_APPLY_LEAPFROG_KICK_E : FSPH_APPLY_LEAPFROG_KICK_U;
  int tail[8], ntail = 0;
  const bool relaxing = s.iteration < P.relaxation_steps;
  const bool cullen = P.sph_viscosity != FSPH_VISC_CONSTANT;
  // compute_acceleration and what the drivers hang on it (hydro/main_driver.cc:123-133,197-206): the adaptive viscosity
  // sweep before it, the relaxation terms after it
  auto acceleration = [&]() {
    if(cullen) {
      // compute_alpha reads the neighbours' new sound speed and velocity divergence: remote copies are refreshed first
      // (the reference fetches them from their owner while it traverses)
      if(s.comm) do_reset_ghosts(s);
      do_sweep(s, FSPH_SWEEP_COMPUTE_ALPHA);
    }
    do_reset_ghosts(s);
    do_sweep(s, 101);
    if(relaxing) {
      A(FSPH_APPLY_ADD_DRAG_ACCELERATION);
      do_sweep(s, FSPH_SWEEP_SHORT_RANGE_REPULSION);
    }
  };
  auto energy_passes = [&]() {
    for(int m = 1; m <= P.pressure_updates_number; ++m) {
      do_sweep(s, energy_sweep);
      // relaxation phase (hydro/main_driver.cc:152,165,221,235): energy dissipated by the drag
      if(relaxing) A(P.thermokinetic_formulation ? FSPH_APPLY_ADD_DRAG_DEDT : FSPH_APPLY_ADD_DRAG_DUDT);
      if(m < P.pressure_updates_number) {
        A(recompute_op);
        do_reset_ghosts(s);
      }
      else tail[ntail++] = recompute_op;
    }
  };
  auto timestep = [&]() {
    if(P.adaptive_timestep) {
      PhaseScope p(s, "timestep");
      apply(s, FSPH_APPLY_COMPUTE_DT);
      reduce(s, FSPH_GET_SET_ADAPTIVE_TIMESTEP, nullptr, 0);
    }
  };
  static const int cols_after_update[] = {FSPH_COL_ID, FSPH_COL_MASS, FSPH_COL_POS};
  static const int cols_after_density[] = {FSPH_COL_RHO};   // P and c_s are recomputed at the end of the iteration
  static const int cols_after_acceleration[] = {FSPH_COL_ACC, FSPH_COL_GACC, FSPH_COL_VSIG};
  static const int cols_after_energy[] = {FSPH_COL_DEDT, FSPH_COL_DUDT};
  static const int cols_after_kick[] = {FSPH_COL_VEL};
  if(initial) {
    do_update_iteration(s);
    stream_flush(s, cols_after_update, 3);
    A(FSPH_APPLY_EOS_INIT);
    if(P.thermokinetic_formulation) A(FSPH_APPLY_SET_TOTAL_ENERGY);
    if(cullen) A(FSPH_APPLY_INITIALIZE_ALPHA);
    do_sweep(s, 100);
    stream_flush(s, cols_after_density, 1);
    A(FSPH_APPLY_SAVE_VELOCITYHALF);
    acceleration();
    gravity();
    stream_flush(s, cols_after_acceleration, 3);
    stream_flush(s, cols_after_kick, 1);   // the first iteration does not kick
    timestep();
    if(P.evolve_internal_energy) energy_passes();
    stream_flush(s, cols_after_energy, 2);
  }
  else {
    {
      // kick e|u, kick v, save v_half, drift: four functors, one pass
      PhaseScope p(s, "elementwise");
      int ops[4], k = 0;
      if(P.evolve_internal_energy) ops[k++] = kick_energy;
      ops[k++] = FSPH_APPLY_LEAPFROG_KICK_V;
      ops[k++] = FSPH_APPLY_SAVE_VELOCITYHALF;
      ops[k++] = FSPH_APPLY_LEAPFROG_DRIFT;
      apply_fused(s, ops, k);
    }
    {
      // columns the rest of this iteration overwrites before reading: acc, gacc, gpot (acceleration sweep),
      // rho, P (density sweep), vsig, dt, the energy derivative, and u when it is recovered from e
      for(int f = 0; f < NF64; ++f) s.dead_col[f] = false;
      for(int d = 0; d < 3; ++d) s.dead_col[F_ACC + d] = s.dead_col[F_GACC + d] = true;
      s.dead_col[F_GPOT] = s.dead_col[F_RHO] = s.dead_col[F_P] = s.dead_col[F_VSIG] = s.dead_col[F_CSOLD] = true;
      if(P.adaptive_timestep) s.dead_col[F_DT] = true;
      if(P.evolve_internal_energy) {
        if(P.thermokinetic_formulation) s.dead_col[F_DEDT] = s.dead_col[F_U] = true;
        else s.dead_col[F_DUDT] = true;
      }
      s.skip_dead_columns = true;
      try { do_update_iteration(s); }
      catch(...) { s.skip_dead_columns = false; throw; }
      s.skip_dead_columns = false;
    }
    stream_flush(s, cols_after_update, 3);
    do_sweep(s, 100);
    stream_flush(s, cols_after_density, 1);
    // without gravity, external forces and relaxation terms nothing touches the acceleration between the sweep and the second kick: the
    // sweep applies it to each body it finishes (one element-wise pass less)
    const bool fuse_kick = !P.enable_fmm && !relaxing && s.ef.nacc == 0;
    s.fuse_kick_v = fuse_kick;
    try { acceleration(); }
    catch(...) { s.fuse_kick_v = false; throw; }
    s.fuse_kick_v = false;
    gravity();
    stream_flush(s, cols_after_acceleration, 3);
    if(!fuse_kick) A(FSPH_APPLY_LEAPFROG_KICK_V);
    stream_flush(s, cols_after_kick, 1);
    do_reset_ghosts(s);
    if(P.evolve_internal_energy) {
      energy_passes();
      tail[ntail++] = kick_energy;
    }
    stream_flush(s, cols_after_energy, 2);
  }
  {
    PhaseScope p(s, "elementwise");
    if(P.sph_variable_h) tail[ntail++] = 100;
    else if(P.sph_update_uniform_h) {
      // needs a global sum in the middle: flush what is queued, then the reference's own sequence
      apply_fused(s, tail, ntail);
      ntail = 0;
      reduce(s, FSPH_GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH, nullptr, 0);
    }
    tail[ntail++] = FSPH_APPLY_CHECK_NANS;
    tail[ntail++] = FSPH_APPLY_CHECK_NEGATIVITY;
    if(P.adaptive_timestep) tail[ntail++] = FSPH_APPLY_COMPUTE_DT;
    apply_fused(s, tail, ntail);
  }
  if(P.adaptive_timestep) {
    PhaseScope p(s, "timestep");
    reduce(s, FSPH_GET_SET_ADAPTIVE_TIMESTEP, nullptr, 0);
  }
  check_status(s);
  // physics::advance_time, default_physics.h:853-860
  s.iteration++;
  if(P.adaptive_timestep) s.totaltime = s.totaltime_next;
  else s.totaltime = s.totaltime + s.dt;
}

template<typename F>
int32_t guarded(fsph_system * s, F && f) {
  if(!s) return 1;
  try {
    FSPH_CUDA_OK(cudaSetDevice(s->device));
    if(!s->in_pipelined_call && !s->prev_out.empty()) {
      // copies of a pipelined step are still in flight and this is not the next pipelined step: let them land first
      // (any other call may read or overwrite the columns they come from)
      const cudaError_t e = cudaStreamSynchronize(s->side_stream);
      s->prev_out.clear();
      FSPH_CUDA_OK(e);
    }
    f();
    return 0;
  }
  catch(const std::exception & e) {
    s->err = e.what();
    s->phases.clear();
    if(s->comm && s->err.find("local group:") == std::string::npos) s->comm->abort();
    return 1;
  }
}

}  // namespace

extern "C" {

void fsph_default_params(fsph_params * p, int32_t dim) {
  // defaults of include/params.h
  std::memset(p, 0, sizeof(*p));
  p->dim = dim;
  p->sph_kernel = FSPH_KERNEL_WENDLAND_C4;
  p->eos_type = FSPH_EOS_IDEAL;
  p->sph_viscosity = FSPH_VISC_CONSTANT;
  p->sph_eta = 1.5;
  p->sph_sinc_index = 4.0;
  p->poly_gamma = 1.4;
  p->sph_viscosity_alpha = 1.0;
  p->sph_viscosity_beta = 2.0;
  p->sph_viscosity_epsilon = 0.01;
  p->timestep_cfl_factor = 0.25;
  p->initial_dt = 0.001;
  p->gravitational_constant = 1.0;
  p->box_length = 3.0;
  p->box_width = 1.0;
  p->box_height = 1.0;
  p->evolve_internal_energy = 1;
  p->thermokinetic_formulation = 1;
  p->pressure_updates_number = 1;
  p->sph_viscosity_alpha_max = 1.0;
  p->sph_viscosity_l = 0.05;
  p->sph_viscosity_delta = 1.0;
  p->poly_gamma2 = 2.5;
  p->ppt_density_thr = 5e14;
  p->relaxation_beta = 1e-6;
  p->relaxation_gamma = 0.0;
  p->relaxation_repulsion_radius = 0.25;
  p->relaxation_repulsion_gamma = 0.0;
  std::strcpy(p->external_force_type, "none");
  std::strcpy(p->density_profile, "constant");
  p->sphere_radius = 1.0;
  p->extforce_wall_powerindex = 5.0;
  p->extforce_wall_steepness = 1e12;
  p->gravity_acceleration_constant = 9.81;
  p->zero_potential_poison_value = 0.0;
  p->mesa_rim_width = 0.25;
  p->rho_initial = 1.0;
  p->pressure_initial = 1.0;
  p->orbital_separation = 2.5e9;
  p->mass_neutron_star = 1.26 * 1.988435e33;
  p->mass_white_dwarf = 1.10 * 1.988435e33;
  p->airfoil_size = 2.0;
  p->airfoil_thickness = 0.05;
  p->airfoil_camber = 0.1;
  p->airfoil_anchor_x = -1.0;
  p->airfoil_anchor_y = 0.0;
  p->airfoil_attack_angle = 0.0;
}

const char * fsph_backend(void) { return "cuda-sm100a"; }

fsph_system * fsph_create(const fsph_params * p, int32_t device) {
  if(!p || p->dim < 1 || p->dim > 3) { g_create_error = "bad dimension"; return nullptr; }
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) {
    cudaGetLastError();
    g_create_error = "no CUDA device: libflecsph_b200 has no CPU path";
    return nullptr;
  }
  if(p->eos_type < 0 || p->eos_type > FSPH_EOS_NONE) { g_create_error = "unknown eos_type"; return nullptr; }
  if(p->sph_viscosity != FSPH_VISC_CONSTANT && p->sph_viscosity != FSPH_VISC_CULLEN) { g_create_error = "unknown sph_viscosity"; return nullptr; }
  if(p->sph_kernel < 0 || p->sph_kernel > FSPH_KERNEL_SINC) { g_create_error = "unknown sph_kernel"; return nullptr; }
  fsph_system * s = nullptr;
  try {
    FSPH_CUDA_OK(cudaSetDevice(device));
    s = new fsph_system;
    s->par = *p;
    s->dim = p->dim;
    s->device = device;
    s->dt = p->initial_dt;
    s->ef = ef_parse(*p);   // external_force::select, eforce.h:358 (throws on a keyword the reference rejects)
    // optional columns
    for(int f = 0; f < F_FIRST_OPTIONAL; ++f) s->col_used[f] = true;
    s->col_used[F_YE] = p->eos_type == FSPH_EOS_WD;
    for(int f = F_DIVV; f <= F_GRADV; ++f) s->col_used[f] = p->sph_viscosity == FSPH_VISC_CULLEN;
    // kernels::select, kernels.h:753-764
    s->kernel_width = (p->sph_kernel == FSPH_KERNEL_QUINTIC_SPLINE || p->sph_kernel == FSPH_KERNEL_GAUSSIAN ||
                       p->sph_kernel == FSPH_KERNEL_SUPER_GAUSSIAN) ? 3.0 : 2.0;
    FSPH_CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->b_status.reserve(128);
    FSPH_CUDA_OK(cudaMemsetAsync(s->b_status.p, 0, 128, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    return s;
  }
  catch(const std::exception & e) {
    g_create_error = e.what();
    delete s;
    return nullptr;
  }
}

void fsph_destroy(fsph_system * s) {
  if(!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  dist_release(*s);
  delete s->comm;
  s->comm = nullptr;
  release_buffers(*s);
  if(s->side_stream) cudaStreamSynchronize(s->side_stream);
  if(s->up_stream) cudaStreamSynchronize(s->up_stream);
  for(cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
  if(s->side_event) cudaEventDestroy(s->side_event);
  if(s->up_event) cudaEventDestroy(s->up_event);
  if(s->side_stream) cudaStreamDestroy(s->side_stream);
  if(s->up_stream) cudaStreamDestroy(s->up_stream);
  if(s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

const char * fsph_last_error(const fsph_system * s) { return s ? s->err.c_str() : g_create_error.c_str(); }

__global__ void k_iota64(int64_t * p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) p[i] = i;
}

int32_t fsph_resize(fsph_system * s, int64_t n) {
  return guarded(s, [&]() {
    if(n < 0 || n >= (int64_t(1) << 31)) throw std::runtime_error("bad body count");
    const size_t b8 = (size_t)std::max<int64_t>(n, 1) * 8, b4 = (size_t)std::max<int64_t>(n, 1) * 4;
    for(int f = 0; f < NF64; ++f) {
      if(!s->col_used[f]) continue;
      s->col[f].reserve(b8);
      FSPH_CUDA_OK(cudaMemsetAsync(s->col[f].p, 0, b8, s->stream));
    }
    s->col_id.reserve(b8); s->col_key.reserve(b8); s->scratch_col.reserve(b8);
    s->col_type.reserve(b4); s->col_nneigh.reserve(b4);
    FSPH_CUDA_OK(cudaMemsetAsync(s->col_key.p, 0, b8, s->stream));
    FSPH_CUDA_OK(cudaMemsetAsync(s->col_type.p, 0, b4, s->stream));
    FSPH_CUDA_OK(cudaMemsetAsync(s->col_nneigh.p, 0, b4, s->stream));
    if(n > 0) FSPH_LAUNCH(k_iota64, (unsigned)((n + 255) / 256), 256, s->stream, s->col_id.as<int64_t>(), n);
    s->n = n;
    s->n_owned = n;
    s->has_ghosts = false;
    invalidate(*s);
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
  });
}

int64_t fsph_nbodies(const fsph_system * s) { return s ? s->n_owned_or_n() : 0; }

int32_t fsph_set_column(fsph_system * s, int32_t col, const void * host) {
  return guarded(s, [&]() {
    dist_drop_ghosts(*s);
    s->pair_pi_valid = false;
    const int64_t n = s->n;
    const int D = s->dim;
    if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
      const double * h = static_cast<const double *>(host);
      std::vector<double> tmp((size_t)n);
      for(int d = 0; d < D; ++d) {
        for(int64_t i = 0; i < n; ++i) tmp[i] = h[i * D + d];
        FSPH_CUDA_OK(cudaMemcpyAsync(s->col[VEC_FIELD[col] + d].p, tmp.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
        FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
      }
      if(col == FSPH_COL_POS) { s->tree_dirty = true; s->search_valid = false; }
    }
    else if(scalar_field(*s, col) >= 0) {
      FSPH_CUDA_OK(cudaMemcpyAsync(s->col[scalar_field(*s, col)].p, host, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
      if(col == FSPH_COL_H || col == FSPH_COL_MASS) { s->tree_dirty = true; s->search_valid = false; }
    }
    else if(col == FSPH_COL_ID) FSPH_CUDA_OK(cudaMemcpyAsync(s->col_id.p, host, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
    else if(col == FSPH_COL_TYPE) FSPH_CUDA_OK(cudaMemcpyAsync(s->col_type.p, host, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
    else throw std::runtime_error("column not writable");
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
  });
}

int32_t fsph_get_column(fsph_system * s, int32_t col, void * host) {
  return guarded(s, [&]() {
    dist_drop_ghosts(*s);
    const int64_t n = s->n;
    const int D = s->dim;
    if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
      double * h = static_cast<double *>(host);
      std::vector<double> tmp((size_t)n);
      for(int d = 0; d < D; ++d) {
        FSPH_CUDA_OK(cudaMemcpyAsync(tmp.data(), s->col[VEC_FIELD[col] + d].p, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
        FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
        for(int64_t i = 0; i < n; ++i) h[i * D + d] = tmp[i];
      }
      return;
    }
    const void * src = nullptr;
    size_t bytes = (size_t)n * 8;
    if(scalar_field(*s, col) >= 0) src = s->col[scalar_field(*s, col)].p;
    else if(col == FSPH_COL_ID) src = s->col_id.p;
    else if(col == FSPH_COL_KEY) src = s->col_key.p;
    else if(col == FSPH_COL_TYPE) { src = s->col_type.p; bytes = (size_t)n * 4; }
    else if(col == FSPH_COL_NNEIGH) { src = s->col_nneigh.p; bytes = (size_t)n * 4; }
    else throw std::runtime_error("bad column");
    FSPH_CUDA_OK(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
  });
}

int32_t fsph_set_component(fsph_system * s, int32_t col, int32_t comp, const void * host) {
  return guarded(s, [&]() {
    dist_drop_ghosts(*s);
    s->pair_pi_valid = false;
    void * dst = nullptr;
    size_t esz = 8;
    if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
      if(comp < 0 || comp >= s->dim) throw std::runtime_error("bad component");
      dst = s->col[VEC_FIELD[col] + comp].p;
    }
    else if(scalar_field(*s, col) >= 0) dst = s->col[scalar_field(*s, col)].p;
    else if(col == FSPH_COL_ID) dst = s->col_id.p;
    else if(col == FSPH_COL_TYPE) { dst = s->col_type.p; esz = 4; }
    else throw std::runtime_error("column not writable");
    FSPH_CUDA_OK(cudaMemcpyAsync(dst, host, (size_t)s->n * esz, cudaMemcpyHostToDevice, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    if(col == FSPH_COL_POS || col == FSPH_COL_H || col == FSPH_COL_MASS) { s->tree_dirty = true; s->search_valid = false; }
  });
}

int32_t fsph_get_component(fsph_system * s, int32_t col, int32_t comp, void * host) {
  return guarded(s, [&]() {
    dist_drop_ghosts(*s);
    const void * src = nullptr;
    size_t esz = 8;
    if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
      if(comp < 0 || comp >= s->dim) throw std::runtime_error("bad component");
      src = s->col[VEC_FIELD[col] + comp].p;
    }
    else if(scalar_field(*s, col) >= 0) src = s->col[scalar_field(*s, col)].p;
    else if(col == FSPH_COL_ID) src = s->col_id.p;
    else if(col == FSPH_COL_KEY) src = s->col_key.p;
    else if(col == FSPH_COL_TYPE) { src = s->col_type.p; esz = 4; }
    else if(col == FSPH_COL_NNEIGH) { src = s->col_nneigh.p; esz = 4; }
    else throw std::runtime_error("bad column");
    FSPH_CUDA_OK(cudaMemcpyAsync(host, src, (size_t)s->n * esz, cudaMemcpyDeviceToHost, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
  });
}

// AoS marshalling of the reference's `body` records (include/physics/body.h:221-251): offsets[col] = offsetof(body, field).
int32_t fsph_upload_aos(fsph_system * s, const void * bodies, int64_t n, int64_t stride, const int64_t * offsets) {
  int32_t rc = fsph_resize(s, n);
  if(rc) return rc;
  return guarded(s, [&]() {
    const char * base = static_cast<const char *>(bodies);
    const int D = s->dim;
    std::vector<double> tmp((size_t)n);
    auto up = [&](void * dst, int64_t off, size_t esz) {
      if(esz == 8) {
        for(int64_t i = 0; i < n; ++i) std::memcpy(&tmp[i], base + i * stride + off, 8);
        FSPH_CUDA_OK(cudaMemcpyAsync(dst, tmp.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
      }
      else {
        int32_t * t4 = reinterpret_cast<int32_t *>(tmp.data());
        for(int64_t i = 0; i < n; ++i) std::memcpy(&t4[i], base + i * stride + off, 4);
        FSPH_CUDA_OK(cudaMemcpyAsync(dst, t4, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
      }
      FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    };
    for(int col = 0; col < FSPH_COL_COUNT_; ++col) {
      if(offsets[col] < 0) continue;
      if(col <= FSPH_COL_GACC) for(int d = 0; d < D; ++d) up(s->col[VEC_FIELD[col] + d].p, offsets[col] + 8 * d, 8);
      else if(scalar_field(*s, col) >= 0) up(s->col[scalar_field(*s, col)].p, offsets[col], 8);
      else if(col == FSPH_COL_ID) up(s->col_id.p, offsets[col], 8);
      else if(col == FSPH_COL_TYPE) up(s->col_type.p, offsets[col], 4);
    }
    invalidate(*s);
  });
}

int32_t fsph_download_aos(fsph_system * s, void * bodies, int64_t stride, const int64_t * offsets) {
  return guarded(s, [&]() {
    dist_drop_ghosts(*s);
    char * base = static_cast<char *>(bodies);
    const int D = s->dim;
    const int64_t n = s->n;
    std::vector<double> tmp((size_t)n);
    auto down = [&](const void * src, int64_t off, size_t esz) {
      FSPH_CUDA_OK(cudaMemcpyAsync(tmp.data(), src, (size_t)n * esz, cudaMemcpyDeviceToHost, s->stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
      const char * t = reinterpret_cast<const char *>(tmp.data());
      for(int64_t i = 0; i < n; ++i) std::memcpy(base + i * stride + off, t + i * esz, esz);
    };
    for(int col = 0; col < FSPH_COL_COUNT_; ++col) {
      if(offsets[col] < 0) continue;
      if(col <= FSPH_COL_GACC) for(int d = 0; d < D; ++d) down(s->col[VEC_FIELD[col] + d].p, offsets[col] + 8 * d, 8);
      else if(scalar_field(*s, col) >= 0) down(s->col[scalar_field(*s, col)].p, offsets[col], 8);
      else if(col == FSPH_COL_ID) down(s->col_id.p, offsets[col], 8);
      else if(col == FSPH_COL_KEY) down(s->col_key.p, offsets[col], 8);
      else if(col == FSPH_COL_TYPE) down(s->col_type.p, offsets[col], 4);
      else if(col == FSPH_COL_NNEIGH) down(s->col_nneigh.p, offsets[col], 4);
    }
  });
}

int32_t fsph_set_time(fsph_system * s, double dt, double totaltime, int64_t iteration) {
  if(!s) return 1;
  s->dt = dt;
  s->totaltime = totaltime;
  s->iteration = iteration;
  return 0;
}
int32_t fsph_get_time(const fsph_system * s, double * dt, double * totaltime, int64_t * iteration) {
  if(!s) return 1;
  if(dt) *dt = s->dt;
  if(totaltime) *totaltime = s->totaltime;
  if(iteration) *iteration = s->iteration;
  return 0;
}
int32_t fsph_advance_time(fsph_system * s) {
  if(!s) return 1;
  s->iteration++;
  if(s->par.adaptive_timestep) s->totaltime = s->totaltime_next;
  else s->totaltime = s->totaltime + s->dt;
  return 0;
}

int32_t fsph_set_time_state(fsph_system * s, const double * st, int64_t iteration) {
  if(!s || !st) return 1;
  s->dt = st[0]; s->dt_saved = st[1]; s->totaltime = st[2]; s->totaltime_next = st[3];
  // System::t_out is ordered screen, scalar, h5data (the order set_adaptive_timestep visits them)
  s->t_out[2] = st[4]; s->t_out[0] = st[5]; s->t_out[1] = st[6];
  s->iteration = iteration;
  return 0;
}
int32_t fsph_get_time_state(const fsph_system * s, double * st, int64_t * iteration) {
  if(!s || !st) return 1;
  st[0] = s->dt; st[1] = s->dt_saved; st[2] = s->totaltime; st[3] = s->totaltime_next;
  st[4] = s->t_out[2]; st[5] = s->t_out[0]; st[6] = s->t_out[1];
  if(iteration) *iteration = s->iteration;
  return 0;
}

int32_t fsph_update_iteration(fsph_system * s) {
  return guarded(s, [&]() { do_update_iteration(*s); check_status(*s); });
}
int32_t fsph_reset_ghosts(fsph_system * s) {
  return guarded(s, [&]() { do_reset_ghosts(*s); });
}
int32_t fsph_sweep(fsph_system * s, int32_t op) {
  return guarded(s, [&]() { do_sweep(*s, op); check_status(*s); });
}
int32_t fsph_apply(fsph_system * s, int32_t op) {
  return guarded(s, [&]() {
    if(op < 0 || op > FSPH_APPLY_ADD_DRAG_DUDT) throw std::runtime_error("unknown apply op");
    apply(*s, op);
    check_status(*s);
  });
}
int32_t fsph_reduce(fsph_system * s, int32_t op, double * out, int32_t nout) {
  return guarded(s, [&]() { reduce(*s, op, out, nout); });
}
int32_t fsph_gravitation_fmm(fsph_system * s) {
  return guarded(s, [&]() { gravitation_fmm(*s); check_status(*s); });
}
int32_t fsph_get_range(fsph_system * s, double * lo, double * hi) {
  // body_system::getRange (bodies_system.h:131-138) is mpi_compute_range of the bodies as they are NOW -- not the range of
  // the last update_iteration (mpisph/test/bs.cc:29-35 asks for it right after read_bodies).  While positions and h are
  // those of the last update_iteration the stored range is that very result.
  return guarded(s, [&]() {
    if(!(s->tree_valid && !s->tree_dirty) && s->n_owned_or_n() > 0) {
      double keep_lo[3], keep_hi[3];
      for(int d = 0; d < 3; ++d) { keep_lo[d] = s->lo[d]; keep_hi[d] = s->hi[d]; }
      compute_range(*s);
      if(s->comm) {
        // MPI_Allreduce MIN / MAX of the ranks' ranges (bodies_system.h:255-262)
        const int P = s->comm->nranks;
        double mine[6] = {s->lo[0], s->lo[1], s->lo[2], s->hi[0], s->hi[1], s->hi[2]};
        std::vector<double> all((size_t)6 * P);
        dist_all_gather_host(*s, mine, all.data(), sizeof(mine));
        for(int q = 0; q < P; ++q)
          for(int d = 0; d < s->dim; ++d) {
            s->lo[d] = std::min(s->lo[d], all[(size_t)6 * q + d]);
            s->hi[d] = std::max(s->hi[d], all[(size_t)6 * q + 3 + d]);
          }
      }
      for(int d = 0; d < s->dim; ++d) { lo[d] = s->lo[d]; hi[d] = s->hi[d]; }
      // the keys of a still valid tree belong to the range they were made with
      if(s->tree_valid)
        for(int d = 0; d < 3; ++d) { s->lo[d] = keep_lo[d]; s->hi[d] = keep_hi[d]; }
      return;
    }
    for(int d = 0; d < s->dim; ++d) { lo[d] = s->lo[d]; hi[d] = s->hi[d]; }
  });
}
double fsph_get_smoothinglength(fsph_system * s) {
  double r = 0;
  guarded(s, [&]() {
    double o[3] = {max_h(*s), 0, 0};
    if(s->comm) global_combine(*s, 1 /* RED_MAX_H */, o);  // MPI_Allreduce MAX, bodies_system.h:122
    r = o[0];
  });
  return r;
}

static int32_t step_impl(fsph_system * s, int32_t initial);

int32_t fsph_step(fsph_system * s, int32_t initial) {
  if(s) s->stream_out.clear();
  return step_impl(s, initial);
}

// device address and element size of one writable component (fsph_set_component)
static void * column_dest(System & s, int col, int comp, size_t & esz) {
  esz = 8;
  if(col >= FSPH_COL_POS && col <= FSPH_COL_GACC) {
    if(comp < 0 || comp >= s.dim) throw std::runtime_error("bad component");
    return s.col[VEC_FIELD[col] + comp].p;
  }
  if(scalar_field(s, col) >= 0) return s.col[scalar_field(s, col)].p;
  if(col == FSPH_COL_ID) return s.col_id.p;
  if(col == FSPH_COL_TYPE) { esz = 4; return s.col_type.p; }
  throw std::runtime_error("column not writable");
}

struct PipelinedScope {
  fsph_system * s;
  explicit PipelinedScope(fsph_system * sys) : s(sys) { s->in_pipelined_call = true; }
  ~PipelinedScope() { s->in_pipelined_call = false; }
};

// One driver iteration with its host traffic pipelined against the neighbouring iterations (include/flecsph_b200.h):
//   up_stream    host-to-device copies of this step's inputs; each waits for the end of the previous step's kernels and for
//                the previous step's device-to-host copy of the same column / into the same host buffer
//   stream       the step; waits for the inputs and for the previous step's copies (it overwrites what they read)
//   side_stream  device-to-host copies of the results, each ordered after the last kernel that writes its column
// Returns without waiting for the device-to-host copies: fsph_stream_wait, or the next pipelined step, orders after them.
int32_t fsph_step_pipelined(fsph_system * s, int32_t initial, const fsph_stream_in * in, int32_t nin, const fsph_stream_out * out,
  int32_t nout) {
  if(!s) return 1;
  PipelinedScope scope(s);
  int32_t rc = guarded(s, [&]() {
    if(nout < 0 || (nout > 0 && !out)) throw std::runtime_error("fsph_step_pipelined: bad output list");
    if(nin < 0 || (nin > 0 && !in)) throw std::runtime_error("fsph_step_pipelined: bad input list");
    if(!s->side_stream) {
      FSPH_CUDA_OK(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
      FSPH_CUDA_OK(cudaEventCreateWithFlags(&s->side_event, cudaEventDisableTiming));
    }
    if(!s->up_stream) {
      FSPH_CUDA_OK(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
      FSPH_CUDA_OK(cudaEventCreateWithFlags(&s->up_event, cudaEventDisableTiming));
    }
    constexpr int HALF = 64;   // copies in flight per step; the events of consecutive steps come from alternate halves
    if(nout > HALF) throw std::runtime_error("fsph_step_pipelined: more than 64 outputs");
    while((int)s->ev_pool.size() < 2 * HALF) {
      cudaEvent_t e;
      FSPH_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      s->ev_pool.push_back(e);
    }
    if(nin > 0) {
      // the host sees the owned bodies only (fsph_set_component drops the remote copies first; here they stay where
      // they are -- the previous step's copies may still be reading the columns -- and the inputs go to the owned range)
      const int64_t first = s->has_ghosts ? s->own_begin : 0, cnt = s->n_owned_or_n();
      s->pair_pi_valid = false;
      // the kernels of the previous step are done with the columns
      FSPH_CUDA_OK(cudaEventRecord(s->up_event, s->stream));
      FSPH_CUDA_OK(cudaStreamWaitEvent(s->up_stream, s->up_event, 0));
      for(int k = 0; k < nin; ++k) {
        size_t esz;
        void * dst = column_dest(*s, in[k].col, in[k].comp, esz);
        for(const auto & p : s->prev_out)
          if(p.host == in[k].host || (p.col == in[k].col && p.comp == in[k].comp)) FSPH_CUDA_OK(cudaStreamWaitEvent(s->up_stream, p.ev, 0));
        FSPH_CUDA_OK(cudaMemcpyAsync(static_cast<char *>(dst) + (size_t)first * esz, in[k].host, (size_t)cnt * esz, cudaMemcpyHostToDevice, s->up_stream));
        if(in[k].col == FSPH_COL_POS || in[k].col == FSPH_COL_H || in[k].col == FSPH_COL_MASS) { s->tree_dirty = true; s->search_valid = false; }
      }
      FSPH_CUDA_OK(cudaEventRecord(s->up_event, s->up_stream));
      FSPH_CUDA_OK(cudaStreamWaitEvent(s->stream, s->up_event, 0));
    }
    // this step overwrites columns the previous step's copies still read: the copies are in order on their stream, so the
    // event of the last one covers them all
    if(!s->prev_out.empty()) FSPH_CUDA_OK(cudaStreamWaitEvent(s->stream, s->prev_out.back().ev, 0));
    const bool upper = !s->prev_out.empty() && s->prev_out.front().ev == s->ev_pool[0];
    s->stream_out.clear();
    for(int k = 0; k < nout; ++k) s->stream_out.push_back({out[k].col, out[k].comp, out[k].host, false, s->ev_pool[(upper ? HALF : 0) + k]});
  });
  if(rc) return rc;
  rc = step_impl(s, initial);
  const int32_t rc2 = guarded(s, [&]() {
    if(rc == 0) {
      stream_flush(*s, nullptr, 0);
      s->prev_out = s->stream_out;
      s->stream_out.clear();
      return;
    }
    const cudaError_t e = cudaStreamSynchronize(s->side_stream);
    s->stream_out.clear();
    s->prev_out.clear();
    FSPH_CUDA_OK(e);
  });
  return rc ? rc : rc2;
}

int32_t fsph_stream_wait(fsph_system * s) {
  if(!s) return 1;
  PipelinedScope scope(s);
  return guarded(s, [&]() {
    cudaError_t e = cudaSuccess;
    if(s->side_stream) e = cudaStreamSynchronize(s->side_stream);
    s->prev_out.clear();
    FSPH_CUDA_OK(e);
  });
}

int32_t fsph_step_streamed(fsph_system * s, int32_t initial, const fsph_stream_out * out, int32_t nout) {
  const int32_t rc = fsph_step_pipelined(s, initial, nullptr, 0, out, nout);
  const int32_t rc2 = fsph_stream_wait(s);
  return rc ? rc : rc2;
}

static int32_t step_impl(fsph_system * s, int32_t initial) {
  return guarded(s, [&]() {
    s->timing = true;
    s->phases.clear();
    Phase tot;
    tot.name = "total";
    cudaEventCreate(&tot.e0);
    cudaEventCreate(&tot.e1);
    cudaEventRecord(tot.e0, s->stream);
    try {
      s->in_step = true;
      do_step(*s, initial);
      s->in_step = false;
    }
    catch(...) {
      s->in_step = false;
      cudaEventDestroy(tot.e0);
      cudaEventDestroy(tot.e1);
      for(auto & p : s->phases) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
      s->phases.clear();
      s->timing = false;
      throw;
    }
    cudaEventRecord(tot.e1, s->stream);
    s->phases.push_back(tot);
    finish_phases(*s);
    s->timing = false;
  });
}

int64_t fsph_tree_num_nodes(fsph_system * s) { return s && s->tree_valid ? s->tree.nnodes : 0; }

int32_t fsph_tree_get_nodes(fsph_system * s, uint64_t * key, double * mass, double * com, double * radius, double * lap,
  double * bmin, double * bmax, int64_t * sub_entities) {
  return guarded(s, [&]() {
    if(!s->tree_valid) throw std::runtime_error("no tree");
    const int nn = s->tree.nnodes, D = s->dim, B = 63 / D;
    const int64_t n = s->n;
    std::vector<uint64_t> pkey((size_t)n);
    std::vector<int32_t> ns(nn), ndp(nn), sb(nn);
    std::vector<double> tmp(nn);
    auto get = [&](void * dst, const void * src, size_t bytes) {
      FSPH_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s->stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    };
    get(pkey.data(), s->col_key.p, (size_t)n * 8);
    get(ns.data(), s->tree.node_start, (size_t)nn * 4);
    get(ndp.data(), s->tree.node_depth, (size_t)nn * 4);
    get(sb.data(), s->tree.sub, (size_t)nn * 4);
    get(mass, s->tree.mass, (size_t)nn * 8);
    get(radius, s->tree.radius, (size_t)nn * 8);
    get(lap, s->tree.lap, (size_t)nn * 8);
    for(int i = 0; i < nn; ++i) {
      key[i] = pkey[ns[i]] >> (D * (B - ndp[i]));
      sub_entities[i] = sb[i];
    }
    for(int d = 0; d < D; ++d) {
      get(tmp.data(), s->tree.com[d], (size_t)nn * 8);
      for(int i = 0; i < nn; ++i) com[(size_t)i * D + d] = tmp[i];
      get(tmp.data(), s->tree.bmin[d], (size_t)nn * 8);
      for(int i = 0; i < nn; ++i) bmin[(size_t)i * D + d] = tmp[i];
      get(tmp.data(), s->tree.bmax[d], (size_t)nn * 8);
      for(int i = 0; i < nn; ++i) bmax[(size_t)i * D + d] = tmp[i];
    }
  });
}

int32_t fsph_tree_root(fsph_system * s, double * com, double * mass) {
  return guarded(s, [&]() {
    if(!s->tree_valid) throw std::runtime_error("no tree");
    // node 0 of the level-ordered node table is the root
    for(int d = 0; d < s->dim; ++d)
      FSPH_CUDA_OK(cudaMemcpyAsync(com + d, s->tree.com[d], 8, cudaMemcpyDeviceToHost, s->stream));
    if(mass) FSPH_CUDA_OK(cudaMemcpyAsync(mass, s->tree.mass, 8, cudaMemcpyDeviceToHost, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
  });
}

int32_t fsph_tree_get_leaf_depth(fsph_system * s, int32_t * depth) {
  return guarded(s, [&]() {
    if(!s->tree_valid) throw std::runtime_error("no tree");
    const int nc = s->tree.ncells;
    std::vector<uint64_t> desc(nc);
    FSPH_CUDA_OK(cudaMemcpyAsync(desc.data(), s->tree.cell_desc, (size_t)nc * 8, cudaMemcpyDeviceToHost, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    for(int c = 0; c < nc; ++c)
      if(desc[c] & 1ull) depth[(uint32_t)desc[c] >> 1] = (int32_t)(desc[c] >> 32);
  });
}

int32_t fsph_neighbor_lists(fsph_system * s, int64_t * offsets, int64_t * ids) {
  return guarded(s, [&]() {
    if(!s->search_valid) throw std::runtime_error("no neighbour search done yet");
    std::vector<int64_t> off, idv;
    export_neighbors(*s, off, ids ? &idv : nullptr);
    std::copy(off.begin(), off.end(), offsets);
    if(ids) std::copy(idv.begin(), idv.end(), ids);
  });
}

uint64_t fsph_launch_count(void) { return fsph::g_launches.load(); }

// ---- several ranks -------------------------------------------------------------------------------------------
int32_t fsph_comm_unique_id(char * out128) {
  try { nccl_unique_id(out128); return 0; }
  catch(const std::exception & e) { g_create_error = e.what(); return 1; }
}
int32_t fsph_comm_init_nccl(fsph_system * s, const char * unique_id, int32_t rank, int32_t nranks) {
  return guarded(s, [&]() {
    if(s->comm) throw std::runtime_error("communicator already set");
    if(nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("bad rank / nranks");
    if(nranks > 1) s->comm = make_nccl_comm(unique_id, rank, nranks);
  });
}
void * fsph_local_group_create(int32_t nranks) { return nranks >= 1 ? local_group_create(nranks) : nullptr; }
void fsph_local_group_destroy(void * g) { local_group_destroy(static_cast<LocalGroup *>(g)); }
int32_t fsph_comm_init_local(fsph_system * s, void * group, int32_t rank) {
  return guarded(s, [&]() {
    if(s->comm) throw std::runtime_error("communicator already set");
    s->comm = make_local_comm(static_cast<LocalGroup *>(group), rank);
  });
}
int32_t fsph_comm_rank(const fsph_system * s) { return s && s->comm ? s->comm->rank : 0; }
int32_t fsph_comm_size(const fsph_system * s) { return s && s->comm ? s->comm->nranks : 1; }
int64_t fsph_nghosts(const fsph_system * s) { return s && s->has_ghosts ? s->n - s->n_owned : 0; }

// host-side planning of the decomposition (no device needed; covered by the world_size-2 CPU tests)
void fsph_dd_choose_splitters(int32_t nranks, int32_t nsamples, const uint64_t * sample_keys, const int64_t * counts, uint64_t * splitters) {
  dd_choose_splitters(nranks, nsamples, sample_keys, counts, splitters);
}
uint64_t fsph_dd_align_splitter(int32_t dim, const uint64_t * below, int32_t nbelow, const uint64_t * above, int32_t nabove, uint64_t splitter) {
  std::vector<uint64_t> b(below, below + nbelow), a(above, above + nabove);
  return dd_align_splitter(dim, b, a, splitter);
}

int32_t fsph_phase_times(fsph_system * s, double * ms, int32_t cap, const char ** names) {
  if(!s) return 0;
  int k = (int)std::min<size_t>(s->phase_ms.size(), (size_t)cap);
  for(int i = 0; i < k; ++i) ms[i] = s->phase_ms[i];
  if(names) *names = s->phase_names.c_str();
  return k;
}

}  // extern "C"
