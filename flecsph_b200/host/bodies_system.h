// bodies_system.h -- `body_system<T,D>` of the reference (mpisph/bodies_system.h:38-379) over the C ABI of
// include/flecsph_b200.h.  Put this directory FIRST on the include path: the unmodified drivers
// (app/drivers/hydro/main_driver.cc:42, app/drivers/newtonian/main_driver.cc) and include/physics/analysis.h:28
// include "bodies_system.h" by name, so this header shadows mpisph/bodies_system.h and every
// `bs.update_iteration()`, `bs.apply_in_smoothinglength(physics::compute_acceleration)`, ... of the driver loop runs
// on the GPU (libflecsph_b200.so); nothing of the reference's tree / traversal / physics bodies executes on the host.
//
// Class name, member names, template signatures, argument meaning and error behaviour follow the reference:
//   * functors are arbitrary host callables in the reference; here they are dispatched BY FUNCTION IDENTITY to device
//     ops (the set the in-scope drivers pass is closed: SURVEY.md 8b).  An unknown per-particle functor is a hard
//     error (no CPU fallback) reported like mpi_assert (mpisph/utils.h:34-38): message on stderr + MPI_Abort(.,1).
//   * get_all() functors that only *read* the vector at I/O cadence (diagnostic::compute_*_stats, user analysis) and
//     getLocalbodies() are the documented host<->device synchronisation points: the AoS mirror is refreshed from the
//     device columns first (and, for getLocalbodies(), pushed back before the next device call).
//   * the `physics::` time globals stay authoritative on the host (advance_time / termination_criteria run there); they
//     are mirrored into the system before each device call and read back after set_adaptive_timestep.
#pragma once

#include "fmm.h"
#include "io.h"
#include "params.h"
#include "utils.h"

#include <chrono>
#include <thread>
#include <vector>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "flecsph_b200.h"

using namespace mpi_utils; // as mpisph/bodies_system.h:27 (diagnostic.h relies on it)

// ---- the get_all callables of analysis.h / diagnostic.h (declared here because this header precedes them in the
// drivers' include order; `weak` so that a program that never includes one of them still links) -------------------
#define FSPH_WEAK __attribute__((weak))
// physics::, integration::, eos:: come from default_physics.h, which "io.h" includes (mpisph/io.h:35 and flatio/io.h)
namespace analysis {
extern point_t linear_momentum, total_ang_mom;
extern double total_mass, total_energy, total_kinetic_energy, total_internal_energy, total_gravitational_energy;
FSPH_WEAK void compute_total_momentum(std::vector<body> &);
FSPH_WEAK void compute_total_mass(std::vector<body> &);
FSPH_WEAK void compute_total_energy(std::vector<body> &);
FSPH_WEAK void compute_total_kinetic_energy(std::vector<body> &);
FSPH_WEAK void compute_total_internal_energy(std::vector<body> &);
FSPH_WEAK void compute_total_gravitational_energy(std::vector<body> &);
FSPH_WEAK void compute_total_ang_mom(std::vector<body> &);
} // namespace analysis

namespace diagnostic {
extern uint64_t N_min, N_max, N_average;
extern double h_min, h_max, h_average, average_dist_in_h, V_min, V_max, V_average;
FSPH_WEAK void compute_neighbors_stats(std::vector<body> &, int64_t);
FSPH_WEAK void compute_smoothinglength_stats(std::vector<body> &, int64_t);
FSPH_WEAK void compute_velocity_stats(std::vector<body> &, int64_t);
} // namespace diagnostic

namespace fsph_host {
[[noreturn]] inline void
fail(const char * what, const char * detail = "") {
  std::fprintf(stderr, "flecsph_b200: %s %s\n", what, detail);
  MPI_Abort(MPI_COMM_WORLD, 1);
  std::abort();
}

// what `bs.tree()->root_node()->coordinates()` needs (analysis.h:396); mass() as in mpisph/test/tree3d.cc:85
struct root_node_view {
  point_t com{};
  double m = 0;
  point_t coordinates() const {
    return com;
  }
  double mass() const {
    return m;
  }
};
} // namespace fsph_host

namespace fsph_host {
// index loop over the AoS mirror on up to 16 host threads (the conversions between the 368-byte records and the device
// columns are memory-bound: ~35 passes over the bodies per synchronisation)
template<typename F>
inline void parallel_for(int64_t n, F f) {
  const unsigned hw = std::thread::hardware_concurrency();
  const unsigned nt = std::min<unsigned>(hw ? hw : 1u, 16u);
  if(n < 200000 || nt <= 1) {
    f((int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (n + nt - 1) / nt;
  for(unsigned t = 0; t < nt; ++t) {
    const int64_t a = (int64_t)t * chunk, b = std::min<int64_t>(n, a + chunk);
    if(a < b) th.emplace_back([=]() { f(a, b); });
  }
  for(auto & x : th) x.join();
}
}   // namespace fsph_host

template<typename T, size_t D>
class body_system
{
  using point_t = flecsi::space_vector_u<T, D>;
  using sweep_fn = void (*)(body &, std::vector<body *> &);
  using apply_fn = void (*)(body &);
  using get_fn = void (*)(std::vector<body> &);
  using get_n_fn = void (*)(std::vector<body> &, int64_t);

public:
  // the reference hands out tree_topology_t*; the drivers only use tree()->root_node()->coordinates()
  struct tree_view {
    body_system * bs;
    fsph_host::root_node_view node;
    fsph_host::root_node_view * root_node() {
      bs->sync_device();
      double com[3] = {0, 0, 0};
      bs->ck(fsph_tree_root(bs->sys_, com, &node.m));
      for(size_t d = 0; d < D; ++d) node.com[d] = com[d];
      return &node;
    }
  };

  body_system() : totalnbodies_(0L), localnbodies_(0L), macangle_(0.0) {
    static_assert(std::is_same<T, double>::value, "the device path is fp64");
    if(param::sph_variable_h) {
      log_one(warn) << "Variable smoothing length ENABLE" << std::endl;
    }
    fsph_params p;
    fsph_default_params(&p, (int32_t)D);
    using namespace param;
    p.sph_kernel = (int32_t)sph_kernel;
    p.eos_type = (int32_t)eos_type;
    p.sph_viscosity = (int32_t)sph_viscosity;
    p.sph_eta = sph_eta;
    p.sph_sinc_index = sph_sinc_index;
    p.poly_gamma = poly_gamma;
    p.sph_viscosity_alpha = sph_viscosity_alpha;
    p.sph_viscosity_beta = sph_viscosity_beta;
    p.sph_viscosity_epsilon = sph_viscosity_epsilon;
    p.timestep_cfl_factor = timestep_cfl_factor;
    p.initial_dt = initial_dt;
    p.out_screen_dt = out_screen_dt;
    p.out_scalar_dt = out_scalar_dt;
    p.out_h5data_dt = out_h5data_dt;
    p.fmm_macangle = fmm_macangle;
    p.gravitational_constant = gravitational_constant;
    p.box_length = box_length;
    p.box_width = box_width;
    p.box_height = box_height;
    p.adaptive_timestep = adaptive_timestep;
    p.sph_variable_h = sph_variable_h;
    p.sph_update_uniform_h = sph_update_uniform_h;
    p.evolve_internal_energy = evolve_internal_energy;
    p.thermokinetic_formulation = thermokinetic_formulation;
    p.enable_fmm = enable_fmm;
    p.relaxation_steps = relaxation_steps;
    p.pressure_updates_number = pressure_updates_number;
    p.periodic_boundary_x = periodic_boundary_x;
    p.periodic_boundary_y = periodic_boundary_y;
    p.periodic_boundary_z = periodic_boundary_z;
    p.sph_viscosity_alpha_max = sph_viscosity_alpha_max;
    p.sph_viscosity_l = sph_viscosity_l;
    p.sph_viscosity_delta = sph_viscosity_delta;
    p.poly_gamma2 = poly_gamma2;
    p.ppt_density_thr = ppt_density_thr;
    p.relaxation_beta = relaxation_beta;
    p.relaxation_gamma = relaxation_gamma;
    p.relaxation_repulsion_radius = relaxation_repulsion_radius;
    p.relaxation_repulsion_gamma = relaxation_repulsion_gamma;
    // external forces (eforce.h): the keyword the driver hands to external_force::select and the parameters its terms read
    std::snprintf(p.external_force_type, sizeof(p.external_force_type), "%s", external_force_type);
    std::snprintf(p.density_profile, sizeof(p.density_profile), "%s", density_profile);
    if(std::strlen(external_force_type) >= sizeof(p.external_force_type))
      fsph_host::fail("external_force_type too long for the device path:", external_force_type);
    p.sphere_radius = sphere_radius;
    p.extforce_wall_powerindex = extforce_wall_powerindex;
    p.extforce_wall_steepness = extforce_wall_steepness;
    p.gravity_acceleration_constant = gravity_acceleration_constant;
    p.zero_potential_poison_value = zero_potential_poison_value;
    p.mesa_rim_width = mesa_rim_width;
    p.rho_initial = rho_initial;
    p.pressure_initial = pressure_initial;
    p.orbital_separation = orbital_separation;
    p.mass_neutron_star = mass_neutron_star;
    p.mass_white_dwarf = mass_white_dwarf;
    p.airfoil_size = airfoil_size;
    p.airfoil_thickness = airfoil_thickness;
    p.airfoil_camber = airfoil_camber;
    p.airfoil_anchor_x = airfoil_anchor_x;
    p.airfoil_anchor_y = airfoil_anchor_y;
    p.airfoil_attack_angle = airfoil_attack_angle;
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const char * dev = std::getenv("FSPH_DEVICE");
    sys_ = fsph_create(&p, dev ? std::atoi(dev) : rank);
    if(!sys_)
      fsph_host::fail("fsph_create failed:", fsph_last_error(nullptr));
    // Several MPI ranks (one per GPU): the ranks' systems join one NCCL communicator -- MPI only carries the 128-byte
    // id, the particle traffic of update_iteration / reset_ghosts goes over NCCL inside the library (fsph_dist.cu).
    int size = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    if(size > 1) {
      char id[128] = {0};
      if(rank == 0 && fsph_comm_unique_id(id) != 0)
        fsph_host::fail("fsph_comm_unique_id failed:", fsph_last_error(nullptr));
      MPI_Bcast(id, 128, MPI_CHAR, 0, MPI_COMM_WORLD);
      ck(fsph_comm_init_nccl(sys_, id, rank, size));
    }
    tree_.bs = this;
  }

  ~body_system() {
    fsph_destroy(sys_);
  }

  void setMacangle(double macangle) {
    macangle_ = macangle;
  }

  void read_bodies(const char * input_prefix, const char * output_prefix, const int startiteration) {
    io::inputDataHDF5(bodies_, input_prefix, output_prefix, totalnbodies_, localnbodies_, startiteration);
    host_newer_ = true;
    device_newer_ = false;
  }

  void write_bodies(const char * output_prefix, int iter, double totaltime) {
    sync_host();
    io::outputDataHDF5(bodies_, output_prefix, iter, totaltime);
  }

  double getSmoothinglength() {
    sync_device();
    return fsph_get_smoothinglength(sys_);
  }

  std::array<point_t, 2> & getRange() {
    sync_device();
    double lo[3], hi[3];
    ck(fsph_get_range(sys_, lo, hi));
    for(size_t d = 0; d < D; ++d) {
      range_[0][d] = lo[d];
      range_[1][d] = hi[d];
    }
    return range_;
  }

  void update_iteration() {
    if(getenv("FSPH_FACADE_TRACE")) {
      // wall time from one update_iteration to the next = one driver iteration on the facade
      static double last = -1.0;
      const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
      if(last >= 0.0) fprintf(stderr, "[facade] iteration wall %.1f ms (%lld bodies)\n", 1e3 * (now - last), (long long)fsph_nbodies(sys_));
      last = now;
    }
    sync_device();
    ck(fsph_update_iteration(sys_));
    localnbodies_ = fsph_nbodies(sys_);
    device_newer_ = true;
  }

  void reset_ghosts() {
    sync_device();
    ck(fsph_reset_ghosts(sys_));
  }

  void gravitation_fmm() {
    assert(gdimension == 3);
    sync_device();
    ck(fsph_gravitation_fmm(sys_));
    device_newer_ = true;
  }

  template<typename EF, typename... ARGS>
  void apply_in_smoothinglength(EF && ef, ARGS &&... args) {
    if constexpr(sizeof...(ARGS) == 1) {
      // The neighbour census of the search micro-benchmark (app/drivers/tree/main_driver.cc:103-114): a capturing
      // lambda (entity &, std::vector<entity *> &, size_t & total) that adds up the list lengths and asserts that every
      // body is in its own list.  The search runs on the device (FSPH_SWEEP_NEIGHBORS_ONLY); the functor body is not
      // executed on the host: the total comes from the device reduction of the neighbour counts, the self-membership
      // from its minimum (a body always satisfies the pair predicate with itself).
      sync_device();
      ck(fsph_sweep(sys_, FSPH_SWEEP_NEIGHBORS_ONLY));
      double o[3] = {0, 0, 0};
      ck(fsph_reduce(sys_, FSPH_GET_NEIGHBORS_STATS, o, 3));
      if(!(o[0] >= 1)) fsph_host::fail("apply_in_smoothinglength: a body is missing from its own neighbour list");
      ((args += (std::remove_reference_t<ARGS>)o[2]), ...);
      return;
    }
    else {
      apply_in_smoothinglength_known(std::forward<EF>(ef));
    }
  }

  template<typename EF>
  void apply_in_smoothinglength_known(EF && ef) {
    int op = -1;
    if constexpr(std::is_convertible<EF, sweep_fn>::value) {
      const sweep_fn f = ef;
      if(f == &physics::compute_density_pressure_soundspeed) op = FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED;
      else if(f == &physics::compute_acceleration) op = FSPH_SWEEP_ACCELERATION;
      else if(f == &physics::compute_dudt) op = FSPH_SWEEP_DUDT;
      else if(f == &physics::compute_dedt) op = FSPH_SWEEP_DEDT;
      else if(f == &viscosity::compute_alpha) op = FSPH_SWEEP_COMPUTE_ALPHA;
      else if(f == &physics::add_short_range_repulsion) op = FSPH_SWEEP_SHORT_RANGE_REPULSION;
    }
    if(op < 0)
      fsph_host::fail("apply_in_smoothinglength: this functor has no device implementation (no CPU fallback)");
    sync_device();
    ck(fsph_sweep(sys_, op));
    device_newer_ = true;
  }

  template<typename EF, typename... ARGS>
  void apply_all(EF && ef, ARGS &&...) {
    static_assert(sizeof...(ARGS) == 0, "device element-wise passes take no extra arguments");
    int op = -1;
    if constexpr(std::is_convertible<EF, apply_fn>::value) {
      const apply_fn f = ef;
      if(f == eos::init) op = FSPH_APPLY_EOS_INIT;
      else if(f == &physics::set_total_energy) op = FSPH_APPLY_SET_TOTAL_ENERGY;
      else if(f == &integration::save_velocityhalf) op = FSPH_APPLY_SAVE_VELOCITYHALF;
      else if(f == &integration::leapfrog_kick_v) op = FSPH_APPLY_LEAPFROG_KICK_V;
      else if(f == &integration::leapfrog_kick_u) op = FSPH_APPLY_LEAPFROG_KICK_U;
      else if(f == &integration::leapfrog_kick_e) op = FSPH_APPLY_LEAPFROG_KICK_E;
      else if(f == &integration::leapfrog_drift) op = FSPH_APPLY_LEAPFROG_DRIFT;
      else if(f == &physics::recompute_pressure_soundspeed) op = FSPH_APPLY_RECOMPUTE_P_CS;
      else if(f == &physics::recompute_pressure_soundspeed_thermokinetic) op = FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC;
      else if(f == &physics::compute_dt) op = FSPH_APPLY_COMPUTE_DT;
      else if(f == &physics::check_nans) op = FSPH_APPLY_CHECK_NANS;
      else if(f == &physics::check_negativity) op = FSPH_APPLY_CHECK_NEGATIVITY;
      else if(f == &viscosity::initialize_alpha) op = FSPH_APPLY_INITIALIZE_ALPHA;
      else if(f == &physics::add_drag_acceleration) op = FSPH_APPLY_ADD_DRAG_ACCELERATION;
      else if(f == &physics::add_drag_dedt) op = FSPH_APPLY_ADD_DRAG_DEDT;
      else if(f == &physics::add_drag_dudt) op = FSPH_APPLY_ADD_DRAG_DUDT;
    }
    if(op < 0)
      fsph_host::fail("apply_all: this functor has no device implementation (no CPU fallback)");
    sync_device();
    ck(fsph_apply(sys_, op));
    device_newer_ = true;
  }

  template<typename EF, typename... ARGS>
  void get_all(EF && ef, ARGS &&... args) {
    if constexpr(sizeof...(ARGS) == 0 && std::is_convertible<EF, get_fn>::value) {
      const get_fn f = ef;
      double o[3] = {0, 0, 0};
      if(f == &physics::set_adaptive_timestep) {
        sync_device();
        ck(fsph_reduce(sys_, FSPH_GET_SET_ADAPTIVE_TIMESTEP, o, 1));
        pull_time();
        return;
      }
      if(f == &physics::compute_smoothinglength) {
        sync_device();
        ck(fsph_reduce(sys_, FSPH_GET_COMPUTE_SMOOTHINGLENGTH, nullptr, 0));
        device_newer_ = true;
        return;
      }
      if(f == &analysis::compute_total_momentum) {
        reduce_to(FSPH_GET_TOTAL_MOMENTUM, o);
        for(size_t d = 0; d < D; ++d) analysis::linear_momentum[d] = o[d];
        return;
      }
      if(f == &analysis::compute_total_ang_mom) {
        reduce_to(FSPH_GET_TOTAL_ANG_MOM, o);
        analysis::total_ang_mom = {0};
        if(D == 2) analysis::total_ang_mom[0] = o[0];
        if(D == 3) for(size_t d = 0; d < D; ++d) analysis::total_ang_mom[d] = o[d];
        return;
      }
      if(f == &analysis::compute_total_mass) { reduce_to(FSPH_GET_TOTAL_MASS, o); analysis::total_mass = o[0]; return; }
      if(f == &analysis::compute_total_energy) { reduce_to(FSPH_GET_TOTAL_ENERGY, o); analysis::total_energy = o[0]; return; }
      if(f == &analysis::compute_total_kinetic_energy) { reduce_to(FSPH_GET_TOTAL_KINETIC_ENERGY, o); analysis::total_kinetic_energy = o[0]; return; }
      if(f == &analysis::compute_total_internal_energy) { reduce_to(FSPH_GET_TOTAL_INTERNAL_ENERGY, o); analysis::total_internal_energy = o[0]; return; }
      if(f == &analysis::compute_total_gravitational_energy) {
        reduce_to(FSPH_GET_TOTAL_GRAVITATIONAL_ENERGY, o);
        analysis::total_gravitational_energy = o[0];
        return;
      }
    }
    if constexpr(sizeof...(ARGS) == 1 && std::is_convertible<EF, get_n_fn>::value) {
      const get_n_fn f = ef;
      if(f == &physics::compute_average_smoothinglength) {
        sync_device();
        ck(fsph_reduce(sys_, FSPH_GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH, nullptr, 0));
        device_newer_ = true;
        return;
      }
      // diagnostic.h:46-107: {min, max, sum} come from one device reduction each
      const int64_t total = (args, ...);
      double o[3] = {0, 0, 0};
      if(f == &diagnostic::compute_neighbors_stats) {
        reduce_to(FSPH_GET_NEIGHBORS_STATS, o);
        diagnostic::N_min = (uint64_t)o[0];
        diagnostic::N_max = std::max(diagnostic::N_max, (uint64_t)o[1]);
        diagnostic::average_dist_in_h /= total;
        diagnostic::N_average = (uint64_t)o[2] / total;
        return;
      }
      if(f == &diagnostic::compute_smoothinglength_stats) {
        reduce_to(FSPH_GET_SMOOTHINGLENGTH_STATS, o);
        diagnostic::h_min = o[0];
        diagnostic::h_max = o[1];
        diagnostic::h_average = o[2] / total;
        return;
      }
      if(f == &diagnostic::compute_velocity_stats) {
        reduce_to(FSPH_GET_VELOCITY_STATS, o);
        diagnostic::V_min = o[0];
        diagnostic::V_max = o[1];
        diagnostic::V_average = o[2] / total;
        return;
      }
    }
    // host analysis at I/O cadence (diagnostic.h:113-123 and user code): same sync point as getLocalbodies()
    sync_host();
    ef(bodies_, std::forward<ARGS>(args)...);
    host_newer_ = true;
  }

  std::vector<body> & getLocalbodies() {
    sync_host();
    host_newer_ = true; // the caller may write through the reference
    return bodies_;
  }

  size_t nbodies() {
    return device_newer_ ? (size_t)fsph_nbodies(sys_) : bodies_.size();
  }
  int64_t getNLocalBodies() {
    return localnbodies_;
  }
  int64_t getNBodies() {
    return totalnbodies_;
  }
  tree_view * tree() {
    return &tree_;
  }
  fsph_system * handle() {
    return sys_;
  }

private:
  friend struct tree_view;

  void ck(int32_t rc) {
    if(rc)
      fsph_host::fail("device call failed:", fsph_last_error(sys_));
  }

  void reduce_to(int op, double * o) {
    sync_device();
    ck(fsph_reduce(sys_, op, o, 3));
  }

  void push_time() {
    const double st[7] = {physics::dt, physics::dt_saved, physics::totaltime, physics::totaltime_next,
      physics::t_h5data_output, physics::t_screen_output, physics::t_scalar_output};
    fsph_set_time_state(sys_, st, physics::iteration);
  }
  void pull_time() {
    double st[7];
    int64_t it;
    fsph_get_time_state(sys_, st, &it);
    physics::dt = st[0];
    physics::dt_saved = st[1];
    physics::totaltime_next = st[3];
    physics::t_h5data_output = st[4];
    physics::t_screen_output = st[5];
    physics::t_scalar_output = st[6];
  }

  // host AoS -> device columns (through the public getters: the fields of body_u are private)
  void sync_device() {
    push_time();
    if(!host_newer_)
      return;
    const int64_t n = (int64_t)bodies_.size();
    ck(fsph_resize(sys_, n));
    std::vector<double> v((size_t)n);
    auto vec = [&](int col, auto getter) {
      for(size_t d = 0; d < D; ++d) {
        fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) v[i] = getter(bodies_[i])[d]; });
        ck(fsph_set_component(sys_, col, (int)d, v.data()));
      }
    };
    auto sca = [&](int col, auto getter) {
      fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) v[i] = getter(bodies_[i]); });
      ck(fsph_set_component(sys_, col, 0, v.data()));
    };
    vec(FSPH_COL_POS, [](body & b) { return b.coordinates(); });
    vec(FSPH_COL_VEL, [](body & b) { return b.getVelocity(); });
    vec(FSPH_COL_VELHALF, [](body & b) { return b.getVelocityhalf(); });
    vec(FSPH_COL_ACC, [](body & b) { return b.getAcceleration(); });
    vec(FSPH_COL_GACC, [](body & b) { return b.getGAcceleration(); });
    sca(FSPH_COL_MASS, [](body & b) { return b.mass(); });
    sca(FSPH_COL_H, [](body & b) { return b.radius(); });
    sca(FSPH_COL_RHO, [](body & b) { return b.getDensity(); });
    sca(FSPH_COL_P, [](body & b) { return b.getPressure(); });
    sca(FSPH_COL_U, [](body & b) { return b.getInternalenergy(); });
    sca(FSPH_COL_E, [](body & b) { return b.getTotalenergy(); });
    sca(FSPH_COL_CS, [](body & b) { return b.getSoundspeed(); });
    sca(FSPH_COL_DUDT, [](body & b) { return b.getDudt(); });
    sca(FSPH_COL_DEDT, [](body & b) { return b.getDedt(); });
    sca(FSPH_COL_DT, [](body & b) { return b.getDt(); });
    sca(FSPH_COL_ALPHA, [](body & b) { return b.getAlpha(); });
    sca(FSPH_COL_VSIG, [](body & b) { return b.getSignalspeed(); });
    sca(FSPH_COL_GPOT, [](body & b) { return b.getGPotential(); });
    sca(FSPH_COL_ADIABATIC, [](body & b) { return b.getAdiabatic(); });
    if(param::eos_type == param::eos_wd) sca(FSPH_COL_YE, [](body & b) { return b.getElectronfraction(); });
    if(param::sph_viscosity == param::visc_cullen) {
      sca(FSPH_COL_DIVV, [](body & b) { return b.getDivergenceV(); });
      sca(FSPH_COL_DDIVVDT, [](body & b) { return b.getDdivvdt(); });
    }
    {
      std::vector<int64_t> l((size_t)n);
      fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) l[i] = (int64_t)bodies_[i].id(); });
      ck(fsph_set_component(sys_, FSPH_COL_ID, 0, l.data()));
      std::vector<int32_t> w((size_t)n);
      fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) w[i] = (int32_t)bodies_[i].getType(); });
      ck(fsph_set_component(sys_, FSPH_COL_TYPE, 0, w.data()));
    }
    host_newer_ = false;
    device_newer_ = false;
  }

  // device columns -> host AoS (storage order of the device = key order, as entities() is after psort)
  void sync_host() {
    if(host_newer_)
      sync_device();
    if(!device_newer_)
      return;
    const int64_t n = fsph_nbodies(sys_);
    bodies_.resize((size_t)n);
    std::vector<double> v((size_t)n);
    auto vec = [&](int col, auto setter) {
      std::vector<point_t> p((size_t)n);
      for(size_t d = 0; d < D; ++d) {
        ck(fsph_get_component(sys_, col, (int)d, v.data()));
        fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) p[i][d] = v[i]; });
      }
      fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) setter(bodies_[i], p[i]); });
    };
    auto sca = [&](int col, auto setter) {
      ck(fsph_get_component(sys_, col, 0, v.data()));
      fsph_host::parallel_for(n, [&](int64_t a, int64_t b) { for(int64_t i = a; i < b; ++i) setter(bodies_[i], v[i]); });
    };
    vec(FSPH_COL_POS, [](body & b, const point_t & p) { b.set_coordinates(p); });
    vec(FSPH_COL_VEL, [](body & b, const point_t & p) { b.setVelocity(p); });
    vec(FSPH_COL_VELHALF, [](body & b, const point_t & p) { b.setVelocityhalf(p); });
    vec(FSPH_COL_ACC, [](body & b, const point_t & p) { b.setAcceleration(p); });
    vec(FSPH_COL_GACC, [](body & b, const point_t & p) { b.setGAcceleration(p); });
    sca(FSPH_COL_MASS, [](body & b, double x) { b.set_mass(x); });
    sca(FSPH_COL_H, [](body & b, double x) { b.set_radius(x); });
    sca(FSPH_COL_RHO, [](body & b, double x) { b.setDensity(x); });
    sca(FSPH_COL_P, [](body & b, double x) { b.setPressure(x); });
    sca(FSPH_COL_U, [](body & b, double x) { b.setInternalenergy(x); });
    sca(FSPH_COL_E, [](body & b, double x) { b.setTotalenergy(x); });
    sca(FSPH_COL_CS, [](body & b, double x) { b.setSoundspeed(x); });
    sca(FSPH_COL_DUDT, [](body & b, double x) { b.setDudt(x); });
    sca(FSPH_COL_DEDT, [](body & b, double x) { b.setDedt(x); });
    sca(FSPH_COL_DT, [](body & b, double x) { b.setDt(x); });
    sca(FSPH_COL_ALPHA, [](body & b, double x) { b.setAlpha(x); });
    sca(FSPH_COL_VSIG, [](body & b, double x) { b.setSignalspeed(x); });
    sca(FSPH_COL_GPOT, [](body & b, double x) { b.setGPotential(x); });
    sca(FSPH_COL_ADIABATIC, [](body & b, double x) { b.setAdiabatic(x); });
    if(param::sph_viscosity == param::visc_cullen) {
      sca(FSPH_COL_DIVV, [](body & b, double x) { b.setDivergenceV(x); });
      sca(FSPH_COL_DDIVVDT, [](body & b, double x) { b.setDdivvdt(x); });
      sca(FSPH_COL_TRIGGER, [](body & b, double x) { b.setTrigger(x); });
      sca(FSPH_COL_XI, [](body & b, double x) { b.setXi(x); });
      sca(FSPH_COL_TRACESS, [](body & b, double x) { b.setTraceSS(x); });
      sca(FSPH_COL_GRADV, [](body & b, double x) { b.setGradV(x); });
    }
    {
      std::vector<int64_t> l((size_t)n);
      ck(fsph_get_component(sys_, FSPH_COL_ID, 0, l.data()));
      for(int64_t i = 0; i < n; ++i) bodies_[i].set_id((size_t)l[i]);
      ck(fsph_get_component(sys_, FSPH_COL_KEY, 0, l.data()));
      for(int64_t i = 0; i < n; ++i) bodies_[i].set_key(key_type((uint64_t)l[i]));
      std::vector<int32_t> w((size_t)n);
      ck(fsph_get_component(sys_, FSPH_COL_TYPE, 0, w.data()));
      for(int64_t i = 0; i < n; ++i) bodies_[i].setType((int)w[i]);
      ck(fsph_get_component(sys_, FSPH_COL_NNEIGH, 0, w.data()));
      for(int64_t i = 0; i < n; ++i) bodies_[i].setNeighbors((size_t)w[i]);
    }
    device_newer_ = false;
  }

  int64_t totalnbodies_; // Total number of particles
  int64_t localnbodies_; // Local number of particles
  double macangle_;
  range_t range_;
  fsph_system * sys_ = nullptr;
  std::vector<body> bodies_; // host AoS mirror (read_bodies / write_bodies / getLocalbodies / host analysis)
  bool host_newer_ = false, device_newer_ = false;
  tree_view tree_;
};
