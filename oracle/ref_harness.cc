/*
 * ref_harness.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Puts the UNMODIFIED reference headers (read where they lie under
 * /root/reference at build time; nothing is copied into this repo) behind the
 * C ABI of include/flecsph_b200.h, so that tests/ and bench.py's cpu_baseline
 * leg can run the reference's own body_system / tree_topology / physics code
 * on the same inputs as the CUDA path.  Built by oracle/Makefile into
 * oracle/_ref/libflecsph_ref_{1,2,3}d.so with -DEXT_GDIMENSION={1,2,3}.
 *
 * What is the reference's code here: everything under namespaces physics,
 * kernels, viscosity, eos, integration, fmm, analysis, flecsi::topology,
 * psort and class body_system (mpisph/bodies_system.h).  What is ours: the
 * single-rank MPI/FleCSI/Boost stand-ins in flecsph_b200/host/shim/, the in-memory
 * io:: stub below (the image has no HDF5), the column marshalling, and
 * fsph_step(), which issues the calls of app/drivers/hydro/main_driver.cc:
 * 102-277 (and newtonian/main_driver.cc:218-222) in the driver's order.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "default_physics.h"

// mpisph/io.h needs HDF5, which this image lacks: pre-define its include guard
// and provide the two entry points body_system names (mpisph/io.h:631,874).
namespace io {
// the in-memory "reader": bodies are already in the vector, report the counts
inline void inputDataHDF5(std::vector<body> & b, const char *, const char *, int64_t & total, int64_t & local, int) {
  total = local = (int64_t)b.size();
}
inline void outputDataHDF5(std::vector<body> &, const char *, int, double) {}
} // namespace io
#define _mpisph_io_h_
#include "bodies_system.h"
#include "analysis.h"

#include "flecsph_b200.h"

namespace flecsi { namespace execution {
void driver(int, char **) {}
void specialization_tlt_init(int, char **) {}
} }

static constexpr int D = gdimension;

struct fsph_system {
  fsph_params par;
  body_system<double, gdimension> * bs = nullptr;
  std::string err;
  // neighbour lists of the last sweep, indexed by storage position
  std::vector<std::vector<int64_t>> nbr;
};

static void
install_params(const fsph_params & p) {
  using namespace param;
  _sph_kernel = static_cast<sph_kernel_keyword>(p.sph_kernel);
  _eos_type = static_cast<eos_type_keyword>(p.eos_type);
  _sph_viscosity = static_cast<sph_viscosity_keyword>(p.sph_viscosity);
  _sph_eta = p.sph_eta;
  _sph_sinc_index = p.sph_sinc_index;
  _poly_gamma = p.poly_gamma;
  _sph_viscosity_alpha = p.sph_viscosity_alpha;
  _sph_viscosity_beta = p.sph_viscosity_beta;
  _sph_viscosity_epsilon = p.sph_viscosity_epsilon;
  _timestep_cfl_factor = p.timestep_cfl_factor;
  _initial_dt = p.initial_dt;
  _out_screen_dt = p.out_screen_dt;
  _out_scalar_dt = p.out_scalar_dt;
  _out_h5data_dt = p.out_h5data_dt;
  _fmm_macangle = p.fmm_macangle;
  _gravitational_constant = p.gravitational_constant;
  _box_length = p.box_length;
  _box_width = p.box_width;
  _box_height = p.box_height;
  _adaptive_timestep = p.adaptive_timestep;
  _sph_variable_h = p.sph_variable_h;
  _sph_update_uniform_h = p.sph_update_uniform_h;
  _evolve_internal_energy = p.evolve_internal_energy;
  _thermokinetic_formulation = p.thermokinetic_formulation;
  _enable_fmm = p.enable_fmm;
  _relaxation_steps = p.relaxation_steps;
  _pressure_updates_number = p.pressure_updates_number;
  _periodic_boundary_x = p.periodic_boundary_x;
  _periodic_boundary_y = p.periodic_boundary_y;
  _periodic_boundary_z = p.periodic_boundary_z;
  _sph_viscosity_alpha_max = p.sph_viscosity_alpha_max;
  _sph_viscosity_l = p.sph_viscosity_l;
  _sph_viscosity_delta = p.sph_viscosity_delta;
  _poly_gamma2 = p.poly_gamma2;
  _ppt_density_thr = p.ppt_density_thr;
  _relaxation_beta = p.relaxation_beta;
  _relaxation_gamma = p.relaxation_gamma;
  _relaxation_repulsion_radius = p.relaxation_repulsion_radius;
  _relaxation_repulsion_gamma = p.relaxation_repulsion_gamma;
  // external forces.  NOTE: the terms of eforce.h latch most of these in function-level statics at their first call
  // (e.g. `const static double box[3]`, eforce.h:51-55), so one process can exercise only one set of values per term.
  std::snprintf(_external_force_type, sizeof(_external_force_type), "%s", p.external_force_type[0] ? p.external_force_type : "none");
  std::snprintf(_density_profile, sizeof(_density_profile), "%s", p.density_profile[0] ? p.density_profile : "constant");
  _sphere_radius = p.sphere_radius;
  _extforce_wall_powerindex = p.extforce_wall_powerindex;
  _extforce_wall_steepness = p.extforce_wall_steepness;
  _gravity_acceleration_constant = p.gravity_acceleration_constant;
  _zero_potential_poison_value = p.zero_potential_poison_value;
  _mesa_rim_width = p.mesa_rim_width;
  _rho_initial = p.rho_initial;
  _pressure_initial = p.pressure_initial;
  _orbital_separation = p.orbital_separation;
  _mass_neutron_star = p.mass_neutron_star;
  _mass_white_dwarf = p.mass_white_dwarf;
  _airfoil_size = p.airfoil_size;
  _airfoil_thickness = p.airfoil_thickness;
  _airfoil_camber = p.airfoil_camber;
  _airfoil_anchor_x = p.airfoil_anchor_x;
  _airfoil_anchor_y = p.airfoil_anchor_y;
  _airfoil_attack_angle = p.airfoil_attack_angle;
  // what set_derived_params does (app/drivers/hydro/main_driver.cc:52-74)
  kernels::select();
  viscosity::select();
  eos::select();
  external_force::select(external_force_type);
  fmm::gc = gravitational_constant;
  physics::dt = p.initial_dt;
  physics::totaltime = 0.0;
  physics::iteration = 0;
  physics::dt_saved = 0.0;
  physics::totaltime_next = 0.0;
  physics::t_h5data_output = 0.0;
  physics::t_screen_output = 0.0;
  physics::t_scalar_output = 0.0;
}

extern "C" {

void
fsph_default_params(fsph_params * p, int32_t dim) {
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
  p->relaxation_repulsion_radius = 0.25;
  std::strcpy(p->external_force_type, "none");
  std::strcpy(p->density_profile, "constant");
  p->sphere_radius = 1.0;
  p->extforce_wall_powerindex = 5.0;
  p->extforce_wall_steepness = 1e12;
  p->gravity_acceleration_constant = 9.81;
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
}

const char *
fsph_backend(void) {
  return "reference";
}

fsph_system *
fsph_create(const fsph_params * p, int32_t) {
  if(p->dim != D)
    return nullptr;
  log_set_output_rank(0);
  fsph_system * s = new fsph_system;
  s->par = *p;
  install_params(*p);
  s->bs = new body_system<double, gdimension>();
  s->bs->setMacangle(p->fmm_macangle);
  return s;
}

void
fsph_destroy(fsph_system * s) {
  if(!s)
    return;
  delete s->bs;
  delete s;
}

const char *
fsph_last_error(const fsph_system * s) {
  return s ? s->err.c_str() : "null system";
}

int32_t
fsph_resize(fsph_system * s, int64_t n) {
  auto & v = s->bs->getLocalbodies();
  size_t old = v.size();
  v.resize(n);
  for(size_t i = old; i < (size_t)n; ++i) {
    std::memset((void *)&v[i], 0, sizeof(body));
    v[i].setType(NORMAL);
    v[i].set_id(i);
  }
  s->bs->read_bodies("", "", 0);
  return 0;
}

int64_t
fsph_nbodies(const fsph_system * s) {
  return s->bs->getLocalbodies().size();
}

static point_t
getvec(const body & b, int col) {
  switch(col) {
    case FSPH_COL_POS: return b.coordinates();
    case FSPH_COL_VEL: return b.getVelocity();
    case FSPH_COL_VELHALF: return b.getVelocityhalf();
    case FSPH_COL_ACC: return b.getAcceleration();
    default: return b.getGAcceleration();
  }
}

int32_t
fsph_set_column(fsph_system * s, int32_t col, const void * host) {
  auto & v = s->bs->getLocalbodies();
  const size_t n = v.size();
  const double * d = static_cast<const double *>(host);
  for(size_t i = 0; i < n; ++i) {
    body & b = v[i];
    point_t pt;
    if(col <= FSPH_COL_GACC)
      for(int k = 0; k < D; ++k)
        pt[k] = d[i * D + k];
    switch(col) {
      case FSPH_COL_POS: b.set_coordinates(pt); break;
      case FSPH_COL_VEL: b.setVelocity(pt); break;
      case FSPH_COL_VELHALF: b.setVelocityhalf(pt); break;
      case FSPH_COL_ACC: b.setAcceleration(pt); break;
      case FSPH_COL_GACC: b.setGAcceleration(pt); break;
      case FSPH_COL_MASS: b.set_mass(d[i]); break;
      case FSPH_COL_H: b.set_radius(d[i]); break;
      case FSPH_COL_RHO: b.setDensity(d[i]); break;
      case FSPH_COL_P: b.setPressure(d[i]); break;
      case FSPH_COL_U: b.setInternalenergy(d[i]); break;
      case FSPH_COL_E: b.setTotalenergy(d[i]); break;
      case FSPH_COL_CS: b.setSoundspeed(d[i]); break;
      case FSPH_COL_DUDT: b.setDudt(d[i]); break;
      case FSPH_COL_DEDT: b.setDedt(d[i]); break;
      case FSPH_COL_DT: b.setDt(d[i]); break;
      case FSPH_COL_ALPHA: b.setAlpha(d[i]); break;
      case FSPH_COL_VSIG: b.setSignalspeed(d[i]); break;
      case FSPH_COL_GPOT: b.setGPotential(d[i]); break;
      case FSPH_COL_ADIABATIC: b.setAdiabatic(d[i]); break;
      case FSPH_COL_YE: b.setElectronfraction(d[i]); break;
      case FSPH_COL_DIVV: b.setDivergenceV(d[i]); break;
      case FSPH_COL_DDIVVDT: b.setDdivvdt(d[i]); break;
      case FSPH_COL_ID: b.set_id(static_cast<const int64_t *>(host)[i]); break;
      case FSPH_COL_TYPE: b.setType(static_cast<const int32_t *>(host)[i]); break;
      default: s->err = "column not writable"; return 1;
    }
  }
  return 0;
}

int32_t
fsph_get_column(fsph_system * s, int32_t col, void * host) {
  auto & v = s->bs->getLocalbodies();
  const size_t n = v.size();
  double * d = static_cast<double *>(host);
  for(size_t i = 0; i < n; ++i) {
    body & b = v[i];
    if(col <= FSPH_COL_GACC) {
      point_t pt = getvec(b, col);
      for(int k = 0; k < D; ++k)
        d[i * D + k] = pt[k];
      continue;
    }
    switch(col) {
      case FSPH_COL_MASS: d[i] = b.mass(); break;
      case FSPH_COL_H: d[i] = b.radius(); break;
      case FSPH_COL_RHO: d[i] = b.getDensity(); break;
      case FSPH_COL_P: d[i] = b.getPressure(); break;
      case FSPH_COL_U: d[i] = b.getInternalenergy(); break;
      case FSPH_COL_E: d[i] = b.getTotalenergy(); break;
      case FSPH_COL_CS: d[i] = b.getSoundspeed(); break;
      case FSPH_COL_DUDT: d[i] = b.getDudt(); break;
      case FSPH_COL_DEDT: d[i] = b.getDedt(); break;
      case FSPH_COL_DT: d[i] = b.getDt(); break;
      case FSPH_COL_ALPHA: d[i] = b.getAlpha(); break;
      case FSPH_COL_VSIG: d[i] = b.getSignalspeed(); break;
      case FSPH_COL_GPOT: d[i] = b.getGPotential(); break;
      case FSPH_COL_ADIABATIC: d[i] = b.getAdiabatic(); break;
      case FSPH_COL_YE: d[i] = b.getElectronfraction(); break;
      case FSPH_COL_DIVV: d[i] = b.getDivergenceV(); break;
      case FSPH_COL_DDIVVDT: d[i] = b.getDdivvdt(); break;
      case FSPH_COL_TRIGGER: d[i] = b.getTrigger(); break;
      case FSPH_COL_XI: d[i] = b.getXi(); break;
      case FSPH_COL_TRACESS: d[i] = b.getTraceSS(); break;
      case FSPH_COL_GRADV: d[i] = b.getGradV(); break;
      case FSPH_COL_ID: static_cast<int64_t *>(host)[i] = b.id(); break;
      case FSPH_COL_KEY: static_cast<uint64_t *>(host)[i] = b.key().value(); break;
      case FSPH_COL_TYPE: static_cast<int32_t *>(host)[i] = b.type(); break;
      case FSPH_COL_NNEIGH:
        static_cast<int32_t *>(host)[i] = i < s->nbr.size() ? (int32_t)s->nbr[i].size() : 0;
        break;
      default: s->err = "bad column"; return 1;
    }
  }
  return 0;
}

int32_t
fsph_set_component(fsph_system * s, int32_t, int32_t, const void *) {
  s->err = "component transfers are a product-library entry point";
  return 1;
}
int32_t
fsph_get_component(fsph_system * s, int32_t, int32_t, void *) {
  s->err = "component transfers are a product-library entry point";
  return 1;
}
uint64_t
fsph_launch_count(void) {
  return 0;
}

int32_t
fsph_upload_aos(fsph_system * s, const void *, int64_t, int64_t, const int64_t *) {
  s->err = "AoS marshalling is a product-library entry point";
  return 1;
}
int32_t
fsph_download_aos(fsph_system * s, void *, int64_t, const int64_t *) {
  s->err = "AoS marshalling is a product-library entry point";
  return 1;
}

int32_t
fsph_set_time(fsph_system *, double dt, double totaltime, int64_t iteration) {
  physics::dt = dt;
  physics::totaltime = totaltime;
  physics::iteration = iteration;
  return 0;
}
int32_t
fsph_get_time(const fsph_system *, double * dt, double * totaltime, int64_t * iteration) {
  if(dt) *dt = physics::dt;
  if(totaltime) *totaltime = physics::totaltime;
  if(iteration) *iteration = physics::iteration;
  return 0;
}
int32_t
fsph_advance_time(fsph_system *) {
  physics::advance_time();
  return 0;
}

int32_t
fsph_update_iteration(fsph_system * s) {
  s->bs->update_iteration();
  return 0;
}
int32_t
fsph_reset_ghosts(fsph_system * s) {
  s->bs->reset_ghosts();
  return 0;
}

extern "C++" {
template<typename F>
static void
sweep_recording(fsph_system * s, F && f) {
  auto & v = s->bs->getLocalbodies();
  s->nbr.assign(v.size(), {});
  body * base = v.data();
  s->bs->apply_in_smoothinglength([&](body & p, std::vector<body *> & nbs) {
    auto & l = s->nbr[&p - base];
    l.resize(nbs.size());
    for(size_t k = 0; k < nbs.size(); ++k)
      l[k] = nbs[k]->id();
    std::sort(l.begin(), l.end());
    f(p, nbs);
  });
}
} // extern C++

int32_t
fsph_sweep(fsph_system * s, int32_t op) {
  switch(op) {
    case FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED:
      sweep_recording(s, physics::compute_density_pressure_soundspeed); break;
    case FSPH_SWEEP_ACCELERATION: sweep_recording(s, physics::compute_acceleration); break;
    case FSPH_SWEEP_DUDT: sweep_recording(s, physics::compute_dudt); break;
    case FSPH_SWEEP_DEDT: sweep_recording(s, physics::compute_dedt); break;
    case FSPH_SWEEP_NEIGHBORS_ONLY: sweep_recording(s, [](body &, std::vector<body *> &) {}); break;
    case FSPH_SWEEP_COMPUTE_ALPHA: sweep_recording(s, viscosity::compute_alpha); break;
    case FSPH_SWEEP_SHORT_RANGE_REPULSION: sweep_recording(s, physics::add_short_range_repulsion); break;
    default: s->err = "unknown sweep op"; return 1;
  }
  return 0;
}

// Same sweeps without the recording wrapper: what the CPU baseline times.
static void
sweep_plain(fsph_system * s, int32_t op) {
  switch(op) {
    case FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED:
      s->bs->apply_in_smoothinglength(physics::compute_density_pressure_soundspeed); break;
    case FSPH_SWEEP_ACCELERATION: s->bs->apply_in_smoothinglength(physics::compute_acceleration); break;
    case FSPH_SWEEP_DUDT: s->bs->apply_in_smoothinglength(physics::compute_dudt); break;
    case FSPH_SWEEP_DEDT: s->bs->apply_in_smoothinglength(physics::compute_dedt); break;
    case FSPH_SWEEP_COMPUTE_ALPHA: s->bs->apply_in_smoothinglength(viscosity::compute_alpha); break;
    case FSPH_SWEEP_SHORT_RANGE_REPULSION: s->bs->apply_in_smoothinglength(physics::add_short_range_repulsion); break;
  }
}

int32_t
fsph_apply(fsph_system * s, int32_t op) {
  auto & bs = *s->bs;
  switch(op) {
    case FSPH_APPLY_EOS_INIT: bs.apply_all(eos::init); break;
    case FSPH_APPLY_SET_TOTAL_ENERGY: bs.apply_all(physics::set_total_energy); break;
    case FSPH_APPLY_SAVE_VELOCITYHALF: bs.apply_all(integration::save_velocityhalf); break;
    case FSPH_APPLY_LEAPFROG_KICK_V: bs.apply_all(integration::leapfrog_kick_v); break;
    case FSPH_APPLY_LEAPFROG_KICK_U: bs.apply_all(integration::leapfrog_kick_u); break;
    case FSPH_APPLY_LEAPFROG_KICK_E: bs.apply_all(integration::leapfrog_kick_e); break;
    case FSPH_APPLY_LEAPFROG_DRIFT: bs.apply_all(integration::leapfrog_drift); break;
    case FSPH_APPLY_RECOMPUTE_P_CS: bs.apply_all(physics::recompute_pressure_soundspeed); break;
    case FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC:
      bs.apply_all(physics::recompute_pressure_soundspeed_thermokinetic); break;
    case FSPH_APPLY_COMPUTE_DT: bs.apply_all(physics::compute_dt); break;
    case FSPH_APPLY_CHECK_NANS: bs.apply_all(physics::check_nans); break;
    case FSPH_APPLY_CHECK_NEGATIVITY: bs.apply_all(physics::check_negativity); break;
    case FSPH_APPLY_INITIALIZE_ALPHA: bs.apply_all(viscosity::initialize_alpha); break;
    case FSPH_APPLY_ADD_DRAG_ACCELERATION: bs.apply_all(physics::add_drag_acceleration); break;
    case FSPH_APPLY_ADD_DRAG_DEDT: bs.apply_all(physics::add_drag_dedt); break;
    case FSPH_APPLY_ADD_DRAG_DUDT: bs.apply_all(physics::add_drag_dudt); break;
    default: s->err = "unknown apply op"; return 1;
  }
  return 0;
}

int32_t
fsph_reduce(fsph_system * s, int32_t op, double * out, int32_t nout) {
  auto & bs = *s->bs;
  switch(op) {
    case FSPH_GET_SET_ADAPTIVE_TIMESTEP:
      bs.get_all(physics::set_adaptive_timestep);
      if(out && nout > 0) out[0] = physics::dt;
      break;
    case FSPH_GET_COMPUTE_SMOOTHINGLENGTH: bs.get_all(physics::compute_smoothinglength); break;
    case FSPH_GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH:
      bs.get_all(physics::compute_average_smoothinglength, (int64_t)bs.getLocalbodies().size()); break;
    case FSPH_GET_TOTAL_MASS:
      bs.get_all(analysis::compute_total_mass);
      if(out && nout > 0) out[0] = analysis::total_mass;
      break;
    case FSPH_GET_TOTAL_ENERGY:
      bs.get_all(analysis::compute_total_energy);
      if(out && nout > 0) out[0] = analysis::total_energy;
      break;
    case FSPH_GET_TOTAL_MOMENTUM:
      bs.get_all(analysis::compute_total_momentum);
      for(int k = 0; k < D && k < nout; ++k) out[k] = analysis::linear_momentum[k];
      break;
    default: s->err = "unknown reduce op"; return 1;
  }
  return 0;
}

int32_t
fsph_gravitation_fmm(fsph_system * s) {
  if constexpr(gdimension == 3) {
    s->bs->gravitation_fmm();
    return 0;
  }
  s->err = "fmm needs 3 dimensions";
  return 1;
}

int32_t
fsph_get_range(fsph_system * s, double * lo, double * hi) {
  auto & r = s->bs->getRange();
  for(int k = 0; k < D; ++k) {
    lo[k] = r[0][k];
    hi[k] = r[1][k];
  }
  return 0;
}

double
fsph_get_smoothinglength(fsph_system * s) {
  return s->bs->getSmoothinglength();
}

// The driver loop body, calls in the order of app/drivers/hydro/main_driver.cc
// (first iteration :102-173, later iterations :175-245, tail :247-277) with
// the gravity call of newtonian/main_driver.cc:218-222 when enable_fmm.
int32_t
fsph_step(fsph_system * s, int32_t initial) {
  using namespace param;
  auto & bs = *s->bs;
  auto energy_passes = [&]() {
    for(int m = 1; m <= pressure_updates_number; ++m) {
      if(thermokinetic_formulation) {
        sweep_plain(s, FSPH_SWEEP_DEDT);
        if(physics::iteration < relaxation_steps)
          bs.apply_all(physics::add_drag_dedt);
        bs.apply_all(physics::recompute_pressure_soundspeed_thermokinetic);
      }
      else {
        sweep_plain(s, FSPH_SWEEP_DUDT);
        if(physics::iteration < relaxation_steps)
          bs.apply_all(physics::add_drag_dudt);
        bs.apply_all(physics::recompute_pressure_soundspeed);
      }
      if(m < pressure_updates_number)
        bs.reset_ghosts();
    }
  };
  if(initial) {
    bs.update_iteration();
    bs.apply_all(eos::init);
    if(thermokinetic_formulation)
      bs.apply_all(physics::set_total_energy);
    if(sph_viscosity != visc_constant)
      bs.apply_all(viscosity::initialize_alpha);
    sweep_plain(s, FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED);
    bs.apply_all(integration::save_velocityhalf);
    if(sph_viscosity != visc_constant)
      sweep_plain(s, FSPH_SWEEP_COMPUTE_ALPHA);
    bs.reset_ghosts();
    sweep_plain(s, FSPH_SWEEP_ACCELERATION);
    if(physics::iteration < relaxation_steps) {
      bs.apply_all(physics::add_drag_acceleration);
      sweep_plain(s, FSPH_SWEEP_SHORT_RANGE_REPULSION);
    }
    if(enable_fmm)
      bs.gravitation_fmm();
    if(adaptive_timestep) {
      bs.apply_all(physics::compute_dt);
      bs.get_all(physics::set_adaptive_timestep);
    }
    if(evolve_internal_energy)
      energy_passes();
  }
  else {
    if(evolve_internal_energy) {
      if(thermokinetic_formulation)
        bs.apply_all(integration::leapfrog_kick_e);
      else
        bs.apply_all(integration::leapfrog_kick_u);
    }
    bs.apply_all(integration::leapfrog_kick_v);
    bs.apply_all(integration::save_velocityhalf);
    bs.apply_all(integration::leapfrog_drift);
    bs.update_iteration();
    sweep_plain(s, FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED);
    if(sph_viscosity != visc_constant)
      sweep_plain(s, FSPH_SWEEP_COMPUTE_ALPHA);
    bs.reset_ghosts();
    sweep_plain(s, FSPH_SWEEP_ACCELERATION);
    if(physics::iteration < relaxation_steps) {
      bs.apply_all(physics::add_drag_acceleration);
      sweep_plain(s, FSPH_SWEEP_SHORT_RANGE_REPULSION);
    }
    if(enable_fmm)
      bs.gravitation_fmm();
    bs.apply_all(integration::leapfrog_kick_v);
    bs.reset_ghosts();
    if(evolve_internal_energy) {
      energy_passes();
      if(thermokinetic_formulation)
        bs.apply_all(integration::leapfrog_kick_e);
      else
        bs.apply_all(integration::leapfrog_kick_u);
    }
  }
  if(sph_variable_h)
    bs.get_all(physics::compute_smoothinglength);
  else if(sph_update_uniform_h)
    bs.get_all(physics::compute_average_smoothinglength, (int64_t)bs.getLocalbodies().size());
  bs.apply_all(physics::check_nans);
  bs.apply_all(physics::check_negativity);
  if(adaptive_timestep) {
    bs.apply_all(physics::compute_dt);
    bs.get_all(physics::set_adaptive_timestep);
  }
  physics::advance_time();
  return 0;
}

// ---- introspection -------------------------------------------------------
using hcell_t = flecsi::topology::hcell<gdimension, key_type, node, body>;

static void
collect_nodes(fsph_system * s, std::vector<node *> & nodes) {
  auto * t = s->bs->tree();
  t->traversal(t->root(), [&](hcell_t * c) {
    if(c->is_node()) {
      nodes.push_back(t->get_node(c));
      return true;
    }
    return false;
  });
  std::sort(nodes.begin(), nodes.end(), [](node * a, node * b) { return a->key().value() < b->key().value(); });
}

int64_t
fsph_tree_num_nodes(fsph_system * s) {
  std::vector<node *> nodes;
  collect_nodes(s, nodes);
  return nodes.size();
}

int32_t
fsph_tree_get_nodes(fsph_system * s, uint64_t * key, double * mass, double * com, double * radius, double * lap,
  double * bmin, double * bmax, int64_t * sub_entities) {
  std::vector<node *> nodes;
  collect_nodes(s, nodes);
  for(size_t i = 0; i < nodes.size(); ++i) {
    node * c = nodes[i];
    key[i] = c->key().value();
    mass[i] = c->mass();
    radius[i] = c->radius();
    lap[i] = c->lap();
    sub_entities[i] = c->sub_entities();
    for(int k = 0; k < D; ++k) {
      com[i * D + k] = c->coordinates()[k];
      bmin[i * D + k] = c->bmin()[k];
      bmax[i * D + k] = c->bmax()[k];
    }
  }
  return 0;
}

int32_t
fsph_tree_get_leaf_depth(fsph_system * s, int32_t * depth) {
  auto * t = s->bs->tree();
  body * base = s->bs->getLocalbodies().data();
  t->traversal(t->root(), [&](hcell_t * c) {
    if(c->is_node())
      return true;
    depth[t->get_entity(c) - base] = (int32_t)c->key().depth();
    return false;
  });
  return 0;
}

int32_t
fsph_neighbor_lists(fsph_system * s, int64_t * offsets, int64_t * ids) {
  int64_t o = 0;
  for(size_t i = 0; i < s->nbr.size(); ++i) {
    offsets[i] = o;
    if(ids)
      std::copy(s->nbr[i].begin(), s->nbr[i].end(), ids + o);
    o += s->nbr[i].size();
  }
  offsets[s->nbr.size()] = o;
  return 0;
}

int32_t
fsph_phase_times(fsph_system *, double *, int32_t, const char ** names) {
  if(names) *names = "";
  return 0;
}

} // extern "C"
