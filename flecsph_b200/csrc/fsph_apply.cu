// fsph_apply.cu -- element-wise passes and scalar reductions (compiled with -fmad=false so that positions and
// velocities after a kick/drift are the same IEEE results as the reference's scalar code; HBM-bound).
//
// Replaces the functors passed to body_system::apply_all / get_all (mpisph/bodies_system.h:318,335):
//   integration::save_velocityhalf / leapfrog_kick_{v,u,e} / leapfrog_drift   include/physics/integration.h:52-110
//   eos::init (ideal -> adiabatic K = P / rho^gamma)                          include/physics/eos/eos.h:68-83,157-159
//   physics::set_total_energy :327-335, recompute_pressure_soundspeed[_thermokinetic] :153-182,
//   compute_dt :602-695, set_adaptive_timestep :761-823, compute_smoothinglength :825-848,
//   compute_average_smoothinglength :879-898, check_nans :918-958, check_negativity :962-979
//                                                                             include/physics/default_physics.h
//   analysis::compute_total_{mass,energy,momentum}                            include/physics/analysis.h:56-150
//   external_force::acceleration / potential where those functors call them    include/physics/eforce.h (fsph_eforce.h)
#include <stdexcept>
#include <cfloat>
#include <cmath>
#include "fsph_internal.h"

namespace fsph {

Cols System::cols() const {
  Cols c;
  for(int f = 0; f < NF64; ++f) c.f[f] = col[f].as<double>();
  c.id = col_id.as<int64_t>();
  c.key = col_key.as<uint64_t>();
  c.type = col_type.as<int32_t>();
  c.nneigh = col_nneigh.as<int32_t>();
  c.own_lo = has_ghosts ? own_begin : 0;
  c.own_hi = has_ghosts ? own_begin + n_owned : n;
  return c;
}

PhysParams System::phys() const {
  PhysParams p;
  p.kernel = par.sph_kernel;
  p.thermokinetic = par.thermokinetic_formulation;
  p.evolve_u = par.evolve_internal_energy;
  p.eos = par.eos_type;
  p.eta = par.sph_eta;
  p.gamma = par.poly_gamma;
  p.alpha = par.sph_viscosity_alpha;
  p.beta = par.sph_viscosity_beta;
  p.epsilon = par.sph_viscosity_epsilon;
  p.cfl = par.timestep_cfl_factor;
  p.kernel_width = kernel_width;
  p.dt = dt;
  p.iteration = iteration;
  p.relaxation_steps = par.relaxation_steps;
  p.visc = par.sph_viscosity;
  p.alpha_max = par.sph_viscosity_alpha_max;
  p.visc_l = par.sph_viscosity_l;
  p.visc_delta = par.sph_viscosity_delta;
  p.gamma2 = par.poly_gamma2;
  p.rho_thr = par.ppt_density_thr;
  p.relax_beta = par.relaxation_beta;
  p.relax_gamma = par.relaxation_gamma;
  p.repulsion_radius = par.relaxation_repulsion_radius;
  p.repulsion_gamma = par.relaxation_repulsion_gamma;
  p.fuse_kick_v = fuse_kick_v ? 1 : 0;
  p.ext_potential = ef.n > 0 ? 1 : 0;
  return p;
}

struct OpList {
  int n;
  int op[8];
};

// Columns an element-wise functor reads / writes (bit f = F64 column f; vector columns stand for their D components).
struct OpAccess { uint64_t rd, wr; };
constexpr uint64_t fbit(int f) { return uint64_t(1) << f; }
constexpr uint64_t vbit(int f) { return fbit(f) | fbit(f + 1) | fbit(f + 2); }
static OpAccess op_access(int op, const PhysParams & pp, const ExtForce & ef) {
  // functors that evaluate the external potential / force read the position as well
  const uint64_t pot_in = ef.n > 0 ? vbit(F_POS) : 0;
  const uint64_t eos_in = fbit(F_RHO) | fbit(F_ADIA) | (pp.eos == FSPH_EOS_WD ? fbit(F_YE) : 0) | fbit(F_P) | fbit(F_CS);
  switch(op) {
    case FSPH_APPLY_EOS_INIT: return {fbit(F_RHO) | fbit(F_P) | (pp.eos == FSPH_EOS_WD ? fbit(F_YE) : 0), fbit(F_ADIA) | fbit(F_U)};
    case FSPH_APPLY_SET_TOTAL_ENERGY: return {vbit(F_VEL) | fbit(F_U) | pot_in, fbit(F_E)};
    case FSPH_APPLY_SAVE_VELOCITYHALF: return {vbit(F_VEL), vbit(F_VELH)};
    case FSPH_APPLY_LEAPFROG_KICK_V: return {vbit(F_VEL) | vbit(F_ACC) | vbit(F_GACC), vbit(F_VEL)};
    case FSPH_APPLY_LEAPFROG_KICK_U: return {fbit(F_U) | fbit(F_DUDT), fbit(F_U)};
    case FSPH_APPLY_LEAPFROG_KICK_E: return {fbit(F_E) | fbit(F_DEDT), fbit(F_E)};
    case FSPH_APPLY_LEAPFROG_DRIFT: return {vbit(F_POS) | vbit(F_VEL), vbit(F_POS)};
    case FSPH_APPLY_RECOMPUTE_P_CS: return {fbit(F_U) | fbit(F_DUDT) | eos_in, fbit(F_P) | fbit(F_CS)};
    case FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC:
      return {vbit(F_VEL) | vbit(F_ACC) | fbit(F_E) | fbit(F_DEDT) | eos_in | pot_in, fbit(F_U) | fbit(F_P) | fbit(F_CS)};
    case FSPH_APPLY_COMPUTE_DT: return {fbit(F_H) | vbit(F_VEL) | vbit(F_ACC) | vbit(F_GACC) | fbit(F_VSIG) | fbit(F_U) | pot_in, fbit(F_DT)};
    case FSPH_APPLY_CHECK_NANS:
      return {vbit(F_POS) | vbit(F_VEL) | vbit(F_VELH) | vbit(F_ACC) | vbit(F_GACC) | fbit(F_MASS) | fbit(F_GPOT) | fbit(F_RHO) | fbit(F_P) |
                fbit(F_E) | fbit(F_DEDT) | fbit(F_DUDT) | fbit(F_ADIA) | fbit(F_VSIG) | fbit(F_U), 0};
    case FSPH_APPLY_CHECK_NEGATIVITY: return {fbit(F_RHO) | fbit(F_P) | fbit(F_U), 0};
    case FSPH_APPLY_INITIALIZE_ALPHA: return {0, fbit(F_ALPHA)};
    case FSPH_APPLY_ADD_DRAG_ACCELERATION: return {vbit(F_VEL) | vbit(F_ACC), vbit(F_ACC)};
    case FSPH_APPLY_ADD_DRAG_DEDT: return {vbit(F_VEL) | fbit(F_DEDT), fbit(F_DEDT)};
    case FSPH_APPLY_ADD_DRAG_DUDT: return {vbit(F_VEL) | fbit(F_DUDT), fbit(F_DUDT)};
    case 100: return {fbit(F_MASS) | fbit(F_RHO), fbit(F_H)};
    case 101: return {0, fbit(F_H)};
    case 102: return {vbit(F_POS) | vbit(F_ACC), vbit(F_ACC)};
    case 103: return {vbit(F_POS) | vbit(F_VEL) | fbit(F_E), fbit(F_U)};
  }
  throw std::runtime_error("unknown apply op");
}

// One functor on the body's columns held in registers (v[f] = column f of this body): the statements of the reference's
// functors, operation by operation (this file is compiled without FMA contraction).
// EF: the system has external force / potential terms (a separate instantiation: the common case without them keeps its
// register budget); the position is copied out of the column registers so that the terms never index `v` dynamically
template<int D, bool EF>
__device__ __forceinline__ void apply_one(double * v, const int op, const PhysParams & pp, const ExtForce & ef, uint32_t * __restrict__ status) {
  const double dt = pp.dt;
  double xp[3] = {0.0, 0.0, 0.0};
  if(EF) {
#pragma unroll
    for(int d = 0; d < D; ++d) xp[d] = v[F_POS + d];
  }
  switch(op) {
    case FSPH_APPLY_EOS_INIT: {
      // eos::init: adiabatic invariant (ideal / polytropic eos.h:68-83; piecewise polytrope :337-351: K1 of the first
      // segment), internal energy of the cold white dwarf (:292-302)
      const double rho = v[F_RHO], P = v[F_P];
      if(pp.eos == FSPH_EOS_PPT) {
        double K1;
        if(rho < pp.rho_thr) K1 = P / pow(rho, pp.gamma);
        else {
          const double K2 = P / pow(rho, pp.gamma2);
          K1 = K2 * pow(pp.rho_thr, pp.gamma2 - pp.gamma);
        }
        v[F_ADIA] = K1;
      }
      else if(pp.eos == FSPH_EOS_WD) {
        const double A_wd = 6.00233181e22, B_wd_nm = 9.73932099e5;
        const double x = cbrt(rho * v[F_YE] / B_wd_nm), x2 = x * x, x3 = x * x * x;
        v[F_U] = A_wd / rho * (8. * x3 * (sqrt(x2 + 1.) - 1.) - (x * (2. * x2 - 3.) * sqrt(x2 + 1.) + 3. * asinh(x)));
      }
      else if(pp.eos != FSPH_EOS_NONE) v[F_ADIA] = P / pow(rho, pp.gamma);
    } break;
    case FSPH_APPLY_SET_TOTAL_ENERGY: {
      double v2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) v2 = v2 + v[F_VEL + d] * v[F_VEL + d];
      // ekin + eint + epot, default_physics.h:330-334
      v[F_E] = .5 * v2 + v[F_U] + (EF ? ef_potential<D>(ef, xp) : 0.0);
    } break;
    case FSPH_APPLY_SAVE_VELOCITYHALF:
#pragma unroll
      for(int d = 0; d < D; ++d) v[F_VELH + d] = v[F_VEL + d];
      break;
    case FSPH_APPLY_LEAPFROG_KICK_V:
#pragma unroll
      for(int d = 0; d < D; ++d) v[F_VEL + d] = v[F_VEL + d] + (0.5 * dt) * (v[F_ACC + d] + v[F_GACC + d]);
      break;
    case FSPH_APPLY_LEAPFROG_KICK_U: v[F_U] = v[F_U] + 0.5 * dt * v[F_DUDT]; break;
    case FSPH_APPLY_LEAPFROG_KICK_E: v[F_E] = v[F_E] + 0.5 * dt * v[F_DEDT]; break;
    case FSPH_APPLY_LEAPFROG_DRIFT:
#pragma unroll
      for(int d = 0; d < D; ++d) v[F_POS + d] = v[F_POS + d] + dt * v[F_VEL + d];
      break;
    case FSPH_APPLY_RECOMPUTE_P_CS: {
      const double u = v[F_U] + 0.5 * dt * v[F_DUDT];
      eos_p_cs(pp, v[F_RHO], u, v[F_ADIA], v[F_YE], v[F_P], v[F_CS]);
    } break;
    case FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC: {
      double v2 = 0.0, va = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        v2 = v2 + v[F_VEL + d] * v[F_VEL + d];
        va = va + v[F_VEL + d] * v[F_ACC + d];
      }
      // recover_internal_energy, default_physics.h:126-144
      const double eint = v[F_E] - .5 * v2 - (EF ? ef_potential<D>(ef, xp) : 0.0);
      if(!(eint > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
      v[F_U] = eint;
      const double u = eint + 0.5 * dt * (v[F_DEDT] - va);
      eos_p_cs(pp, v[F_RHO], u, v[F_ADIA], v[F_YE], v[F_P], v[F_CS]);
    } break;
    case FSPH_APPLY_COMPUTE_DT: {
      const double tiny = 1e-24, mc = 0.6;
      const double dx = v[F_H] / (pp.eta * pp.kernel_width);
      double v2 = 0.0, a2 = 0.0, vg = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        const double w = v[F_VEL + d], a = v[F_ACC + d] + v[F_GACC + d];
        v2 = v2 + w * w;
        a2 = a2 + a * a;
        vg = vg + w * v[F_GACC + d];
      }
      const double vn = D == 1 ? fabs(v[F_VEL]) : sqrt(v2);
      const double acc = D == 1 ? fabs(v[F_ACC] + v[F_GACC]) : sqrt(a2);
      const double dt_v = dx / (vn + tiny);
      const double dt_a = sqrt(dx / (acc + tiny));
      const double dt_c = dx / (tiny + v[F_VSIG] * (1 + mc * pp.alpha));
      double dtmin = pp.cfl * fmin(fmin(dt_v, dt_a), dt_c);
      if(pp.evolve_u && pp.thermokinetic) {
        const double eint = v[F_U];
        const double epot = EF ? ef_potential<D>(ef, xp) : 0.0;
        if(!(eint > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
        int it;
        for(it = 0; it < 20; ++it) {
          double delta_epot = 0.0;
          if(EF) {
            // external_force::potential(pos + dtmin*vel) - epot, default_physics.h:655
            double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for(int d = 0; d < D; ++d) xn[d] = xp[d] + dtmin * v[F_VEL + d];
            delta_epot = ef_potential<D>(ef, xn) - epot;
          }
          const double delta_egrv = -dtmin * vg;
          if(delta_epot + delta_egrv < eint * 0.5) break;
          dtmin *= 0.25;
        }
        if(it >= 20) atomicOr(status, (uint32_t)ERR_DT_LOOP);
      }
      v[F_DT] = dtmin;
    } break;
    case FSPH_APPLY_CHECK_NANS: {
      bool bad = false;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        bad = bad || isnan(v[F_POS + d]) || isnan(v[F_VEL + d]) || isnan(v[F_VELH + d]) || isnan(v[F_ACC + d]) || isnan(v[F_GACC + d]);
      }
      bad = bad || isnan(v[F_MASS]) || isnan(v[F_GPOT]) || isnan(v[F_RHO]) || isnan(v[F_P]) || isnan(v[F_E]) || isnan(v[F_DEDT]) ||
            isnan(v[F_DUDT]) || isnan(v[F_ADIA]) || isnan(v[F_VSIG]);
      if(pp.evolve_u) bad = bad || isnan(v[F_U]);
      if(bad) atomicOr(status, (uint32_t)ERR_NAN);
    } break;
    case FSPH_APPLY_CHECK_NEGATIVITY:
      if(v[F_RHO] < 0 || v[F_P] < 0 || (pp.evolve_u && v[F_U] < 0)) atomicOr(status, (uint32_t)ERR_NEGATIVE);
      break;
    case FSPH_APPLY_INITIALIZE_ALPHA:   // viscosity.h:172-180
      v[F_ALPHA] = pp.visc == FSPH_VISC_CONSTANT ? pp.alpha : 0.0;
      break;
    case FSPH_APPLY_ADD_DRAG_ACCELERATION:
    case FSPH_APPLY_ADD_DRAG_DEDT:
    case FSPH_APPLY_ADD_DRAG_DUDT: {
      // external_force::acceleration_drag (eforce.h:441-451): a = 0 - (beta + gamma v^2) v
      double v2 = v[F_VEL] * v[F_VEL];
#pragma unroll
      for(int d = 1; d < D; ++d) v2 = v2 + v[F_VEL + d] * v[F_VEL + d];
      const double k = pp.relax_beta + pp.relax_gamma * v2;
      double va = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        const double w = v[F_VEL + d], a = 0.0 - k * w;
        if(op == FSPH_APPLY_ADD_DRAG_ACCELERATION) v[F_ACC + d] = v[F_ACC + d] + a;
        va = va + w * a;
      }
      if(op == FSPH_APPLY_ADD_DRAG_DEDT) v[F_DEDT] = v[F_DEDT] + va;
      if(op == FSPH_APPLY_ADD_DRAG_DUDT) v[F_DUDT] = v[F_DUDT] + va;
    } break;
    case 100: {  // compute_smoothinglength
      const double q = v[F_MASS] / v[F_RHO];
      const double r = D == 1 ? q : (D == 2 ? sqrt(q) : cbrt(q));
      v[F_H] = r * pp.eta * pp.kernel_width;
    } break;
    case 101: v[F_H] = pp.dt; break;  // uniform h (value carried in pp.dt)
    case 103: {
      // recover_internal_energy (default_physics.h:126-144) as compute_density_pressure_soundspeed calls it (:313-314) when
      // there is an external potential: issued right before the density sweep, which then takes u as it finds it
      double v2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) v2 = v2 + v[F_VEL + d] * v[F_VEL + d];
      const double eint = v[F_E] - .5 * v2 - (EF ? ef_potential<D>(ef, xp) : 0.0);
      if(!(eint > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
      v[F_U] = eint;
    } break;
    case 102: {  // acc_a += external_force::acceleration(particle), the last statement of compute_acceleration (default_physics.h:404)
      if(EF) {
        double a[3] = {0.0, 0.0, 0.0};
        ef_acceleration<D>(ef, xp, a);
#pragma unroll
        for(int d = 0; d < D; ++d) v[F_ACC + d] = v[F_ACC + d] + a[d];
      }
    } break;
  }
}

// Several element-wise functors of the driver loop in ONE pass over HBM: every column any of them reads is loaded once
// up front (all loads in flight together), the functors run in order on the registers, every column any of them wrote is
// stored once.  `need` / `dirty`: the column sets (bit f = F64 column f), from op_access on the host.
#ifndef FSPH_APPLY_OCC
#define FSPH_APPLY_OCC 4
#endif
template<int D, bool EF>
__global__ void __launch_bounds__(128, FSPH_APPLY_OCC) k_apply(Cols c, int64_t n, OpList ops, PhysParams pp, ExtForce ef, uint64_t need, uint64_t dirty,
  uint32_t * __restrict__ status) {
  // n = number of owned bodies, starting at c.own_lo (remote copies are refreshed from their owner: reset_ghosts)
  const int64_t i = c.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= c.own_lo + n) return;
  double v[NF64];
#pragma unroll
  for(int f = 0; f < NF64; ++f) {
    v[f] = 0.0;
    if(f < F_MASS && (f % 3) >= D) continue;   // unused vector components
    if((need >> f) & 1ull) v[f] = c.f[f][i];
  }
  for(int k = 0; k < ops.n; ++k) apply_one<D, EF>(v, ops.op[k], pp, ef, status);
#pragma unroll
  for(int f = 0; f < NF64; ++f) {
    if(f < F_MASS && (f % 3) >= D) continue;
    if((dirty >> f) & 1ull) c.f[f][i] = v[f];
  }
}

void apply_fused(System & s, const int * op_list, int nops) {
  if(s.n <= 0 || nops <= 0) return;
  if(nops > 8) throw std::runtime_error("apply_fused: too many ops");
  const int64_t n_own = s.n_owned_or_n();
  unsigned nb = (unsigned)((n_own + 127) / 128);
  uint32_t * st = s.b_status.as<uint32_t>();
  const PhysParams pp = s.phys();
  OpList ops;
  ops.n = nops;
  uint64_t need = 0, dirty = 0;
  for(int j = 0; j < 8; ++j) ops.op[j] = j < nops ? op_list[j] : -1;
  for(int k = 0; k < nops; ++k) {
    const OpAccess a = op_access(op_list[k], pp, s.ef);
    need |= a.rd | a.wr;   // a functor may write a column only for some bodies (e.g. eos::init): keep the others' values
    dirty |= a.wr;
  }
  for(int f = 0; f < NF64; ++f)
    if(!s.col_used[f]) { need &= ~fbit(f); dirty &= ~fbit(f); }
  const bool efo = s.ef.n > 0;
#define FSPH_APPLY_GO(DD, EE) FSPH_LAUNCH((k_apply<DD, EE>), nb, 128, s.stream, s.cols(), n_own, ops, pp, s.ef, need, dirty, st)
  if(s.dim == 1) { if(efo) FSPH_APPLY_GO(1, true); else FSPH_APPLY_GO(1, false); }
  else if(s.dim == 2) { if(efo) FSPH_APPLY_GO(2, true); else FSPH_APPLY_GO(2, false); }
  else { if(efo) FSPH_APPLY_GO(3, true); else FSPH_APPLY_GO(3, false); }
#undef FSPH_APPLY_GO
  FSPH_CUDA_OK(cudaGetLastError());
  for(int k = 0; k < nops; ++k) {
    // the cached pair viscosities depend on v_1/2, rho and c_s (fsph_sweeps.cu)
    if(op_list[k] == FSPH_APPLY_SAVE_VELOCITYHALF || op_list[k] == FSPH_APPLY_RECOMPUTE_P_CS ||
       op_list[k] == FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC)
      s.pair_pi_valid = false;
  }
  for(int k = 0; k < nops; ++k)
    if(op_list[k] == FSPH_APPLY_LEAPFROG_DRIFT || op_list[k] == 100 || op_list[k] == 101) {
      // positions / smoothing lengths moved: the tree and the cached neighbour structure are stale
      s.tree_dirty = true;
      s.search_valid = false;
    }
}

void apply(System & s, int op) { apply_fused(s, &op, 1); }

// ---- reductions: per-block partials in a fixed order, finished on the host (deterministic) ------------------
enum RedKind { RED_MIN_DT = 0, RED_MAX_H, RED_SUM_H, RED_MASS, RED_ENERGY, RED_MOMENTUM, RED_EKIN, RED_EINT, RED_EGRAV, RED_ANGMOM,
               RED_H_STATS, RED_V_STATS, RED_N_STATS };
// kinds >= RED_H_STATS reduce {min, max, sum} in slots 0, 1, 2
__host__ __device__ inline double red_combine(int kind, int j, double a, double b) {
  if(kind == RED_MIN_DT) return fmin(a, b);
  if(kind == RED_MAX_H) return fmax(a, b);
  if(kind >= RED_H_STATS) return j == 0 ? fmin(a, b) : (j == 1 ? fmax(a, b) : a + b);
  return a + b;
}

template<int D>
__global__ void __launch_bounds__(256) k_reduce(Cols c, int64_t n, int kind, int thermokinetic, int fmm, ExtForce ef, double * __restrict__ part) {
  double v[3] = {0, 0, 0};
  if(kind == RED_MIN_DT) v[0] = 1e24;
  if(kind == RED_MAX_H) v[0] = -DBL_MAX;
  if(kind >= RED_H_STATS) { v[0] = DBL_MAX; v[1] = kind == RED_N_STATS ? 0.0 : DBL_MIN; }
  for(int64_t i = c.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.own_hi; i += (int64_t)gridDim.x * blockDim.x) {
    switch(kind) {
      case RED_MIN_DT: v[0] = fmin(v[0], c.f[F_DT][i]); break;
      case RED_MAX_H: v[0] = fmax(v[0], c.f[F_H][i]); break;
      case RED_SUM_H: v[0] += c.f[F_H][i]; break;
      case RED_MASS:
        if(c.type[i] == 0) v[0] += c.f[F_MASS][i];
        break;
      case RED_ENERGY:
        if(c.type[i] == 0) {
          const double m = c.f[F_MASS][i];
          if(thermokinetic) v[0] += m * c.f[F_E][i];
          else {
            double v2 = 0.0;
#pragma unroll
            for(int d = 0; d < D; ++d) v2 += c.f[F_VEL + d][i] * c.f[F_VEL + d][i];
            // m*(ekin + eint + epot), analysis.h:104-112
            double epot = 0.0;
            if(ef.n > 0) {
              double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
              for(int d = 0; d < D; ++d) x[d] = c.f[F_POS + d][i];
              epot = ef_potential<D>(ef, x);
            }
            v[0] += m * (.5 * v2 + c.f[F_U][i] + epot);
          }
          if(fmm) v[0] += .5 * m * c.f[F_GPOT][i];
        }
        break;
      case RED_MOMENTUM:
        if(c.type[i] == 0) {
#pragma unroll
          for(int d = 0; d < D; ++d) v[d] += c.f[F_MASS][i] * c.f[F_VEL + d][i];
        }
        break;
      case RED_EKIN:
        if(c.type[i] == 0) {
          double v2 = 0.0;
#pragma unroll
          for(int d = 0; d < D; ++d) v2 += c.f[F_VEL + d][i] * c.f[F_VEL + d][i];
          v[0] += .5 * c.f[F_MASS][i] * v2;
        }
        break;
      case RED_EINT:
        if(c.type[i] == 0) v[0] += c.f[F_MASS][i] * c.f[F_U][i];
        break;
      case RED_EGRAV:
        if(c.type[i] == 0) v[0] += c.f[F_GPOT][i] * c.f[F_MASS][i];
        break;
      case RED_ANGMOM:
        if(c.type[i] == 0) {
          const double m = c.f[F_MASS][i];
          if(D == 2) v[0] += m * (c.f[F_POS][i] * c.f[F_VEL + 1][i] - c.f[F_POS + 1][i] * c.f[F_VEL][i]);
          if(D == 3) {
            const double r0 = c.f[F_POS][i], r1 = c.f[F_POS + 1][i], r2 = c.f[F_POS + (D > 2 ? 2 : 0)][i];
            const double w0 = c.f[F_VEL][i], w1 = c.f[F_VEL + 1][i], w2 = c.f[F_VEL + (D > 2 ? 2 : 0)][i];
            v[0] += m * (r1 * w2 - r2 * w1);
            v[1] += m * (r2 * w0 - r0 * w2);
            v[2] += m * (r0 * w1 - r1 * w0);
          }
        }
        break;
      case RED_H_STATS: {
        const double h = c.f[F_H][i];
        v[0] = fmin(v[0], h); v[1] = fmax(v[1], h); v[2] += h;
      } break;
      case RED_V_STATS: {
        double v2 = 0.0;
#pragma unroll
        for(int d = 0; d < D; ++d) v2 += c.f[F_VEL + d][i] * c.f[F_VEL + d][i];
        const double vn = sqrt(v2);
        v[0] = fmin(v[0], vn); v[1] = fmax(v[1], vn); v[2] += vn;
      } break;
      case RED_N_STATS: {
        const double nn = (double)c.nneigh[i];
        v[0] = fmin(v[0], nn); v[1] = fmax(v[1], nn); v[2] += nn;
      } break;
    }
  }
  __shared__ double sm[8][3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for(int j = 0; j < 3; ++j) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
      double u = __shfl_xor_sync(0xffffffffu, v[j], o);
      v[j] = red_combine(kind, j, v[j], u);
    }
    if(lane == 0) sm[w][j] = v[j];
  }
  __syncthreads();
  if(threadIdx.x < 3) {
    double r = sm[0][threadIdx.x];
    for(int ww = 1; ww < 8; ++ww) {
      double u = sm[ww][threadIdx.x];
      r = red_combine(kind, (int)threadIdx.x, r, u);
    }
    part[(size_t)blockIdx.x * 3 + threadIdx.x] = r;
  }
}

static void run_reduce_local(System & s, int kind, double out[3]) {
  if(s.n <= 0) {
    // empty system: the identity of each reduction (no launch with an empty grid)
    out[0] = kind == RED_MIN_DT ? 1e24 : (kind == RED_MAX_H ? -DBL_MAX : (kind >= RED_H_STATS ? DBL_MAX : 0.0));
    out[1] = kind >= RED_H_STATS ? (kind == RED_N_STATS ? 0.0 : DBL_MIN) : 0.0;
    out[2] = 0.0;
    return;
  }
  int nb = (int)std::min<int64_t>((s.n + 255) / 256, 148 * 4);
  double * part = reduce_scratch(s, (size_t)nb * 3 + 16);
  if(s.dim == 1) FSPH_LAUNCH((k_reduce<1>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, s.ef, part);
  else if(s.dim == 2) FSPH_LAUNCH((k_reduce<2>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, s.ef, part);
  else FSPH_LAUNCH((k_reduce<3>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, s.ef, part);
  FSPH_CUDA_OK(cudaGetLastError());
  std::vector<double> h((size_t)nb * 3);
  FSPH_CUDA_OK(cudaMemcpyAsync(h.data(), part, h.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  for(int j = 0; j < 3; ++j) {
    double r = h[j];
    for(int b = 1; b < nb; ++b) {
      double u = h[(size_t)b * 3 + j];
      r = red_combine(kind, j, r, u);
    }
    out[j] = r;
  }
}

void global_combine(System & s, int kind, double v[3]) {
  const int P = s.comm->nranks;
  std::vector<double> all((size_t)3 * P);
  dist_all_gather_host(s, v, all.data(), 3 * sizeof(double));
  for(int j = 0; j < 3; ++j) {
    double r = all[j];
    for(int q = 1; q < P; ++q) r = red_combine(kind, j, r, all[(size_t)3 * q + j]);
    v[j] = r;
  }
}

// get_all reductions end in an MPI_Allreduce in the reference (default_physics.h:769,891; analysis.h:60-200)
static void run_reduce(System & s, int kind, double out[3]) {
  run_reduce_local(s, kind, out);
  if(s.comm) global_combine(s, kind, out);
}

double max_h(System & s) {
  double o[3];
  run_reduce_local(s, RED_MAX_H, o);
  return o[0];
}

// physics::set_adaptive_timestep, default_physics.h:761-823 (host part, after the device min)
static void set_adaptive_timestep(System & s, double dtmin) {
  double & dt = s.dt;
  if(dtmin < dt) dt = std::min(dtmin, dt / 2.0);
  if(dtmin > 2.0 * dt) dt = dt * 2.0;
  s.totaltime_next = s.totaltime + dt;
  const double odt[3] = {s.par.out_screen_dt, s.par.out_scalar_dt, s.par.out_h5data_dt};
  for(int j = 0; j < 3; ++j) {
    if(odt[j] > 0) {
      if(s.totaltime == s.t_out[j]) {
        if(s.dt_saved > 0) dt = s.dt_saved;
        s.t_out[j] += odt[j];
      }
      if(s.totaltime + dt > s.t_out[j] - 0.1 * odt[j]) {
        s.dt_saved = dt;
        s.totaltime_next = j == 0 ? s.t_out[j] : std::min(s.totaltime_next, s.t_out[j]);
      }
    }
  }
  dt = s.totaltime_next - s.totaltime;
}

void reduce(System & s, int op, double * out, int nout) {
  double o[3] = {0, 0, 0};
  switch(op) {
    case FSPH_GET_SET_ADAPTIVE_TIMESTEP:
      run_reduce(s, RED_MIN_DT, o);
      set_adaptive_timestep(s, std::min(1e24, o[0]));
      if(out && nout > 0) out[0] = s.dt;
      return;
    case FSPH_GET_COMPUTE_SMOOTHINGLENGTH:
      apply(s, 100);
      return;
    case FSPH_GET_COMPUTE_AVERAGE_SMOOTHINGLENGTH: {
      apply(s, 100);
      run_reduce(s, RED_SUM_H, o);
      // nparticles = body_system::getNBodies(): with periodic boundaries the reference's count includes the WALL copies
      // (bodies_system.h:166-168 re-counts after pboundary_generate), and so does the sum above
      double cnt[3] = {(double)s.n_owned_or_n(), 0, 0};
      if(s.comm) global_combine(s, RED_SUM_H, cnt);
      const double new_h = 1. / cnt[0] * o[0];
      const double keep = s.dt;
      s.dt = new_h;
      apply(s, 101);
      s.dt = keep;
      return;
    }
    case FSPH_GET_TOTAL_MASS: run_reduce(s, RED_MASS, o); break;
    case FSPH_GET_TOTAL_ENERGY: run_reduce(s, RED_ENERGY, o); break;
    case FSPH_GET_TOTAL_MOMENTUM: run_reduce(s, RED_MOMENTUM, o); break;
    case FSPH_GET_TOTAL_KINETIC_ENERGY: run_reduce(s, RED_EKIN, o); break;
    case FSPH_GET_TOTAL_INTERNAL_ENERGY: run_reduce(s, RED_EINT, o); break;
    case FSPH_GET_TOTAL_GRAVITATIONAL_ENERGY: run_reduce(s, RED_EGRAV, o); o[0] *= 0.5; break;
    case FSPH_GET_TOTAL_ANG_MOM: run_reduce(s, RED_ANGMOM, o); break;
    case FSPH_GET_SMOOTHINGLENGTH_STATS: run_reduce(s, RED_H_STATS, o); break;
    case FSPH_GET_VELOCITY_STATS: run_reduce(s, RED_V_STATS, o); break;
    case FSPH_GET_NEIGHBORS_STATS: run_reduce(s, RED_N_STATS, o); break;
    default: throw std::runtime_error("unknown reduce op");
  }
  for(int j = 0; j < nout && j < 3; ++j) out[j] = o[j];
}

void check_status(System & s) {
  uint32_t st = 0;
  uint32_t * d = s.b_status.as<uint32_t>();
  FSPH_CUDA_OK(cudaMemcpyAsync(&st, d, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(!st) return;
  FSPH_CUDA_OK(cudaMemsetAsync(d, 0, 4, s.stream));
  std::string m = "device guard failed:";
  if(st & ERR_DUP_KEYS) m += " duplicate keys;";
  if(st & ERR_RHO_NONPOS) m += " density of a particle is not a positive number (default_physics.h:215);";
  if(st & ERR_U_NONPOS) m += " internal energy non-positive (default_physics.h:134);";
  if(st & ERR_NAN) m += " NaN in a particle field (check_nans);";
  if(st & ERR_NEGATIVE) m += " negative density / pressure / internal energy (check_negativity);";
  if(st & ERR_QUEUE_OVERFLOW) m += " neighbour-search frontier overflow;";
  if(st & ERR_DT_LOOP) m += " eint-based dt estimator loop did not converge (default_physics.h:661);";
  if(st & ERR_RANGE) m += " a body lies outside the key range;";
  if(st & ERR_SEARCH_CAPACITY) m += " neighbour structure exceeded its reserved capacity;";
  if(st & ERR_DD_CAPACITY) m += " domain-decomposition buffer capacity / inconsistent global leaf depths;";
  throw std::runtime_error(m);
}

}  // namespace fsph
