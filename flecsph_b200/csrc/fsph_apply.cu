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
  return p;
}

struct OpList {
  int n;
  int op[8];
};

template<int D>
__device__ __forceinline__ void apply_one(const Cols & c, const int64_t i, const int op, const PhysParams & pp, uint32_t * __restrict__ status) {
  const double dt = pp.dt;
  switch(op) {
    case FSPH_APPLY_EOS_INIT:
      c.f[F_ADIA][i] = c.f[F_P][i] / pow(c.f[F_RHO][i], pp.gamma);
      break;
    case FSPH_APPLY_SET_TOTAL_ENERGY: {
      double v2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) v2 = v2 + c.f[F_VEL + d][i] * c.f[F_VEL + d][i];
      c.f[F_E][i] = .5 * v2 + c.f[F_U][i] + 0.0;
    } break;
    case FSPH_APPLY_SAVE_VELOCITYHALF:
#pragma unroll
      for(int d = 0; d < D; ++d) c.f[F_VELH + d][i] = c.f[F_VEL + d][i];
      break;
    case FSPH_APPLY_LEAPFROG_KICK_V:
#pragma unroll
      for(int d = 0; d < D; ++d) c.f[F_VEL + d][i] = c.f[F_VEL + d][i] + (0.5 * dt) * (c.f[F_ACC + d][i] + c.f[F_GACC + d][i]);
      break;
    case FSPH_APPLY_LEAPFROG_KICK_U: c.f[F_U][i] = c.f[F_U][i] + 0.5 * dt * c.f[F_DUDT][i]; break;
    case FSPH_APPLY_LEAPFROG_KICK_E: c.f[F_E][i] = c.f[F_E][i] + 0.5 * dt * c.f[F_DEDT][i]; break;
    case FSPH_APPLY_LEAPFROG_DRIFT:
#pragma unroll
      for(int d = 0; d < D; ++d) c.f[F_POS + d][i] = c.f[F_POS + d][i] + dt * c.f[F_VEL + d][i];
      break;
    case FSPH_APPLY_RECOMPUTE_P_CS: {
      const double u = c.f[F_U][i] + 0.5 * dt * c.f[F_DUDT][i];
      eos_p_cs(pp, c.f[F_RHO][i], u, c.f[F_ADIA][i], c.f[F_P][i], c.f[F_CS][i]);
    } break;
    case FSPH_APPLY_RECOMPUTE_P_CS_THERMOKINETIC: {
      double v2 = 0.0, va = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        v2 = v2 + c.f[F_VEL + d][i] * c.f[F_VEL + d][i];
        va = va + c.f[F_VEL + d][i] * c.f[F_ACC + d][i];
      }
      const double eint = c.f[F_E][i] - .5 * v2 - 0.0;
      if(!(eint > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
      c.f[F_U][i] = eint;
      const double u = eint + 0.5 * dt * (c.f[F_DEDT][i] - va);
      eos_p_cs(pp, c.f[F_RHO][i], u, c.f[F_ADIA][i], c.f[F_P][i], c.f[F_CS][i]);
    } break;
    case FSPH_APPLY_COMPUTE_DT: {
      const double tiny = 1e-24, mc = 0.6;
      const double dx = c.f[F_H][i] / (pp.eta * pp.kernel_width);
      double v2 = 0.0, a2 = 0.0, vg = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        const double v = c.f[F_VEL + d][i], a = c.f[F_ACC + d][i] + c.f[F_GACC + d][i];
        v2 = v2 + v * v;
        a2 = a2 + a * a;
        vg = vg + v * c.f[F_GACC + d][i];
      }
      const double vn = D == 1 ? fabs(c.f[F_VEL][i]) : sqrt(v2);
      const double acc = D == 1 ? fabs(c.f[F_ACC][i] + c.f[F_GACC][i]) : sqrt(a2);
      const double dt_v = dx / (vn + tiny);
      const double dt_a = sqrt(dx / (acc + tiny));
      const double dt_c = dx / (tiny + c.f[F_VSIG][i] * (1 + mc * pp.alpha));
      double dtmin = pp.cfl * fmin(fmin(dt_v, dt_a), dt_c);
      if(pp.evolve_u && pp.thermokinetic) {
        const double eint = c.f[F_U][i];
        if(!(eint > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
        int it;
        for(it = 0; it < 20; ++it) {
          const double delta_epot = 0.0, delta_egrv = -dtmin * vg;
          if(delta_epot + delta_egrv < eint * 0.5) break;
          dtmin *= 0.25;
        }
        if(it >= 20) atomicOr(status, (uint32_t)ERR_DT_LOOP);
      }
      c.f[F_DT][i] = dtmin;
    } break;
    case FSPH_APPLY_CHECK_NANS: {
      bool bad = false;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        bad = bad || isnan(c.f[F_POS + d][i]) || isnan(c.f[F_VEL + d][i]) || isnan(c.f[F_VELH + d][i]) ||
              isnan(c.f[F_ACC + d][i]) || isnan(c.f[F_GACC + d][i]);
      }
      bad = bad || isnan(c.f[F_MASS][i]) || isnan(c.f[F_GPOT][i]) || isnan(c.f[F_RHO][i]) || isnan(c.f[F_P][i]) ||
            isnan(c.f[F_E][i]) || isnan(c.f[F_DEDT][i]) || isnan(c.f[F_DUDT][i]) || isnan(c.f[F_ADIA][i]) || isnan(c.f[F_VSIG][i]);
      if(pp.evolve_u) bad = bad || isnan(c.f[F_U][i]);
      if(bad) atomicOr(status, (uint32_t)ERR_NAN);
    } break;
    case FSPH_APPLY_CHECK_NEGATIVITY:
      if(c.f[F_RHO][i] < 0 || c.f[F_P][i] < 0 || (pp.evolve_u && c.f[F_U][i] < 0)) atomicOr(status, (uint32_t)ERR_NEGATIVE);
      break;
    case 100: {  // compute_smoothinglength
      const double q = c.f[F_MASS][i] / c.f[F_RHO][i];
      const double r = D == 1 ? q : (D == 2 ? sqrt(q) : cbrt(q));
      c.f[F_H][i] = r * pp.eta * pp.kernel_width;
    } break;
    case 101: c.f[F_H][i] = pp.dt; break;  // uniform h (value carried in pp.dt)
  }
}

// Several element-wise functors of the driver loop in one pass over HBM (same per-particle order of operations).
template<int D>
__global__ void __launch_bounds__(256) k_apply(Cols c, int64_t n, OpList ops, PhysParams pp, uint32_t * __restrict__ status) {
  // n = number of owned bodies, starting at c.own_lo (remote copies are refreshed from their owner: reset_ghosts)
  const int64_t i = c.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= c.own_lo + n) return;
  for(int k = 0; k < ops.n; ++k) apply_one<D>(c, i, ops.op[k], pp, status);
}

void apply_fused(System & s, const int * op_list, int nops) {
  if(s.n <= 0 || nops <= 0) return;
  if(nops > 8) throw std::runtime_error("apply_fused: too many ops");
  const int64_t n_own = s.n_owned_or_n();
  unsigned nb = (unsigned)((n_own + 255) / 256);
  uint32_t * st = s.b_status.as<uint32_t>();
  // Measured on B200 (64 M particles): one kernel looping over a run-time op list is slower than one launch per
  // functor (the dependent load -> compute -> store chains of consecutive ops serialise on HBM latency), so the
  // list is issued as separate streaming passes; a compile-time fused pass is future work.
  for(int k = 0; k < nops; ++k) {
    OpList ops;
    ops.n = 1;
    for(int j = 0; j < 8; ++j) ops.op[j] = -1;
    ops.op[0] = op_list[k];
    if(s.dim == 1) FSPH_LAUNCH((k_apply<1>), nb, 256, s.stream, s.cols(), n_own, ops, s.phys(), st);
    else if(s.dim == 2) FSPH_LAUNCH((k_apply<2>), nb, 256, s.stream, s.cols(), n_own, ops, s.phys(), st);
    else FSPH_LAUNCH((k_apply<3>), nb, 256, s.stream, s.cols(), n_own, ops, s.phys(), st);
  }
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
__global__ void __launch_bounds__(256) k_reduce(Cols c, int64_t n, int kind, int thermokinetic, int fmm, double * __restrict__ part) {
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
            v[0] += m * (c.f[F_U][i] + .5 * v2);
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
  if(s.dim == 1) FSPH_LAUNCH((k_reduce<1>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, part);
  else if(s.dim == 2) FSPH_LAUNCH((k_reduce<2>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, part);
  else FSPH_LAUNCH((k_reduce<3>), nb, 256, s.stream, s.cols(), s.n, kind, s.par.thermokinetic_formulation, s.par.enable_fmm, part);
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
