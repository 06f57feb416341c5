// fsph_sweeps.cu -- the fp64 pair sweeps over the neighbour structure of fsph_search.cu.
//
// Replaces the functors the drivers pass to body_system::apply_in_smoothinglength
// (mpisph/bodies_system.h:304; include/physics/default_physics.h):
//   compute_density_pressure_soundspeed :309-320  (compute_density :193-228, recover_internal_energy :126-145,
//                                                  eos ideal eos/eos.h:166-183, compute_signalspeed :236-265)
//   compute_acceleration :348-408, compute_dudt :421-487, compute_dedt :502-566
// with the SPH kernels of include/physics/kernels.h (cubic :106-153, quintic :217-268, Wendland C2 :282-365,
// C4 :375-463, C6 :480-566, gaussian :168-203, super gaussian :582-612, sinc :626-699) and the constant artificial
// viscosity of include/physics/viscosity.h:81-137.
//
// One CTA of 256 threads per block of the neighbour structure (thread k <-> body k of the block; a block is a run of
// whole work cells, fsph_search.cu).  The block's union candidate list is staged in shared memory in tiles of 512 as
// SoA columns (coalesced gathers: candidates are runs of consecutive Morton indices); each warp then reads its
// neighbour-list rows of the tile (32 x uint16, coalesced) and every lane evaluates its own neighbours from shared
// memory, two pairs per iteration.  FP64-pipe bound; FMA contraction is allowed here (1e-10 bar).
//
// Pair caches.  x, h and m are frozen between the sweeps of an iteration, and so is everything the artificial viscosity
// reads between the acceleration and the energy sweep (rho, c_s, v_1/2).  When device memory allows, the density
// sweep therefore also stores A_ab = m_b [x_b != x_a] F(r_ab, h_ab) for every pair (grad W = r_ab F), the
// acceleration sweep reads A and stores Pi_ab, and the energy sweep reads both: its pair costs ~20 fp64 instructions
// instead of ~90, and the acceleration sweep skips the kernel gradient.  The caches are parallel to the list arena
// (one double per list slot) and are used only while the neighbour structure and the fields they were computed from
// are unchanged (System::pair_a_epoch / pair_pi_valid); without them every sweep recomputes (CM = 0).
#include <stdexcept>
#include <cfloat>
#include <cstdlib>
#include "fsph_internal.h"

namespace fsph {

constexpr double TINY = 1e10 * DBL_MIN;  // kernels.h:64, viscosity.h:36
constexpr double PI = 3.14159265358979323846;

enum SweepOp { OP_DENSITY = 0, OP_ACC = 1, OP_DUDT = 2, OP_DEDT = 3, OP_VSIG = 5, OP_ACCV = 6, OP_ALPHA = 7, OP_REPULSION = 8 };
// VISC (template flag): 0 = constant artificial viscosity, 1 = Cullen-Dehnen (per-body alpha; the density sweep also
// accumulates the velocity divergence)
// cache mode of a sweep.  density: 1 = write A.  acceleration: 1 = read A, write Pi.  energy: 1 = read A and Pi,
// 2 = read A only (Pi is stale, e.g. second pressure pass).  0 = no cache.
enum CacheMode { CM_NONE = 0, CM_FULL = 1, CM_A_ONLY = 2 };

// ---- SPH kernels: value Wi(r, 1/h) and the scalar Fi(r, 1/h, 1/r) with grad W = r_ab * Fi --------------------

template<int D> __device__ __forceinline__ double ipowD(double ih) { return D == 1 ? ih : (D == 2 ? ih * ih : ih * ih * ih); }
// Reciprocal and reciprocal square root from the MUFU seeds (~20 bits) plus two Newton steps (error ~1e-16
// relative, far inside the 1e-10 bar) -- several times cheaper than the IEEE division / sqrt sequences with
// their special-case branches.  Arguments are positive normal numbers here.
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
}
template<int D, int K> struct SphKernel;

template<int D> struct SphKernel<D, FSPH_KERNEL_CUBIC_SPLINE> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 4. / 3. : (D == 2 ? 40. / (7. * PI) : 8. / PI); }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = 2. * r * ih;
    double res = 0.;
    if(q < 2.0) {
      res = q < 1.0 ? 1.0 - 1.5 * q * q + .75 * q * q * q : 0.25 * (2 - q) * (2 - q) * (2 - q);
      res *= sigma() * ipowD<D>(ih);
    }
    return res;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 2. * r * ih;
    double res = 0.;
    if(q < 2.0) {
      const double sg = 2. * sigma() * ipowD<D>(ih) * ih;
      const double dW = q < 1.0 ? -3.0 * q + 9. / 4. * q * q : -.75 * (2 - q) * (2 - q);
      res = sg * dW * ir;
    }
    return res;
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_QUINTIC_SPLINE> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 1. / 40. : (D == 2 ? 63. / (478. * PI) : 81. / (359. * PI)); }
  static __device__ __forceinline__ double p5(double x) { double x2 = x * x; return x2 * x2 * x; }
  static __device__ __forceinline__ double p4(double x) { double x2 = x * x; return x2 * x2; }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = 3. * r * ih;
    double res = 0.;
    if(q < 3.) {
      res = p5(3 - q);
      if(q < 2.) res += -6 * p5(2 - q);
      if(q < 1.) res += 15 * p5(1 - q);
      res *= sigma() * ipowD<D>(ih);
    }
    return res;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 3. * r * ih;
    double res = 0.;
    if(q < 3.) {
      const double sg = 3. * sigma() * ipowD<D>(ih) * ih;
      double dW = -5. * p4(3 - q);
      if(q < 2.) dW += 30. * p4(2 - q);
      if(q < 1.) dW += -75. * p4(1 - q);
      res = sg * dW * ir;
    }
    return res;
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_WENDLAND_C2> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 1.25 : (D == 2 ? 7. / PI : 21. / (2. * PI)); }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double res = D == 1 ? sigma() * ih * o2 * o * (3 * q + 1) : sigma() * ipowD<D>(ih) * o2 * o2 * (4 * q + 1);
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double sg = D == 1 ? sigma() * ih * ih : 2. * sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -12. * q * o2 : -10. * q * o2 * o;
    return q < 1.0 ? sg * dW * ir : 0.0;
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_WENDLAND_C4> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 1.5 : (D == 2 ? 9. / PI : 495. / (32. * PI)); }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = r * ih, o = 1 - q, o2 = o * o;
    double res;
    if(D == 1) res = sigma() * ih * (o2 * o) * o2 * (1 + q * (5 + q * 8));
    else res = sigma() * ipowD<D>(ih) * (o2 * o2 * o2) * (1 + q * (6 + q * (35. / 3.)));
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double sg = 14. * sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -q * o2 * o2 * (1 + 4. * q) : -4. / 3. * q * (o2 * o) * o2 * (1 + 5. * q);
    return q < 1.0 ? sg * dW * ir : 0.0;
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_WENDLAND_C6> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 55. / 32. : (D == 2 ? 78. / (7. * PI) : 1365. / (64. * PI)); }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = r * ih, o = 1 - q, o2 = o * o, o4 = o2 * o2;
    double res;
    if(D == 1) res = sigma() * ih * (o2 * o) * o4 * (1 + q * (7 + q * (19 + q * 21)));
    else res = sigma() * ipowD<D>(ih) * o4 * o4 * (1 + q * (8 + q * (25 + q * 32)));
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = r * ih, o = 1 - q, o2 = o * o, o3 = o2 * o, o4 = o2 * o2;
    const double sg = sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -6. * q * o3 * o3 * (3 + q * (18 + q * 35)) : -22. * q * o4 * o3 * (1 + q * (7 + q * 16));
    return q < 1.0 ? sg * dW * ir : 0.0;
  }
};

// gaussian / super gaussian (kernels.h:168-203, 582-612): support 3 r / h < 3, h is the full support radius
template<int D> struct SphKernel<D, FSPH_KERNEL_GAUSSIAN> {
  static __device__ __forceinline__ double sigma() {
    return D == 1 ? 1.69260614115414981387661988500545775 : (D == 2 ? 2.86514256233641438778811727710055596 : 4.85098600188835377710867224691152783);
  }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = 3. * r * ih;
    return q <= 3. ? sigma() * ipowD<D>(ih) * exp(-q * q) : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 3. * r * ih;
    const double sg = 3. * sigma() * ipowD<D>(ih) * ih;
    return q < 3. ? sg * (-2. * q * exp(-q * q)) * ir : 0.0;
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_SUPER_GAUSSIAN> {
  static __device__ __forceinline__ double sigma() {
    return D == 1 ? 1.69225265632490690298438831093977954 : (D == 2 ? 2.86196342089351035329931251213987773 : 4.83280745997963041039733264826054666);
  }
  static __device__ __forceinline__ double Wi(double r, double ih) {
    const double q = 3. * r * ih, q2 = q * q;
    return q < 3. ? sigma() * ipowD<D>(ih) * exp(-q2) * (D / 2.0 + 1 - q2) : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 3. * r * ih;
    const double sg = 3. * sigma() * ipowD<D>(ih) * ih;
    const double dW = exp(-q * q) * (2. * q * q * q - (D + 4.) * q);
    return q < 3. ? sg * dW * ir : 0.0;
  }
};
// sinc kernel (kernels.h:626-699): index n and its normalisation (set_sinc_kernel_normalization, :67-77) are run-time
// parameters; they live in constant memory, refreshed by sweep() before the launches
static __constant__ double c_sinc[2];  // {n, sigma(n, D)}
template<int D> struct SphKernel<D, FSPH_KERNEL_SINC> {
  static __device__ __forceinline__ double Wi(double r, double ih) {
    double q = fabs(r * ih);
    if(!(q < 1.)) return 0.0;
    q *= PI;
    double sinc = 1.0;
    if(q > 1e-6) sinc = sin(q) / q;
    else if(q > 1e-24) { const double q2 = q * q; sinc += q2 * (.05 * q2 - 1.) / 6.; }
    return c_sinc[1] * ipowD<D>(ih) * pow(sinc, c_sinc[0]);
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    double q = fabs(r * ih);
    if(!(q < 1.)) return 0.0;
    const double n = c_sinc[0], sg = c_sinc[1] * ipowD<D>(ih) * ih;
    q *= PI;
    double dW = 0.0;
    if(q > 1e-6) {
      const double sinc = sin(q) / q, cosx = cos(q) / q;
      dW = -n * pow(sinc, n - 1) * PI * (sinc / q - cosx);
    }
    else if(q > 1e-24) {
      const double q2 = q * q;
      dW = -n * q / 3. * PI * (1 - .5 * q2 * (.2 - (n - 1) / 18));
    }
    return sg * dW * ir;
  }
};


// ---- candidate tile columns per sweep ---------------------------------------------------------------------
// which neighbour fields a sweep gathers into shared memory; positions come first, the rest in a fixed order
#ifndef FSPH_SW_ROWS_SMEM
#define FSPH_SW_ROWS_SMEM 1
#endif
constexpr int ROW_STAGE = 48;   // list rows per (tile, warp) staged in shared memory (longer lists are read from global memory)
template<int D, int OP, int CM, int VISC> struct TileCols {
  static constexpr bool ACCX = OP == OP_ACC || OP == OP_ACCV;
  static constexpr bool ENER = OP == OP_DUDT || OP == OP_DEDT;
  static constexpr bool PIV = (ACCX) || (ENER && CM != CM_FULL);         // the sweep evaluates Pi itself
  static constexpr bool GRAD = ((ACCX || ENER) && CM == CM_NONE) || OP == OP_ALPHA;   // ... the kernel gradient itself
  static constexpr bool hasH = OP == OP_DENSITY || PIV || GRAD || OP == OP_REPULSION;
  static constexpr bool hasM = OP == OP_DENSITY || GRAD || OP == OP_REPULSION;
  static constexpr bool hasRHO = PIV;
  static constexpr bool hasPR = ACCX || OP == OP_DEDT;
  static constexpr bool hasCS = PIV || OP == OP_VSIG || OP == OP_ALPHA;
  static constexpr bool hasVH = PIV;
  static constexpr bool hasV = ENER || OP == OP_VSIG || OP == OP_ALPHA;
  static constexpr bool hasCSO = OP == OP_VSIG || OP == OP_ACCV;
  static constexpr bool hasAL = PIV && VISC == 1;                        // alpha_b of the Cullen-Dehnen viscosity
  static constexpr bool hasDV = OP == OP_ALPHA;                          // div v of the neighbour
  static constexpr bool hasIDX = OP == OP_VSIG || OP == OP_ACCV || OP == OP_REPULSION;
  static constexpr int iH = D;
  static constexpr int iM = iH + (hasH ? 1 : 0);
  static constexpr int iRHO = iM + (hasM ? 1 : 0);
  static constexpr int iPR = iRHO + (hasRHO ? 1 : 0);
  static constexpr int iCS = iPR + (hasPR ? 1 : 0);
  static constexpr int iVH = iCS + (hasCS ? 1 : 0);
  static constexpr int iV = iVH + (hasVH ? D : 0);
  static constexpr int iCSO = iV + (hasV ? D : 0);
  static constexpr int iAL = iCSO + (hasCSO ? 1 : 0);
  static constexpr int iDV = iAL + (hasAL ? 1 : 0);
  static constexpr int NF = iDV + (hasDV ? 1 : 0);
  static __device__ __forceinline__ int col(int j) {
    if(j < D) return F_POS + j;
    if(hasH && j == iH) return F_H;
    if(hasM && j == iM) return F_MASS;
    if(hasRHO && j == iRHO) return F_RHO;
    if(hasPR && j == iPR) return F_PRHO2;   // P / rho^2, refreshed by k_prho2 right before the sweep
    if(hasCS && j == iCS) return F_CS;
    if(hasVH && j >= iVH && j < iVH + D) return F_VELH + (j - iVH);
    if(hasV && j >= iV && j < iV + D) return F_VEL + (j - iV);
    if(hasCSO && j == iCSO) return F_CSOLD;
    if(hasAL && j == iAL) return F_ALPHA;
    return F_DIVV;
  }
  __host__ __device__ static constexpr size_t tile_bytes() { return (size_t)NF * NL_TILE * 8 + (hasIDX ? NL_TILE * 4 : 0); }
  // The warps' list rows of a tile are fetched by TMA bulk copies into shared memory (ROW_STAGE rows + the count row per
  // warp) while the tile's candidate columns are being gathered: measured to pay for the density and the energy sweeps
  // (density 22.1 -> 21.1 ms at 64 M); the acceleration sweep, at its register limit, loses (39.7 -> 49.6 ms) and keeps
  // reading its rows from global memory.
  static constexpr bool ROWS_SMEM = FSPH_SW_ROWS_SMEM && !ACCX && OP != OP_ALPHA &&
                                    tile_bytes() + (size_t)(BLK_BODIES / 32) * (ROW_STAGE + 1) * 64 + 64 <= 74 * 1024;
  static constexpr size_t smem_bytes() { return tile_bytes() + (ROWS_SMEM ? (size_t)(BLK_BODIES / 32) * (ROW_STAGE + 1) * 64 + 64 : 0); }
};

// state of the particle a thread owns + its accumulators
template<int D> struct Own {
  double xa[D], va[D], vha[D];
  double ha, rho_a, cs_a, Prho2_a, alpha_a;
  double acc0, acc1, accv[D], vsig;
  double gv[D * D];   // compute_alpha only: velocity-gradient sums
  int ia;
};

// Pi_ab of the constant artificial viscosity (viscosity.h:81-137): mu = h dot / (r^2 + eps h^2 + TINY);
// Pi = (-alpha c + beta mu) mu / rho -- one division for both
// Cullen-Dehnen (viscosity.h:150-157): Pi = -alpha_ab (c - 2 mu) mu / rho with the bodies' own alphas
template<int VISC>
__device__ __forceinline__ double visc_pi(const PhysParams & pp, double dotvr, double dist2, double hab, double rho_ab, double c_ab,
  double alpha_ab) {
  double Pi = 0.0;
  if(dotvr < 0.0) {
    const double den = dist2 + pp.epsilon * hab * hab + TINY;
    const double inv = rcp_nr(den * rho_ab);
    const double mu = hab * dotvr * (inv * rho_ab);
    if(VISC == 1) Pi = -alpha_ab * (c_ab - 2.0 * mu) * mu * (inv * den);
    else Pi = (-pp.alpha * c_ab + pp.beta * mu) * mu * (inv * den);
  }
  return Pi;
}

// one neighbour pair (a, candidate ci of the current tile); `valid` = false makes it a no-op (padding slot of the
// two-pairs-per-iteration loop).  cA / cP: the pair's cache slots (read or written according to CM).
template<int D, int K, int OP, int CM, int VISC>
__device__ __forceinline__ void pair_body(Own<D> & o, const double (*tile)[NL_TILE], const int32_t * tidx, const PhysParams & pp,
  const int ci, const bool valid, double & cA, double & cP) {
  using TC = TileCols<D, OP, CM, VISC>;
  double rab[D];
  double dist2 = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    rab[d] = o.xa[d] - tile[d][ci];
    dist2 += rab[d] * rab[d];
  }
  if(OP == OP_DENSITY) {
    // r, 1/r and 1/h_ab: one reciprocal square root + one reciprocal (dist2 clamped away from 0: W(0) is finite and the
    // gradient factor of the particle itself / coincident particles is dropped through m_b [x_b != x_a] = 0)
    const double ir = rsqrt_nr(fmax(dist2, 1e-280));
    const double r = D == 1 ? fabs(rab[0]) : dist2 * ir;
    const double hab = .5 * (o.ha + tile[TC::iH][ci]);
    const double ih = rcp_nr(hab);
    const double mb = tile[TC::iM][ci];
    const double t = mb * SphKernel<D, K>::Wi(r, ih);
    o.acc0 += valid ? t : 0.0;
    if(CM == CM_FULL || VISC == 1) {
      const double A = (dist2 == 0.0 || !valid) ? 0.0 : mb * SphKernel<D, K>::Fi(r, ih, ir);
      if(CM == CM_FULL) cA = A;
      if(VISC == 1) {
        // compute_divv (default_physics.h:273-300): sum_b m_b grad W_ab, contracted with v_a at the end
#pragma unroll
        for(int d = 0; d < D; ++d) o.accv[d] += A * rab[d];
      }
    }
  }
  else if(OP == OP_ALPHA) {
    // viscosity::compute_alpha (viscosity.h:310-363) with compute_xi (:193-262): signal velocity, R_a and the
    // velocity-gradient matrix
    const double hab = .5 * (o.ha + tile[TC::iH][ci]);
    const double ir = rsqrt_nr(fmax(dist2, 1e-280));
    const double r = D == 1 ? fabs(rab[0]) : dist2 * ir;
    const double ih = rcp_nr(hab);
    const double mb = valid ? tile[TC::iM][ci] : 0.0;
    const double Wab = SphKernel<D, K>::Wi(r, ih);
    const double Fab = dist2 == 0.0 ? 0.0 : SphKernel<D, K>::Fi(r, ih, ir);
    const double dvb = tile[TC::iDV][ci];
    const double sg = dvb > 0.0 ? 1.0 : (dvb < 0.0 ? -1.0 : dvb);   // signnum_c
    o.acc0 += sg * mb * Wab;
    double vab[D], vr = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      vab[d] = o.va[d] - tile[TC::iV + d][ci];
      vr += vab[d] * rab[d];
    }
#pragma unroll
    for(int i = 0; i < D; ++i)
#pragma unroll
      for(int j = 0; j < D; ++j) o.gv[D * i + j] += mb * (-vab[i]) * (rab[j] * Fab);
    const double cab = fmax(o.cs_a, tile[TC::iCS][ci]);
    vr = vr * ir;
    const double mn = (0.0 < vr) ? 0.0 : vr;
    const double val = cab - mn;
    if(valid && dist2 != 0.0 && o.vsig < val) o.vsig = val;   // the body itself contributes NaN in the reference: dropped
  }
  else if(OP == OP_REPULSION) {
    // add_short_range_repulsion (default_physics.h:718-754): the body itself is skipped by id
    const double hab = .5 * (o.ha + tile[TC::iH][ci]);
    const double r = D == 1 ? fabs(rab[0]) : sqrt(dist2);
    if(valid && tidx[ci] != o.ia && !(r > hab * pp.repulsion_radius)) {
      const double f = tile[TC::iM][ci] / (r * r * r);
#pragma unroll
      for(int d = 0; d < D; ++d) o.accv[d] += f * rab[d];
    }
  }
  else if(OP == OP_VSIG) {
    // compute_signalspeed: b contributes its new soundspeed iff it was processed before a, i.e. iff its
    // (descending-key processing order) index is larger
    const double csb = tidx[ci] > o.ia ? tile[TC::iCS][ci] : tile[TC::iCSO][ci];
    const double cab = fmax(o.cs_a, csb);
    const double dist = D == 1 ? fabs(rab[0]) : sqrt(dist2);
    double vr = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) vr += (o.va[d] - tile[TC::iV + d][ci]) * rab[d];
    vr = vr / dist;                            // 0/0 = NaN for the particle itself
    const double mn = (0.0 < vr) ? 0.0 : vr;  // std::min(vr, 0.0), NaN-preserving
    const double val = cab - mn;
    if(valid && o.vsig < val) o.vsig = val;   // std::max(vsig, val): NaN (self) dropped
  }
  else {
    // ---- A = m_b [x_b != x_a] F(r, h_ab) and Pi_ab: computed here or taken from the caches ---------------------
    double A, Pi, ir = 0.0;
    if(TC::GRAD) {
      const double hab = .5 * (o.ha + tile[TC::iH][ci]);
      ir = rsqrt_nr(fmax(dist2, 1e-280));
      const double r = D == 1 ? fabs(rab[0]) : dist2 * ir;
      const double mb = (dist2 == 0.0 || !valid) ? 0.0 : tile[TC::iM][ci];   // m_b * (pos_b != pos_a)
      A = mb * SphKernel<D, K>::Fi(r, rcp_nr(hab), ir);
    }
    else A = valid ? cA : 0.0;
    if(TC::PIV) {
      const double hab = .5 * (o.ha + tile[TC::iH][ci]);
      double dotvr = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) dotvr += (o.vha[d] - tile[TC::iVH + d][ci]) * rab[d];
      const double alpha_ab = VISC == 1 ? .5 * (o.alpha_a + tile[TC::hasAL ? TC::iAL : 0][ci]) : 0.0;
      Pi = visc_pi<VISC>(pp, dotvr, dist2, hab, .5 * (o.rho_a + tile[TC::iRHO][ci]), .5 * (o.cs_a + tile[TC::iCS][ci]), alpha_ab);
      if(TC::ACCX && CM == CM_FULL) cP = Pi;
    }
    else Pi = cP;
    if(TC::ACCX) {
      const double coef = -(o.Prho2_a + tile[TC::iPR][ci] + Pi) * A;
#pragma unroll
      for(int d = 0; d < D; ++d) o.accv[d] += coef * rab[d];
      if(OP == OP_ACCV) {
        // compute_signalspeed fused into the acceleration sweep (fsph_step only: v == v_1/2 there, and all new
        // soundspeeds exist by now)
        if(!TC::GRAD) ir = rsqrt_nr(fmax(dist2, 1e-280));
        const double csb = tidx[ci] > o.ia ? tile[TC::iCS][ci] : tile[TC::iCSO][ci];
        const double cab = fmax(o.cs_a, csb);
        double vr = 0.0;
#pragma unroll
        for(int d = 0; d < D; ++d) vr += (o.vha[d] - tile[TC::iVH + d][ci]) * rab[d];
        vr = vr * ir;                              // dist2 == 0 (self) is excluded below instead of via NaN
        const double mn = (0.0 < vr) ? 0.0 : vr;
        const double val = cab - mn;
        if(valid && dist2 != 0.0 && o.vsig < val) o.vsig = val;
      }
    }
    else if(OP == OP_DUDT) {
      double vdw = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) vdw += (o.va[d] - tile[TC::iV + d][ci]) * rab[d];
      vdw *= A;
      o.acc0 += vdw;
      o.acc1 += vdw * Pi;
    }
    else {  // OP_DEDT
      double vadw = 0.0, vbdw = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        vadw += o.va[d] * rab[d];
        vbdw += tile[TC::iV + d][ci] * rab[d];
      }
      vadw *= A;
      vbdw *= A;
      o.acc0 -= o.Prho2_a * vbdw + vadw * tile[TC::iPR][ci] + .5 * Pi * (vbdw + vadw);
    }
  }
}

// state of the particle `ia` of a sweep (the columns op OP needs), accumulators cleared
template<int D, int OP, int VISC>
__device__ __forceinline__ void load_own(Own<D> & o, const Cols & c, const int ia, const bool active) {
  o.ia = ia;
  o.ha = o.rho_a = o.cs_a = o.Prho2_a = o.alpha_a = 0.0;
  o.acc0 = o.acc1 = o.vsig = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) { o.xa[d] = 0; o.va[d] = 0; o.vha[d] = 0; o.accv[d] = 0; }
#pragma unroll
  for(int d = 0; d < D * D; ++d) o.gv[d] = 0;
  if(active) {
#pragma unroll
    for(int d = 0; d < D; ++d) o.xa[d] = c.f[F_POS + d][ia];
    o.ha = c.f[F_H][ia];
    if(OP == OP_VSIG || OP == OP_DUDT || OP == OP_DEDT || OP == OP_ALPHA || (OP == OP_DENSITY && VISC == 1)) {
#pragma unroll
      for(int d = 0; d < D; ++d) o.va[d] = c.f[F_VEL + d][ia];
    }
    if(OP == OP_VSIG || OP == OP_ALPHA) o.cs_a = c.f[F_CS][ia];
    if(OP == OP_ALPHA) o.rho_a = c.f[F_RHO][ia];
    if(VISC == 1 && (OP == OP_ACC || OP == OP_DUDT || OP == OP_DEDT || OP == OP_ACCV)) o.alpha_a = c.f[F_ALPHA][ia];
    if(OP == OP_ACC || OP == OP_DUDT || OP == OP_DEDT || OP == OP_ACCV) {
      o.rho_a = c.f[F_RHO][ia];
      o.cs_a = c.f[F_CS][ia];
      o.Prho2_a = c.f[F_PRHO2][ia];
#pragma unroll
      for(int d = 0; d < D; ++d) o.vha[d] = c.f[F_VELH + d][ia];
    }
  }
}

// the per-particle tail of each functor: writes the results of the sweep
template<int D, int OP, int VISC>
__device__ __forceinline__ void store_own(const Own<D> & o, const Cols & c, const PhysParams & pp, uint32_t * __restrict__ status) {
  const int ia = o.ia;
  if(OP == OP_DENSITY) {
    const double rho = o.acc0;
    if(!(rho > 0)) atomicOr(status, (uint32_t)ERR_RHO_NONPOS);
    c.f[F_RHO][ia] = rho;
    double u;
    if(pp.evolve_u && pp.thermokinetic && !pp.ext_potential) {
      // recover_internal_energy without an external potential (with one: element-wise pass 103 before this sweep)
      double v2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) { const double v = c.f[F_VEL + d][ia]; v2 += v * v; }
      u = c.f[F_E][ia] - .5 * v2 - 0.0;
      if(!(u > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
      c.f[F_U][ia] = u;
    }
    else u = c.f[F_U][ia];
    double Pn = c.f[F_P][ia], csn = c.f[F_CS][ia];
    eos_p_cs(pp, rho, u, (pp.eos == FSPH_EOS_POLYTROPIC || pp.eos == FSPH_EOS_PPT) ? c.f[F_ADIA][ia] : 0.0,
      pp.eos == FSPH_EOS_WD ? c.f[F_YE][ia] : 0.0, Pn, csn);
    c.f[F_P][ia] = Pn;
    c.f[F_CSOLD][ia] = c.f[F_CS][ia];
    c.f[F_CS][ia] = csn;
    if(VISC == 1) {
      // compute_divv (default_physics.h:273-300): div v = (v_a . sum_b m_b grad W_ab) / rho_a and its rate of change
      double dv = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) dv += o.va[d] * o.accv[d];
      dv = dv / rho;
      c.f[F_DDIVVDT][ia] = (dv - c.f[F_DIVV][ia]) / pp.dt;
      c.f[F_DIVV][ia] = dv;
    }
  }
  else if(OP == OP_ALPHA) {
    const double divV = c.f[F_DIVV][ia], h_a = o.ha, alpha_a = c.f[F_ALPHA][ia];
    const double R_a = o.acc0 / o.rho_a;
    double g[D * D];
#pragma unroll
    for(int k = 0; k < D * D; ++k) g[k] = o.gv[k] / o.rho_a;
    c.f[F_GRADV][ia] = g[0];
    // traceless symmetric part of the velocity gradient
    double traceSS = 0.0;
#pragma unroll
    for(int i = 0; i < D; ++i)
#pragma unroll
      for(int j = 0; j < D; ++j) {
        double sym = 0.5 * (g[D * i + j] + g[D * j + i]);
        if(i == j) sym -= divV / D;
        traceSS += sym * sym;
      }
    c.f[F_TRACESS][ia] = traceSS;
    const double omr = 1.0 - R_a, q4 = omr * omr * omr * omr, t2 = 2.0 * q4 * divV, res = t2 * t2;
    const double xi = res < 1e-16 ? 0.0 : res / (res + traceSS);
    const double trig = fmax(-c.f[F_DDIVVDT][ia], 0.0);
    c.f[F_XI][ia] = xi;
    c.f[F_TRIGGER][ia] = trig;
    const double Atrig = xi * trig, vsig = o.vsig, vh = vsig / h_a;
    const double alpha_loc = pp.alpha_max * Atrig / (pp.visc_delta * (vh * vh) + Atrig);
    if(alpha_a <= alpha_loc) c.f[F_ALPHA][ia] = alpha_loc;
    else {
      const double decayt = h_a / (2.0 * pp.visc_l * vsig);
      const double dalphadt = (alpha_a - alpha_loc) / decayt;
      c.f[F_ALPHA][ia] = alpha_a * exp(-dalphadt * pp.dt);
    }
  }
  else if(OP == OP_REPULSION) {
#pragma unroll
    for(int d = 0; d < D; ++d) c.f[F_ACC + d][ia] = c.f[F_ACC + d][ia] + o.accv[d] * pp.repulsion_gamma;
  }
  else if(OP == OP_VSIG) c.f[F_VSIG][ia] = o.vsig;
  else if(OP == OP_ACC || OP == OP_ACCV) {
#pragma unroll
    for(int d = 0; d < D; ++d) {
      c.f[F_ACC + d][ia] = o.accv[d];  // external_force::acceleration is added by an element-wise pass right after (do_sweep)
      c.f[F_GACC + d][ia] = 0.0;
    }
    c.f[F_GPOT][ia] = 0.0;
    if(OP == OP_ACCV) {
      c.f[F_VSIG][ia] = o.vsig;
      if(pp.fuse_kick_v) {
        // leapfrog_kick_v (integration.h:71-79) right here: v += dt/2 (a + g) with g = 0, operation by operation as the
        // element-wise pass does it (no sweep of this step reads v of another body before the ghost refresh that follows)
#pragma unroll
        for(int d = 0; d < D; ++d)
          c.f[F_VEL + d][ia] = __dadd_rn(c.f[F_VEL + d][ia], __dmul_rn(0.5 * pp.dt, __dadd_rn(o.accv[d], 0.0)));
      }
    }
  }
  else if(OP == OP_DUDT) c.f[F_DUDT][ia] = o.Prho2_a * o.acc0 + .5 * o.acc1;
  else {
    double gv = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) gv += c.f[F_GACC + d][ia] * o.va[d];
    c.f[F_DEDT][ia] = o.acc0 + gv;
  }
}

template<int OP> struct SweepOcc { static constexpr int MINB = (OP == OP_DENSITY || OP == OP_VSIG || OP == OP_REPULSION) ? 4 : 3; };

template<int D, int K, int OP, int CM, int VISC>
__global__ void __launch_bounds__(BLK_BODIES, SweepOcc<OP>::MINB) k_sweep(Cols c, Search sr, PhysParams pp, uint32_t * __restrict__ status,
  double * __restrict__ pair_a, double * __restrict__ pair_pi) {
  using TC = TileCols<D, OP, CM, VISC>;
  constexpr int NF = TC::NF;
  constexpr int NT = BLK_BODIES;
  constexpr bool A_RD = (OP == OP_ACC || OP == OP_ACCV || OP == OP_DUDT || OP == OP_DEDT) && CM != CM_NONE;
  constexpr bool A_WR = OP == OP_DENSITY && CM == CM_FULL;
  constexpr bool P_RD = (OP == OP_DUDT || OP == OP_DEDT) && CM == CM_FULL;
  constexpr bool P_WR = (OP == OP_ACC || OP == OP_ACCV) && CM == CM_FULL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double (* const tile)[NL_TILE] = reinterpret_cast<double (*)[NL_TILE]>(smem_raw);
  int32_t * const tidx = reinterpret_cast<int32_t *>(smem_raw + (size_t)NF * NL_TILE * 8);

  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int body0 = sr.blk_body[blk], PB = sr.blk_nbody[blk], C = sr.blk_ncand[blk];
  const uint64_t co = sr.blk_cand_off[blk], dir0 = sr.blk_dir_off[blk];
  const bool active = tid < PB;
  const int nwarp_act = (PB + 31) >> 5;

  Own<D> o;
  load_own<D, OP, VISC>(o, c, body0 + tid, active);

  if(TC::ROWS_SMEM) {
    // ---- variant with the list rows staged in shared memory by TMA -------------------------------------------------
    uint16_t * const lrow = reinterpret_cast<uint16_t *>(smem_raw + TC::tile_bytes()) + (size_t)warp * (ROW_STAGE + 1) * 32;
    const uint32_t lrow_s = (uint32_t)__cvta_generic_to_shared(lrow);
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(smem_raw + TC::tile_bytes() + (size_t)NT / 32 * (ROW_STAGE + 1) * 64) + warp * 8;
    uint32_t phase = 0;
    if(lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int ntile = nl_ntiles(C);
    for(int ti = 0; ti < ntile; ++ti) {
      int tb, tc;
      nl_tile_range(C, ti, tb, tc);
      const uint64_t rb = warp < nwarp_act ? sr.dir[dir0 + ti].row[warp] : 0u;
      const int cmax = warp < nwarp_act ? (int)sr.dir[dir0 + ti].cmax[warp] : 0;
      const bool staged = cmax <= ROW_STAGE && warp < nwarp_act;
      __syncthreads();   // the previous tile's columns and staged rows are no longer read
      if(staged && lane == 0) {
        // the warp's 1 + cmax rows are contiguous: one bulk copy, completion signalled on the warp's mbarrier
        const uint32_t bytes = (uint32_t)(1 + cmax) * 64u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(lrow_s), "l"(sr.nl + rb * 32), "r"(bytes), "r"(mbar_s) : "memory");
      }
      {
        constexpr int LPT = NL_TILE / NT;
        int pidx[LPT];
#pragma unroll
        for(int u = 0; u < LPT; ++u) {
          const int i = tid + u * NT;
          pidx[u] = i < tc ? sr.cand[co + tb + i] : -1;
        }
#pragma unroll
        for(int u = 0; u < LPT; ++u) {
          const int i = tid + u * NT;
          if(pidx[u] >= 0) {
            if(TC::hasIDX) tidx[i] = pidx[u];
#pragma unroll
            for(int j = 0; j < NF; ++j) tile[j][i] = c.f[TC::col(j)][pidx[u]];
          }
        }
      }
      __syncthreads();
      const uint16_t * list = sr.nl + rb * 32 + lane;   // lists longer than the staging area: straight from global memory
      if(staged) {
        uint32_t done = 0;
        while(!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(mbar_s), "r"(phase) : "memory");
        phase ^= 1u;
        list = lrow + lane;
      }
      const int cnt = warp < nwarp_act ? list[0] : 0;
      int e0 = cmax > 0 ? list[32] : 0, e1 = cmax > 1 ? list[64] : 0;
      for(int j = 0; j < cmax; j += 2) {
        const bool v0 = j < cnt, v1 = j + 1 < cnt;
        const int ci0 = v0 ? e0 : 0, ci1 = v1 ? e1 : 0;
        const size_t slot0 = (size_t)(rb + 1 + j) * 32 + lane, slot1 = slot0 + 32;
        double a0 = 0.0, a1 = 0.0, p0 = 0.0, p1 = 0.0;
        if(A_RD) { a0 = v0 ? pair_a[slot0] : 0.0; a1 = v1 ? pair_a[slot1] : 0.0; }
        if(P_RD) { p0 = v0 ? pair_pi[slot0] : 0.0; p1 = v1 ? pair_pi[slot1] : 0.0; }
        if(j + 2 < cmax) e0 = list[(3 + j) * 32];
        if(j + 3 < cmax) e1 = list[(4 + j) * 32];
        pair_body<D, K, OP, CM, VISC>(o, tile, tidx, pp, ci0, v0, a0, p0);
        pair_body<D, K, OP, CM, VISC>(o, tile, tidx, pp, ci1, v1, a1, p1);
        if(A_WR) { if(v0) pair_a[slot0] = a0; if(v1) pair_a[slot1] = a1; }
        if(P_WR) { if(v0) pair_pi[slot0] = p0; if(v1) pair_pi[slot1] = p1; }
      }
    }
  }
  else {
    const int ntile = nl_ntiles(C);
    for(int ti = 0; ti < ntile; ++ti) {
      // tile ti starts at candidate tb of the block and holds tc of them
      int tb, tc;
      nl_tile_range(C, ti, tb, tc);
      // this warp's rows of the tile
      const uint64_t rb = warp < nwarp_act ? sr.dir[dir0 + ti].row[warp] : 0u;
      const uint16_t * rows = sr.nl + rb * 32 + lane;
      const int cnt = warp < nwarp_act ? rows[0] : 0;
      const int cmax = __reduce_max_sync(0xffffffffu, cnt);
      __syncthreads();
      {
        // all candidate indices of this thread first, then the column gathers: 2 x NF loads in flight
        constexpr int LPT = NL_TILE / NT;
        int pidx[LPT];
  #pragma unroll
        for(int u = 0; u < LPT; ++u) {
          const int i = tid + u * NT;
          pidx[u] = i < tc ? sr.cand[co + tb + i] : -1;
        }
  #pragma unroll
        for(int u = 0; u < LPT; ++u) {
          const int i = tid + u * NT;
          if(pidx[u] >= 0) {
            if(TC::hasIDX) tidx[i] = pidx[u];
  #pragma unroll
            for(int j = 0; j < NF; ++j) tile[j][i] = c.f[TC::col(j)][pidx[u]];
          }
        }
      }
      __syncthreads();
      // uniform pair loop over the lanes' lists, two pairs per iteration (independent fp64 dependency chains); the
      // entries of the next iteration are loaded before this one is evaluated
      int e0 = cmax > 0 ? rows[32] : 0, e1 = cmax > 1 ? rows[64] : 0;
      for(int j = 0; j < cmax; j += 2) {
        const bool v0 = j < cnt, v1 = j + 1 < cnt;
        const int ci0 = v0 ? e0 : 0, ci1 = v1 ? e1 : 0;
        const size_t slot0 = (size_t)(rb + 1 + j) * 32 + lane, slot1 = slot0 + 32;
        double a0 = 0.0, a1 = 0.0, p0 = 0.0, p1 = 0.0;
        if(A_RD) { a0 = v0 ? pair_a[slot0] : 0.0; a1 = v1 ? pair_a[slot1] : 0.0; }
        if(P_RD) { p0 = v0 ? pair_pi[slot0] : 0.0; p1 = v1 ? pair_pi[slot1] : 0.0; }
        if(j + 2 < cmax) e0 = rows[(3 + j) * 32];
        if(j + 3 < cmax) e1 = rows[(4 + j) * 32];
        pair_body<D, K, OP, CM, VISC>(o, tile, tidx, pp, ci0, v0, a0, p0);
        pair_body<D, K, OP, CM, VISC>(o, tile, tidx, pp, ci1, v1, a1, p1);
        if(A_WR) { if(v0) pair_a[slot0] = a0; if(v1) pair_a[slot1] = a1; }
        if(P_WR) { if(v0) pair_pi[slot0] = p0; if(v1) pair_pi[slot1] = p1; }
      }
    }
  }
  if(active) store_own<D, OP, VISC>(o, c, pp, status);
}

static __global__ void __launch_bounds__(256) k_prho2(Cols c, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) {
    const double rho = c.f[F_RHO][i];
    c.f[F_PRHO2][i] = c.f[F_P][i] / (rho * rho);
  }
}

static __global__ void __launch_bounds__(256) k_fill(double * p, int64_t n, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) p[i] = v;
}

template<int D, int K, int OP, int CM, int VISC = 0>
static void launch_sweep(System & s) {
  const Search & sr = s.search;
  if(OP == OP_ACC || OP == OP_ACCV || OP == OP_DUDT || OP == OP_DEDT)
    FSPH_LAUNCH(k_prho2, (unsigned)((s.n + 255) / 256), 256, s.stream, s.cols(), s.n);
  uint32_t * st = s.b_status.as<uint32_t>();
  FSPH_LAUNCH_SMEM((k_sweep<D, K, OP, CM, VISC>), sr.nblk, BLK_BODIES, (TileCols<D, OP, CM, VISC>::smem_bytes()), s.stream, s.cols(), sr,
    s.phys(), st, s.b_pair_a.as<double>(), s.b_pair_pi.as<double>());
}
// the Cullen-Dehnen variants are instantiated for every kernel as well: the viscosity is a run-time parameter
template<int D, int K, int OP, int CM>
static void launch_sweep_v(System & s) {
  if(s.par.sph_viscosity == FSPH_VISC_CULLEN) launch_sweep<D, K, OP, CM, 1>(s);
  else launch_sweep<D, K, OP, CM, 0>(s);
}

// Pair caches: decide once per neighbour structure whether they fit (two doubles per list slot next to everything else
// the step needs); FSPH_PAIR_CACHE=0/1 overrides (tests run both ways).
static bool pair_cache_ready(System & s) {
  if(s.pair_cache_epoch == s.search_epoch) return s.pair_cache_on;
  s.pair_cache_epoch = s.search_epoch;
  s.pair_a_epoch = 0;
  s.pair_pi_valid = false;
  int want = -1;
  if(const char * e = getenv("FSPH_PAIR_CACHE")) want = atoi(e);
  const size_t need = (size_t)(s.search.nrows + 2) * 32 * 8;
  bool on = want != 0;
  if(on && (s.b_pair_a.cap < need || s.b_pair_pi.cap < need)) {
    size_t fr = 0, tot = 0;
    FSPH_CUDA_OK(cudaMemGetInfo(&fr, &tot));
    const size_t have = s.b_pair_a.cap + s.b_pair_pi.cap;
    // keep 12 % of the device (and at least 2 GB) free for the buffers that still grow during a run
    const size_t reserve = std::max<size_t>(tot / 8, (size_t)2 << 30);
    if(want < 0 && fr + have < 2 * (need + need / 8) + reserve) on = false;
    if(on) {
      try {
        s.b_pair_a.reserve(need);
        s.b_pair_pi.reserve(need);
      }
      catch(const std::exception &) {
        cudaGetLastError();
        s.b_pair_a.release();
        s.b_pair_pi.release();
        on = false;
      }
    }
  }
  if(!on) { s.b_pair_a.release(); s.b_pair_pi.release(); }
  s.pair_cache_on = on;
  return on;
}

template<int D, int K>
static void sweep_dk(System & s, int op) {
  const bool cache = pair_cache_ready(s);
  const bool a_ok = cache && s.pair_a_epoch == s.search_epoch;   // A of this neighbour structure is stored
  switch(op) {
    case FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED:
    case 100:   // fsph_step: density / P / cs only (the signal speed rides on the acceleration sweep)
      s.pair_pi_valid = false;   // rho, cs change
      if(cache && !a_ok) {
        launch_sweep_v<D, K, OP_DENSITY, CM_FULL>(s);
        s.pair_a_epoch = s.search_epoch;
      }
      else launch_sweep_v<D, K, OP_DENSITY, CM_NONE>(s);
      if(op != 100) launch_sweep<D, K, OP_VSIG, CM_NONE>(s);
      break;
    case FSPH_SWEEP_ACCELERATION:
      if(a_ok) { launch_sweep_v<D, K, OP_ACC, CM_FULL>(s); s.pair_pi_valid = true; }
      else launch_sweep_v<D, K, OP_ACC, CM_NONE>(s);
      break;
    case 101:
      if(a_ok) { launch_sweep_v<D, K, OP_ACCV, CM_FULL>(s); s.pair_pi_valid = true; }
      else launch_sweep_v<D, K, OP_ACCV, CM_NONE>(s);
      break;
    case FSPH_SWEEP_DUDT:
      if(s.iteration < s.par.relaxation_steps)
        FSPH_LAUNCH(k_fill, (unsigned)((s.n + 255) / 256), 256, s.stream, s.col[F_DUDT].as<double>(), s.n, 0.0);
      else if(a_ok && s.pair_pi_valid) launch_sweep_v<D, K, OP_DUDT, CM_FULL>(s);
      else if(a_ok) launch_sweep_v<D, K, OP_DUDT, CM_A_ONLY>(s);
      else launch_sweep_v<D, K, OP_DUDT, CM_NONE>(s);
      break;
    case FSPH_SWEEP_DEDT:
      if(a_ok && s.pair_pi_valid) launch_sweep_v<D, K, OP_DEDT, CM_FULL>(s);
      else if(a_ok) launch_sweep_v<D, K, OP_DEDT, CM_A_ONLY>(s);
      else launch_sweep_v<D, K, OP_DEDT, CM_NONE>(s);
      break;
    case FSPH_SWEEP_COMPUTE_ALPHA:
      if(s.par.sph_viscosity != FSPH_VISC_CULLEN) throw std::runtime_error("compute_alpha needs sph_viscosity = cullen (its columns are not allocated otherwise)");
      s.pair_pi_valid = false;   // alpha changes
      launch_sweep<D, K, OP_ALPHA, CM_NONE>(s);
      break;
    case FSPH_SWEEP_SHORT_RANGE_REPULSION:
      launch_sweep<D, FSPH_KERNEL_WENDLAND_C4, OP_REPULSION, CM_NONE>(s);   // no kernel function in this sweep
      break;
    default: throw std::runtime_error("unknown sweep op");
  }
}

template<int D>
static void sweep_d(System & s, int op) {
  switch(s.par.sph_kernel) {
    case FSPH_KERNEL_CUBIC_SPLINE: sweep_dk<D, FSPH_KERNEL_CUBIC_SPLINE>(s, op); break;
    case FSPH_KERNEL_QUINTIC_SPLINE: sweep_dk<D, FSPH_KERNEL_QUINTIC_SPLINE>(s, op); break;
    case FSPH_KERNEL_WENDLAND_C2: sweep_dk<D, FSPH_KERNEL_WENDLAND_C2>(s, op); break;
    case FSPH_KERNEL_WENDLAND_C4: sweep_dk<D, FSPH_KERNEL_WENDLAND_C4>(s, op); break;
    case FSPH_KERNEL_WENDLAND_C6: sweep_dk<D, FSPH_KERNEL_WENDLAND_C6>(s, op); break;
    case FSPH_KERNEL_GAUSSIAN: sweep_dk<D, FSPH_KERNEL_GAUSSIAN>(s, op); break;
    case FSPH_KERNEL_SUPER_GAUSSIAN: sweep_dk<D, FSPH_KERNEL_SUPER_GAUSSIAN>(s, op); break;
    case FSPH_KERNEL_SINC: {
      // set_sinc_kernel_normalization, kernels.h:67-77
      const double si = s.par.sph_sinc_index, sq = sqrt(si), s2 = si * si;
      const double b0[3] = {-1.5404568e-2, 5.2245027e-2, 2.7012593e-2}, b1[3] = {3.6632876e-1, 1.3090245e-1, 2.0410827e-2};
      const double b2[3] = {-4.6519576e-4, 1.9358485e-2, 3.7451957e-3}, b3[3] = {7.3658324e-2, -6.1642906e-3, 4.7013839e-2};
      double h2[2] = {si, 0.0};
      if(D == 1) h2[1] = 2. * (b0[0] + b1[0] * sq + b2[0] * si + b3[0] / sq);
      else if(D == 2) h2[1] = 4. * (b0[1] + b1[1] * si + b2[1] / si + b3[1] / s2);
      else h2[1] = 8. * (b0[2] + b1[2] * sq + b2[2] * si + b3[2] * sq * si);
      FSPH_CUDA_OK(cudaMemcpyToSymbolAsync(c_sinc, h2, sizeof(h2), 0, cudaMemcpyHostToDevice, s.stream));
      sweep_dk<D, FSPH_KERNEL_SINC>(s, op);
    } break;
    default: throw std::runtime_error("unknown sph_kernel");
  }
}

// This file is compiled once per dimension (-DFSPH_SWEEPS_DIM=1|2|3, flecsph_b200/build.py): every (kernel, sweep, cache
// mode, viscosity) combination is a separate kernel, and the three translation units build in parallel.
#ifndef FSPH_SWEEPS_DIM
#error "compile with -DFSPH_SWEEPS_DIM=1, 2 or 3"
#endif
#if FSPH_SWEEPS_DIM == 1
void sweep_1d(System & s, int op) { sweep_d<1>(s, op); }
#elif FSPH_SWEEPS_DIM == 2
void sweep_2d(System & s, int op) { sweep_d<2>(s, op); }
#else
void sweep_1d(System & s, int op);
void sweep_2d(System & s, int op);
void sweep(System & s, int op) {
  if(!s.tree_valid) throw std::runtime_error("sweep before update_iteration");
  if(!s.search_valid) neighbor_search(s);
  if(op == FSPH_SWEEP_NEIGHBORS_ONLY) return;
  if(s.dim == 1) sweep_1d(s, op);
  else if(s.dim == 2) sweep_2d(s, op);
  else sweep_d<3>(s, op);
  FSPH_CUDA_OK(cudaGetLastError());
}
#endif

}  // namespace fsph
