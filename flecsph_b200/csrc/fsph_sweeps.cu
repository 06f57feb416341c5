// fsph_sweeps.cu -- the fp64 pair sweeps over the neighbour structure of fsph_search.cu.
//
// Replaces the functors the drivers pass to body_system::apply_in_smoothinglength
// (mpisph/bodies_system.h:304; include/physics/default_physics.h):
//   compute_density_pressure_soundspeed :309-320  (compute_density :193-228, recover_internal_energy :126-145,
//                                                  eos ideal eos/eos.h:166-183, compute_signalspeed :236-265)
//   compute_acceleration :348-408, compute_dudt :421-487, compute_dedt :502-566
// with the SPH kernels of include/physics/kernels.h (cubic :106-153, quintic :217-268, Wendland C2 :282-365,
// C4 :375-463, C6 :480-566) and the constant artificial viscosity of include/physics/viscosity.h:81-137.
//
// One CTA per work cell (thread k <-> particle k of the cell).  The cell's candidate particles are staged
// in shared memory in tiles of 256 as SoA columns (coalesced gathers: candidates are runs of consecutive
// Morton indices); each thread then walks the set bits of its own mask words and reads its neighbours'
// fields from shared memory.  FP64-pipe bound; FMA contraction is allowed here (1e-10 bar).
#include <stdexcept>
#include <cfloat>
#include "fsph_internal.h"

namespace fsph {

// candidates staged per shared-memory tile: a smaller tile lets more CTAs share an SM (the tile is most of a CTA's
// shared memory) at the price of more tile rounds; measured on B200 (64 M Sedov): density 256, the wide pair sweeps 128
#ifndef FSPH_SW_TILE_DENS
#define FSPH_SW_TILE_DENS 256
#endif
#ifndef FSPH_SW_TILE_PAIR
#define FSPH_SW_TILE_PAIR 128
#endif
// particles per thread of the pair sweeps.  2 (k_sweep2: key neighbours share a thread and read common neighbours
// once) halves the shared-memory wavefronts per pair but was 1.5x SLOWER on B200 (64 M Sedov: acceleration 103 vs 64 ms):
// half as many warps per SM leave the fp64 dependency chains exposed.  Kept for the record; the default is 1.
// 1: every thread keeps the mask words of the current tile in registers (loaded straight from global memory) instead of
// staging them in shared memory
#ifndef FSPH_SW_MASK_REGS
#define FSPH_SW_MASK_REGS 1
#endif
#ifndef FSPH_SWEEP_PPT
#define FSPH_SWEEP_PPT 1
#endif
template<int OP> struct SwTile { static constexpr int N = (OP == 0 /* OP_DENSITY */ || OP == 5 /* OP_VSIG */) ? FSPH_SW_TILE_DENS : FSPH_SW_TILE_PAIR; };
// neighbour-list entries decoded per round and thread (measured at 64 M: 16 is best for the energy sweeps, 32 for the rest)
template<int OP> struct SwLcap { static constexpr int N = (OP == 2 /* OP_DUDT */ || OP == 3 /* OP_DEDT */) ? 16 : 32; };
constexpr int SW_LCAP = 32;   // k_sweep2 (experimental)
constexpr double TINY = 1e10 * DBL_MIN;  // kernels.h:64, viscosity.h:36
constexpr double PI = 3.14159265358979323846;

enum SweepOp { OP_DENSITY = 0, OP_ACC = 1, OP_DUDT = 2, OP_DEDT = 3, OP_VSIG = 5, OP_ACCV = 6 };

// ---- SPH kernels: value W(r,h) and the scalar F with grad W = r_ab * F ------------------------------------

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
// 1/h and 1/(r + TINY) from a single fp64 reciprocal
__device__ __forceinline__ void inv_h_r(double h, double r, double & ih, double & ir) {
  const double rt = r + TINY;
  const double inv = rcp_nr(h * rt);
  ih = inv * rt;
  ir = inv * h;
}

template<int D, int K> struct SphKernel;

template<int D> struct SphKernel<D, FSPH_KERNEL_CUBIC_SPLINE> {
  static __device__ __forceinline__ double sigma() { return D == 1 ? 4. / 3. : (D == 2 ? 40. / (7. * PI) : 8. / PI); }
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = 2. * r * ih;
    double res = 0.;
    if(q < 2.0) {
      res = q < 1.0 ? 1.0 - 1.5 * q * q + .75 * q * q * q : 0.25 * (2 - q) * (2 - q) * (2 - q);
      res *= sigma() * ipowD<D>(ih);
    }
    return res;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    const double q = 2. * r * ih;
    double res = 0.;
    if(q < 2.0) {
      const double sg = 2. * sigma() * ipowD<D>(ih) * ih;
      const double dW = q < 1.0 ? -3.0 * q + 9. / 4. * q * q : -.75 * (2 - q) * (2 - q);
      res = sg * dW * ir;
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
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
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
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
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
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double res = D == 1 ? sigma() * ih * o2 * o * (3 * q + 1) : sigma() * ipowD<D>(ih) * o2 * o2 * (4 * q + 1);
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double sg = D == 1 ? sigma() * ih * ih : 2. * sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -12. * q * o2 : -10. * q * o2 * o;
    return q < 1.0 ? sg * dW * ir : 0.0;
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
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = r * ih, o = 1 - q, o2 = o * o;
    double res;
    if(D == 1) res = sigma() * ih * (o2 * o) * o2 * (1 + q * (5 + q * 8));
    else res = sigma() * ipowD<D>(ih) * (o2 * o2 * o2) * (1 + q * (6 + q * (35. / 3.)));
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    const double q = r * ih, o = 1 - q, o2 = o * o;
    const double sg = 14. * sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -q * o2 * o2 * (1 + 4. * q) : -4. / 3. * q * (o2 * o) * o2 * (1 + 5. * q);
    return q < 1.0 ? sg * dW * ir : 0.0;
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
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = r * ih, o = 1 - q, o2 = o * o, o4 = o2 * o2;
    double res;
    if(D == 1) res = sigma() * ih * (o2 * o) * o4 * (1 + q * (7 + q * (19 + q * 21)));
    else res = sigma() * ipowD<D>(ih) * o4 * o4 * (1 + q * (8 + q * (25 + q * 32)));
    return q < 1.0 ? res : 0.0;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    const double q = r * ih, o = 1 - q, o2 = o * o, o3 = o2 * o, o4 = o2 * o2;
    const double sg = sigma() * ipowD<D>(ih) * ih;
    const double dW = D == 1 ? -6. * q * o3 * o3 * (3 + q * (18 + q * 35)) : -22. * q * o4 * o3 * (1 + q * (7 + q * 16));
    return q < 1.0 ? sg * dW * ir : 0.0;
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
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = 3. * r * ih;
    return q <= 3. ? sigma() * ipowD<D>(ih) * exp(-q * q) : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 3. * r * ih;
    const double sg = 3. * sigma() * ipowD<D>(ih) * ih;
    return q < 3. ? sg * (-2. * q * exp(-q * q)) * ir : 0.0;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    return Fi(r, ih, ir);
  }
};
template<int D> struct SphKernel<D, FSPH_KERNEL_SUPER_GAUSSIAN> {
  static __device__ __forceinline__ double sigma() {
    return D == 1 ? 1.69225265632490690298438831093977954 : (D == 2 ? 2.86196342089351035329931251213987773 : 4.83280745997963041039733264826054666);
  }
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
    const double q = 3. * r * ih, q2 = q * q;
    return q < 3. ? sigma() * ipowD<D>(ih) * exp(-q2) * (D / 2.0 + 1 - q2) : 0.0;
  }
  static __device__ __forceinline__ double Fi(double r, double ih, double ir) {
    const double q = 3. * r * ih;
    const double sg = 3. * sigma() * ipowD<D>(ih) * ih;
    const double dW = exp(-q * q) * (2. * q * q * q - (D + 4.) * q);
    return q < 3. ? sg * dW * ir : 0.0;
  }
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    return Fi(r, ih, ir);
  }
};
// sinc kernel (kernels.h:626-699): index n and its normalisation (set_sinc_kernel_normalization, :67-77) are run-time
// parameters; they live in constant memory, refreshed by sweep() before the launches
__constant__ double c_sinc[2];  // {n, sigma(n, D)}
template<int D> struct SphKernel<D, FSPH_KERNEL_SINC> {
  static __device__ __forceinline__ double W(double r, double h) {
    const double ih = rcp_nr(h);
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
  static __device__ __forceinline__ double F(double r, double h) {
    double ih, ir;
    inv_h_r(h, r, ih, ir);
    return Fi(r, ih, ir);
  }
};

// ---- candidate tile columns per sweep ---------------------------------------------------------------------
template<int D, int OP> struct TileCols {
  // ACCV = acceleration + signal speed: ACC's columns + velocity + previous soundspeed
  static constexpr int NF = OP == OP_DENSITY ? D + 2 : (OP == OP_VSIG ? 2 * D + 2 : (OP == OP_ACC ? 2 * D + 5 : (OP == OP_ACCV ? 3 * D + 6 : 3 * D + 5)));
  static __device__ __forceinline__ int col(int j) {
    if(j < D) return F_POS + j;
    if(OP == OP_DENSITY) return j == D ? F_H : F_MASS;
    if(OP == OP_VSIG) return j < 2 * D ? F_VEL + (j - D) : (j == 2 * D ? F_CS : F_CSOLD);
    const int r = j - D;
    if(r == 0) return F_H;
    if(r == 1) return F_MASS;
    if(r == 2) return F_RHO;
    if(r == 3) return F_PRHO2;   // P / rho^2, refreshed by k_prho2 right before the sweep
    if(r == 4) return F_CS;
    if(r < 5 + D) return F_VELH + (r - 5);
    if(r < 5 + 2 * D) return F_VEL + (r - 5 - D);
    return F_CSOLD;
  }
};

template<int D>
__device__ __forceinline__ double norm_vec(const double * v) {
  if(D == 1) return fabs(v[0]);
  double s = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) s += v[d] * v[d];
  return sqrt(s);
}

// state of the particle a thread owns + its accumulators
template<int D> struct Own {
  double xa[D], va[D], vha[D];
  double ha, rho_a, cs_a, Prho2_a;
  double acc0, acc1, accv[D], vsig;
  int ia;
};

// one neighbour pair (a, candidate ci of the current tile); `valid` = false makes it a no-op (padding slot of
// the two-pairs-per-iteration loop)
template<int D, int K, int OP>
__device__ __forceinline__ void pair_body(Own<D> & o, const double (*tile)[SwTile<OP>::N], const int32_t * tidx, const PhysParams & pp,
  const int ci, const bool valid) {
  double rab[D];
#pragma unroll
  for(int d = 0; d < D; ++d) rab[d] = o.xa[d] - tile[d][ci];
  if(OP == OP_DENSITY) {
    double d2 = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) d2 += rab[d] * rab[d];
    const double r = D == 1 ? fabs(rab[0]) : d2 * rsqrt_nr(fmax(d2, 1e-280));
    const double hab = .5 * (o.ha + tile[D][ci]);
    const double t = tile[D + 1][ci] * SphKernel<D, K>::W(r, hab);
    o.acc0 += valid ? t : 0.0;
  }
  else if(OP == OP_VSIG) {
    // compute_signalspeed: b contributes its new soundspeed iff it was processed before a, i.e. iff its
    // (descending-key processing order) index is larger
    const double csb = tidx[ci] > o.ia ? tile[2 * D][ci] : tile[2 * D + 1][ci];
    const double cab = fmax(o.cs_a, csb);
    const double dist = norm_vec<D>(rab);
    double vr = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) vr += (o.va[d] - tile[D + d][ci]) * rab[d];
    vr = vr / dist;                            // 0/0 = NaN for the particle itself
    const double mn = (0.0 < vr) ? 0.0 : vr;  // std::min(vr, 0.0), NaN-preserving
    const double val = cab - mn;
    if(valid && o.vsig < val) o.vsig = val;   // std::max(vsig, val): NaN (self) dropped
  }
  else {
    const double hab = .5 * (o.ha + tile[D][ci]);
    double dotvr = 0.0, dist2 = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      dotvr += (o.vha[d] - tile[D + 5 + d][ci]) * rab[d];
      dist2 += rab[d] * rab[d];
    }
    const double mb = (dist2 == 0.0 || !valid) ? 0.0 : tile[D + 1][ci];   // m_b * (pos_b != pos_a)
    const double rho_ab = .5 * (o.rho_a + tile[D + 2][ci]);
    double Pi = 0.0;
    if(dotvr < 0.0) {
      // mu = h dot / (r^2 + eps h^2 + TINY);  Pi = (-alpha c + beta mu) mu / rho  -- one division for both
      const double den = dist2 + pp.epsilon * hab * hab + TINY;
      const double inv = rcp_nr(den * rho_ab);
      const double mu = hab * dotvr * (inv * rho_ab);
      const double c_ab = .5 * (o.cs_a + tile[D + 4][ci]);
      Pi = (-pp.alpha * c_ab + pp.beta * mu) * mu * (inv * den);
    }
    // r, 1/r and 1/h_ab: one reciprocal square root + one reciprocal (dist2 clamped away from 0: the particle
    // itself / coincident particles carry m_b = 0 and the kernel gradient is 0 at r = 0 anyway)
    const double ir = rsqrt_nr(fmax(dist2, 1e-280));
    const double r = dist2 * ir;
    const double ih = rcp_nr(hab);
    const double Fk = SphKernel<D, K>::Fi(r, ih, ir);
    const double Prho2_b = tile[D + 3][ci];
    if(OP == OP_ACC || OP == OP_ACCV) {
      const double coef = -mb * (o.Prho2_a + Prho2_b + Pi) * Fk;
#pragma unroll
      for(int d = 0; d < D; ++d) o.accv[d] += coef * rab[d];
      if(OP == OP_ACCV) {
        // compute_signalspeed fused into the acceleration sweep (fsph_step only): all new soundspeeds exist by now
        const double csb = tidx[ci] > o.ia ? tile[D + 4][ci] : tile[3 * D + 5][ci];
        const double cab = fmax(o.cs_a, csb);
        double vr = 0.0;
#pragma unroll
        for(int d = 0; d < D; ++d) vr += (o.va[d] - tile[2 * D + 5 + d][ci]) * rab[d];
        vr = vr * ir;                              // dist2 == 0 (self) is excluded below instead of via NaN
        const double mn = (0.0 < vr) ? 0.0 : vr;
        const double val = cab - mn;
        if(valid && dist2 != 0.0 && o.vsig < val) o.vsig = val;
      }
    }
    else if(OP == OP_DUDT) {
      double vdw = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) vdw += (o.va[d] - tile[2 * D + 5 + d][ci]) * rab[d];
      vdw *= Fk;
      o.acc0 += mb * vdw;
      o.acc1 += mb * vdw * Pi;
    }
    else {  // OP_DEDT
      double vadw = 0.0, vbdw = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        vadw += o.va[d] * rab[d];
        vbdw += tile[2 * D + 5 + d][ci] * rab[d];
      }
      vadw *= Fk;
      vbdw *= Fk;
      o.acc0 -= mb * (o.Prho2_a * vbdw + vadw * Prho2_b + .5 * Pi * (vbdw + vadw));
    }
  }
}

// state of the particle `ia` of a sweep (the columns op OP needs), accumulators cleared
template<int D, int OP>
__device__ __forceinline__ void load_own(Own<D> & o, const Cols & c, const int ia, const bool active) {
  o.ia = ia;
  o.ha = o.rho_a = o.cs_a = o.Prho2_a = 0.0;
  o.acc0 = o.acc1 = o.vsig = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) { o.xa[d] = 0; o.va[d] = 0; o.vha[d] = 0; o.accv[d] = 0; }
  if(active) {
#pragma unroll
    for(int d = 0; d < D; ++d) o.xa[d] = c.f[F_POS + d][ia];
    o.ha = c.f[F_H][ia];
    if(OP == OP_VSIG || OP == OP_DUDT || OP == OP_DEDT || OP == OP_ACCV) {
#pragma unroll
      for(int d = 0; d < D; ++d) o.va[d] = c.f[F_VEL + d][ia];
    }
    if(OP == OP_VSIG) o.cs_a = c.f[F_CS][ia];
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
template<int D, int OP>
__device__ __forceinline__ void store_own(const Own<D> & o, const Cols & c, const PhysParams & pp, uint32_t * __restrict__ status) {
  const int ia = o.ia;
  if(OP == OP_DENSITY) {
    const double rho = o.acc0;
    if(!(rho > 0)) atomicOr(status, (uint32_t)ERR_RHO_NONPOS);
    c.f[F_RHO][ia] = rho;
    double u;
    if(pp.evolve_u && pp.thermokinetic) {
      // recover_internal_energy (no external potential)
      double v2 = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) { const double v = c.f[F_VEL + d][ia]; v2 += v * v; }
      u = c.f[F_E][ia] - .5 * v2 - 0.0;
      if(!(u > 0)) atomicOr(status, (uint32_t)ERR_U_NONPOS);
      c.f[F_U][ia] = u;
    }
    else u = c.f[F_U][ia];
    double Pn, csn;
    eos_p_cs(pp, rho, u, pp.eos == FSPH_EOS_POLYTROPIC ? c.f[F_ADIA][ia] : 0.0, Pn, csn);
    c.f[F_P][ia] = Pn;
    c.f[F_CSOLD][ia] = c.f[F_CS][ia];
    c.f[F_CS][ia] = csn;
  }
  else if(OP == OP_VSIG) c.f[F_VSIG][ia] = o.vsig;
  else if(OP == OP_ACC || OP == OP_ACCV) {
#pragma unroll
    for(int d = 0; d < D; ++d) {
      c.f[F_ACC + d][ia] = o.accv[d];  // + external_force::acceleration == 0 (external_force_type "none")
      c.f[F_GACC + d][ia] = 0.0;
    }
    c.f[F_GPOT][ia] = 0.0;
    if(OP == OP_ACCV) c.f[F_VSIG][ia] = o.vsig;
  }
  else if(OP == OP_DUDT) c.f[F_DUDT][ia] = o.Prho2_a * o.acc0 + .5 * o.acc1;
  else {
    double gv = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) gv += c.f[F_GACC + d][ia] * o.va[d];
    c.f[F_DEDT][ia] = o.acc0 + gv;
  }
}

template<int D, int K, int OP, int NW>
__global__ void __launch_bounds__(32 * NW) k_sweep(Tree t, Cols c, Search sr, PhysParams pp, int g_begin, uint32_t * __restrict__ status) {
  using TC = TileCols<D, OP>;
  constexpr int NF = TC::NF;
  constexpr int NT = 32 * NW;
  constexpr int TILE = SwTile<OP>::N;
  constexpr int LCAP = SwLcap<OP>::N;
  constexpr bool NEED_IDX = OP == OP_VSIG || OP == OP_ACCV;
  __shared__ double tile[NF][TILE];
  __shared__ int32_t tidx[NEED_IDX ? TILE : 1];
#if !FSPH_SW_MASK_REGS
  __shared__ uint32_t mtile[TILE / 32][NT];   // this tile's mask words, [word][particle]
#endif
  __shared__ uint8_t nlist[LCAP][NT];         // decoded neighbour lists of one round, [entry][particle]

  const int g = g_begin + blockIdx.x, tid = threadIdx.x;
  int start, P;
  {
    int v = t.cell[sr.wc_cell[g]];
    if(v >= 0) { start = v; P = 1; }
    else { start = t.node_start[~v]; P = t.sub[~v]; }
  }
  const int Ppad = (P + 31) & ~31;
  const uint64_t co = sr.cand_off[g], mo = sr.mask_off[g];
  const int C = sr.ncand[g];
  const bool active = tid < P;

  Own<D> o;
  load_own<D, OP>(o, c, start + tid, active);

  for(int tb = 0; tb < C; tb += TILE) {
    const int tc = min(TILE, C - tb);
    __syncthreads();
    {
      // all candidate indices of this thread first, then the column gathers: (TILE / NT) x NF loads in flight
      constexpr int LPT = (TILE + NT - 1) / NT;
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
          if(NEED_IDX) tidx[i] = pidx[u];
#pragma unroll
          for(int j = 0; j < NF; ++j) tile[j][i] = c.f[TC::col(j)][pidx[u]];
        }
      }
    }
#if FSPH_SW_MASK_REGS
    // this thread's mask words of the tile straight into registers (coalesced: the mask is laid out [word][particle]);
    // the loads are in flight together with the tile gathers above
    uint32_t mw[TILE / 32];
    {
      const int nwt = (tc + 31) >> 5;
      const uint32_t * mrow = sr.mask + mo + (uint64_t)(tb >> 5) * Ppad + tid;
#pragma unroll
      for(int w = 0; w < TILE / 32; ++w) mw[w] = (active && w < nwt) ? mrow[(size_t)w * Ppad] : 0u;
    }
    __syncthreads();
    for(;;) {
      int cnt = 0;
#pragma unroll
      for(int w = 0; w < TILE / 32; ++w) {
        while(mw[w] != 0u && cnt < LCAP) {
          nlist[cnt++][tid] = (uint8_t)(w * 32 + __ffs(mw[w]) - 1);
          mw[w] &= mw[w] - 1;
        }
      }
      const int jmax = __reduce_max_sync(0xffffffffu, cnt);
      if(jmax == 0) break;
      for(int j = 0; j < jmax; j += 2) {
        const bool v0 = j < cnt, v1 = j + 1 < cnt;
        const int ci0 = v0 ? nlist[j][tid] : 0;
        const int ci1 = v1 ? nlist[j + 1][tid] : ci0;
        pair_body<D, K, OP>(o, tile, tidx, pp, ci0, v0);
        pair_body<D, K, OP>(o, tile, tidx, pp, ci1, v1);
      }
      if(jmax < LCAP) break;
    }
#else
    {
      const int nwt = (tc + 31) >> 5;
      const uint32_t * mrow = sr.mask + mo + (uint64_t)(tb >> 5) * Ppad;
      for(int i = tid; i < nwt * Ppad; i += NT) mtile[i / Ppad][i % Ppad] = mrow[i];
    }
    __syncthreads();
    // Decode this thread's mask bits into a short list of tile-local candidate indices (<= LCAP per round),
    // then run a uniform, branch-free pair loop over the lists: two pairs per iteration (independent fp64
    // dependency chains), trip count = the warp's longest list.
    const int nw = active ? (tc + 31) >> 5 : 0;
    int w = 0;
    uint32_t m = nw > 0 ? mtile[0][tid] : 0u;
    for(;;) {
      int cnt = 0;
      while(cnt < LCAP) {
        while(m == 0 && w + 1 < nw) m = mtile[++w][tid];
        if(m == 0) break;
        nlist[cnt++][tid] = (uint8_t)(w * 32 + __ffs(m) - 1);
        m &= m - 1;
      }
      const int jmax = __reduce_max_sync(0xffffffffu, cnt);
      if(jmax == 0) break;
      for(int j = 0; j < jmax; j += 2) {
        const bool v0 = j < cnt, v1 = j + 1 < cnt;
        const int ci0 = v0 ? nlist[j][tid] : 0;
        const int ci1 = v1 ? nlist[j + 1][tid] : ci0;
        pair_body<D, K, OP>(o, tile, tidx, pp, ci0, v0);
        pair_body<D, K, OP>(o, tile, tidx, pp, ci1, v1);
      }
      if(jmax < LCAP) break;
    }
#endif
  }
  if(active) store_own<D, OP>(o, c, pp, status);
}

// Two particles per thread.  The pair loop is bound by shared-memory wavefronts (ncu: l1tex data pipe 77 % of peak,
// fp64 pipe 55 %): every pair gathers its neighbour's 5-15 doubles from the tile at a lane-private address.  Key
// neighbours (2t, 2t+1) share ~80 % of their neighbours, so a thread that owns both reads each common neighbour
// once and evaluates it for both (a neighbour of only one of them costs a predicated-off evaluation).
template<int D, int K, int OP, int NW>
__global__ void __launch_bounds__(32 * NW) k_sweep2(Tree t, Cols c, Search sr, PhysParams pp, int g_begin, uint32_t * __restrict__ status) {
  using TC = TileCols<D, OP>;
  constexpr int NF = TC::NF;
  constexpr int NT = 32 * NW;
  constexpr int TILE = SwTile<OP>::N;
  constexpr bool NEED_IDX = OP == OP_VSIG || OP == OP_ACCV;
  __shared__ double tile[NF][TILE];
  __shared__ int32_t tidx[NEED_IDX ? TILE : 1];
  __shared__ uint32_t mtile[TILE / 32][2 * NT];  // this tile's mask words, [word][particle]
  __shared__ uint16_t nlist[SW_LCAP][NT];        // decoded union lists of one round: candidate | owner bits << 8

  const int g = g_begin + blockIdx.x, tid = threadIdx.x;
  int start, P;
  {
    int v = t.cell[sr.wc_cell[g]];
    if(v >= 0) { start = v; P = 1; }
    else { start = t.node_start[~v]; P = t.sub[~v]; }
  }
  const int Ppad = (P + 31) & ~31;
  const uint64_t co = sr.cand_off[g], mo = sr.mask_off[g];
  const int C = sr.ncand[g];
  const bool act0 = 2 * tid < P, act1 = 2 * tid + 1 < P;

  Own<D> o0, o1;
  load_own<D, OP>(o0, c, start + 2 * tid, act0);
  load_own<D, OP>(o1, c, start + 2 * tid + 1, act1);

  for(int tb = 0; tb < C; tb += TILE) {
    const int tc = min(TILE, C - tb);
    __syncthreads();
    {
      constexpr int LPT = (TILE + NT - 1) / NT;
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
          if(NEED_IDX) tidx[i] = pidx[u];
#pragma unroll
          for(int j = 0; j < NF; ++j) tile[j][i] = c.f[TC::col(j)][pidx[u]];
        }
      }
    }
    {
      const int nwt = (tc + 31) >> 5;
      const uint32_t * mrow = sr.mask + mo + (uint64_t)(tb >> 5) * Ppad;
      for(int i = tid; i < nwt * Ppad; i += NT) mtile[i / Ppad][i % Ppad] = mrow[i];
    }
    __syncthreads();
    const int nw = act0 ? (tc + 31) >> 5 : 0;
    int w = 0;
    uint32_t m0 = nw > 0 ? mtile[0][2 * tid] : 0u;
    uint32_t m1 = (nw > 0 && act1) ? mtile[0][2 * tid + 1] : 0u;
    for(;;) {
      int cnt = 0;
      while(cnt < SW_LCAP) {
        while((m0 | m1) == 0 && w + 1 < nw) {
          ++w;
          m0 = mtile[w][2 * tid];
          m1 = act1 ? mtile[w][2 * tid + 1] : 0u;
        }
        const uint32_t mu = m0 | m1;
        if(mu == 0) break;
        const int b = __ffs(mu) - 1;
        nlist[cnt++][tid] = (uint16_t)((w * 32 + b) | (((m0 >> b) & 1u) << 8) | (((m1 >> b) & 1u) << 9));
        const uint32_t clr = ~(1u << b);
        m0 &= clr;
        m1 &= clr;
      }
      const int jmax = __reduce_max_sync(0xffffffffu, cnt);
      if(jmax == 0) break;
      for(int j = 0; j < jmax; ++j) {
        const int e = j < cnt ? nlist[j][tid] : 0;
        const int ci = e & 0xff;
        pair_body<D, K, OP>(o0, tile, tidx, pp, ci, (e >> 8) & 1);
        pair_body<D, K, OP>(o1, tile, tidx, pp, ci, (e >> 9) & 1);
      }
      if(jmax < SW_LCAP) break;
    }
  }
  if(act0) store_own<D, OP>(o0, c, pp, status);
  if(act1) store_own<D, OP>(o1, c, pp, status);
}

__global__ void __launch_bounds__(256) k_prho2(Cols c, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) {
    const double rho = c.f[F_RHO][i];
    c.f[F_PRHO2][i] = c.f[F_P][i] / (rho * rho);
  }
}

__global__ void __launch_bounds__(256) k_fill(double * p, int64_t n, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) p[i] = v;
}

template<int D, int K, int OP>
static void launch_sweep(System & s) {
  const Search & sr = s.search;
  if(OP != OP_DENSITY && OP != OP_VSIG) FSPH_LAUNCH(k_prho2, (unsigned)((s.n + 255) / 256), 256, s.stream, s.cols(), s.n);
  uint32_t * st = s.b_status.as<uint32_t>();
  for(int cls = 0; cls < 4; ++cls) {
    const int b = sr.class_off[cls], n = sr.class_off[cls + 1] - b;
    if(n <= 0) continue;
#if FSPH_SWEEP_PPT == 2
    // two particles per thread: classes of <= 32 / 64 particles fit one warp, <= 96 / 128 two warps
    switch(cls) {
      case 0: FSPH_LAUNCH((k_sweep<D, K, OP, 1>), n, 32, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
      case 1: FSPH_LAUNCH((k_sweep2<D, K, OP, 1>), n, 32, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
      default: FSPH_LAUNCH((k_sweep2<D, K, OP, 2>), n, 64, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
    }
#else
    switch(cls) {
      case 0: FSPH_LAUNCH((k_sweep<D, K, OP, 1>), n, 32, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
      case 1: FSPH_LAUNCH((k_sweep<D, K, OP, 2>), n, 64, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
      case 2: FSPH_LAUNCH((k_sweep<D, K, OP, 3>), n, 96, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
      default: FSPH_LAUNCH((k_sweep<D, K, OP, 4>), n, 128, s.stream, s.tree, s.cols(), sr, s.phys(), b, st); break;
    }
#endif
  }
}

template<int D, int K>
static void sweep_dk(System & s, int op) {
  switch(op) {
    case FSPH_SWEEP_DENSITY_PRESSURE_SOUNDSPEED:
      launch_sweep<D, K, OP_DENSITY>(s);
      launch_sweep<D, K, OP_VSIG>(s);
      break;
    case FSPH_SWEEP_ACCELERATION: launch_sweep<D, K, OP_ACC>(s); break;
    case 100: launch_sweep<D, K, OP_DENSITY>(s); break;   // fsph_step: density / P / cs only ...
    case 101: launch_sweep<D, K, OP_ACCV>(s); break;      // ... and the signal speed rides on the acceleration sweep
    case FSPH_SWEEP_DUDT:
      if(s.iteration < s.par.relaxation_steps)
        FSPH_LAUNCH(k_fill, (unsigned)((s.n + 255) / 256), 256, s.stream, s.col[F_DUDT].as<double>(), s.n, 0.0);
      else launch_sweep<D, K, OP_DUDT>(s);
      break;
    case FSPH_SWEEP_DEDT: launch_sweep<D, K, OP_DEDT>(s); break;
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

void sweep(System & s, int op) {
  if(!s.tree_valid) throw std::runtime_error("sweep before update_iteration");
  if(!s.search_valid) neighbor_search(s);
  if(op == FSPH_SWEEP_NEIGHBORS_ONLY) return;
  if(s.dim == 1) sweep_d<1>(s, op);
  else if(s.dim == 2) sweep_d<2>(s, op);
  else sweep_d<3>(s, op);
  FSPH_CUDA_OK(cudaGetLastError());
}

}  // namespace fsph
