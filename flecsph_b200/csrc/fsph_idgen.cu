// fsph_idgen.cu -- initial data on the device (compiled with -fmad=false: positions equal the host generators' bit for bit).
//
// Replaces, for the lattice-in-a-box / lattice-in-a-sphere initial data of the benchmark configurations, the host loops
// of the reference's single-rank generators:
//   particle_lattice::generator_lattice_3d   include/lattice.h:224-293 (rectangular, HCP, FCC)
//   particle_lattice::in_domain_3d           include/lattice.h:107-133
//   the per-particle assignments of app/id_generators/sedov/main.cc:262-304, noh/main.cc:236-292,
//   collapse/main.cc (uniform density / pressure, blast energy inside r < r_blast, radial inflow).
// The loop variables of the reference are accumulated sequentially (`p += d`), so the coordinates each loop visits
// are tabulated on the host with the same additions (one z table, one y table per stacking state, one x table per
// (stacking state, row parity): a few thousand doubles); the device forms their product, applies in_domain_3d and
// compacts the accepted sites in generation order (flag, scan, emit).  At 64 M bodies this replaces 40-100 s of host
// work and an 8 GB upload by ~10 ms on the device.
#include <cmath>
#include <stdexcept>
#include <vector>
#include "fsph_internal.h"

namespace fsph {

struct LatticeTabs {
  const double * zs;      // [nz]
  const int32_t * zl_of;  // [nz] stacking state of each layer
  const double * ys;      // [nstate][LY]
  const int32_t * ny;     // [nstate]
  const double * xs;      // [nstate][2][LX]
  const int32_t * nx;     // [nstate][2]
  int nz, LY, LX, iz0, iz1;
  double bmin[3], bmax[3];
  int domain_type;
};

__device__ __forceinline__ bool site_in_domain(const LatticeTabs & T, double x, double y, double z) {
  // in_domain_3d, lattice.h:107-133 (b_tol = 1e-6)
  const double dx = 1e-6 * (T.bmax[0] - T.bmin[0]);
  const double xmin = T.bmin[0] - dx, xmax = T.bmax[0] - dx, ymin = T.bmin[1] - dx, ymax = T.bmax[1] - dx, zmin = T.bmin[2] - dx,
               zmax = T.bmax[2] - dx;
  bool in = x > xmin && x < xmax && y > ymin && y < ymax && z > zmin && z < zmax;
  if(T.domain_type == 1) {
    const double r = 0.5 * fmin(fmin(xmax - xmin, ymax - ymin), zmax - zmin);
    const double x0 = 0.5 * (xmin + xmax), y0 = 0.5 * (ymin + ymax), z0 = 0.5 * (zmin + zmax);
    in = in && (((x - x0) * (x - x0) + (y - y0) * (y - y0)) + (z - z0) * (z - z0) < r * r);
  }
  return in;
}

// site s <-> (layer, row, column) of the padded product; MODE 0: flag, MODE 1: emit at the scanned offset
template<int MODE>
__global__ void __launch_bounds__(256) k_lattice(LatticeTabs T, int64_t nsite, uint32_t * __restrict__ flag, const uint32_t * __restrict__ off,
  Cols c, int64_t id_offset) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(s >= nsite) return;
  const int ix = (int)(s % T.LX);
  const int64_t r = s / T.LX;
  const int iy = (int)(r % T.LY), iz = T.iz0 + (int)(r / T.LY);
  const int zl = T.zl_of[iz];
  bool in = iy < T.ny[zl] && ix < T.nx[zl * 2 + (iy & 1)];
  double x = 0, y = 0, z = 0;
  if(in) {
    x = T.xs[(size_t)(zl * 2 + (iy & 1)) * T.LX + ix];
    y = T.ys[(size_t)zl * T.LY + iy];
    z = T.zs[iz];
    in = site_in_domain(T, x, y, z);
  }
  if(MODE == 0) flag[s] = in ? 1u : 0u;
  else if(in) {
    const uint32_t o = off[s];
    c.f[F_POS][o] = x;
    c.f[F_POS + 1][o] = y;
    c.f[F_POS + 2][o] = z;
    c.id[o] = id_offset + o;
  }
}

__global__ void __launch_bounds__(256) k_count_within(Cols c, int64_t n, double cx, double cy, double cz, double radius,
  unsigned long long * __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  if(i < n) {
    const double x = c.f[F_POS][i] - cx, y = c.f[F_POS + 1][i] - cy, z = c.f[F_POS + 2][i] - cz;
    in = sqrt((x * x + y * y) + z * z) < radius;
  }
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}

struct IdFill {
  double mass, h, rho, u, u_blast_add, blast_radius, gamma, dt, radial_velocity, rho_eos, pressure;
};

__global__ void __launch_bounds__(256) k_id_fill(Cols c, int64_t n, IdFill f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const double x = c.f[F_POS][i], y = c.f[F_POS + 1][i], z = c.f[F_POS + 2][i];
  const double r = sqrt((x * x + y * y) + z * z);
  c.f[F_MASS][i] = f.mass;
  c.f[F_H][i] = f.h;
  c.f[F_RHO][i] = f.rho;
  // sedov/main.cc:297-304: the blast energy is shared by the bodies inside the blast radius
  const double u = f.u + (r < f.blast_radius ? f.u_blast_add : 0.0);
  c.f[F_U][i] = u;
  c.f[F_P][i] = isnan(f.pressure) ? f.rho_eos * u * (f.gamma - 1) : f.pressure;
  c.f[F_DT][i] = f.dt;
  // noh/main.cc:272-276: v = -v0 r / |r|
  const double rr = r > 0 ? r : 1.0, on = r > 0 ? 1.0 : 0.0;
  const bool still = f.radial_velocity == 0.0;
  c.f[F_VEL][i] = still ? 0.0 : f.radial_velocity * x / rr * on;
  c.f[F_VEL + 1][i] = still ? 0.0 : f.radial_velocity * y / rr * on;
  c.f[F_VEL + 2][i] = still ? 0.0 : f.radial_velocity * z / rr * on;
}

// values of `for(p = start; p < hi; p += step)` (the loop's own rounding)
static std::vector<double> axis(double start, double hi, double step) {
  std::vector<double> v;
  for(double p = start; p < hi; p += step) v.push_back(p);
  return v;
}

}  // namespace fsph

using namespace fsph;
struct fsph_system : public System {};
extern "C" int32_t fsph_resize(fsph_system * s, int64_t n);

extern "C" {

// generator_lattice_3d (lattice.h:224-293) into a 3-D system: positions in generation order, ids = id_offset + index.
// slab_k of slab_K: only that contiguous group of z layers (one group per rank of a multi-GPU run).  *n_out = bodies made.
int32_t fsph_generate_lattice(fsph_system * s, int32_t lattice_type, int32_t domain_type, const double * bmin, const double * bmax,
  double sep, int32_t slab_k, int32_t slab_K, int64_t id_offset, int64_t * n_out) {
  if(!s) return 1;
  try {
    if(s->dim != 3) throw std::runtime_error("fsph_generate_lattice: 3-D systems only");
    if(lattice_type < 0 || lattice_type > 2 || slab_K < 1 || slab_k < 0 || slab_k >= slab_K) throw std::runtime_error("bad lattice arguments");
    FSPH_CUDA_OK(cudaSetDevice(s->device));
    const double dx = sep, dy = sep * sqrt(3.) / 2., dz = sep * sqrt(2. / 3.);
    const int nstate = lattice_type == 0 ? 1 : (lattice_type == 1 ? 2 : 4);
    std::vector<double> zs = axis(bmin[2], bmax[2], lattice_type == 0 ? dx : dz);
    std::vector<int32_t> zl_of(zs.size());
    {
      double zl = 0;
      for(size_t iz = 0; iz < zs.size(); ++iz) {
        zl_of[iz] = (int32_t)zl;
        if(lattice_type == 1) zl = 1 - zl;
        else if(lattice_type == 2) zl = (zl + 1) * (zl < 3);
      }
    }
    std::vector<std::vector<double>> ys(nstate), xs(nstate * 2);
    size_t LY = 1, LX = 1;
    for(int z = 0; z < nstate; ++z) {
      ys[z] = lattice_type == 0 ? axis(bmin[1], bmax[1], dx) : axis(bmin[1] - z * dy / 3, bmax[1], dy);
      LY = std::max(LY, ys[z].size());
      for(int yo = 0; yo < 2; ++yo) {
        xs[z * 2 + yo] = lattice_type == 0 ? axis(bmin[0], bmax[0], dx) : axis(bmin[0] + (yo - z) * dx / 2., bmax[0], dx);
        LX = std::max(LX, xs[z * 2 + yo].size());
      }
    }
    const int nz = (int)zs.size();
    const int iz0 = (int)(((int64_t)slab_k * nz) / slab_K), iz1 = (int)(((int64_t)(slab_k + 1) * nz) / slab_K);
    // pack the tables: zs | ys | xs (doubles), then zl_of | ny | nx (ints)
    std::vector<double> hd(zs);
    const size_t o_ys = hd.size();
    hd.resize(o_ys + nstate * LY, 0.0);
    const size_t o_xs = hd.size();
    hd.resize(o_xs + (size_t)nstate * 2 * LX, 0.0);
    std::vector<int32_t> hi(zl_of);
    const size_t o_ny = hi.size();
    hi.resize(o_ny + nstate);
    const size_t o_nx = hi.size();
    hi.resize(o_nx + nstate * 2);
    for(int z = 0; z < nstate; ++z) {
      std::copy(ys[z].begin(), ys[z].end(), hd.begin() + o_ys + z * LY);
      hi[o_ny + z] = (int32_t)ys[z].size();
      for(int yo = 0; yo < 2; ++yo) {
        std::copy(xs[z * 2 + yo].begin(), xs[z * 2 + yo].end(), hd.begin() + o_xs + (size_t)(z * 2 + yo) * LX);
        hi[o_nx + z * 2 + yo] = (int32_t)xs[z * 2 + yo].size();
      }
    }
    DevBuf dtab, itab, flags;
    dtab.reserve(hd.size() * 8);
    itab.reserve(hi.size() * 4);
    FSPH_CUDA_OK(cudaMemcpyAsync(dtab.p, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice, s->stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(itab.p, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, s->stream));
    LatticeTabs T;
    T.zs = dtab.as<double>(); T.ys = T.zs + o_ys; T.xs = T.zs + o_xs;
    T.zl_of = itab.as<int32_t>(); T.ny = T.zl_of + o_ny; T.nx = T.zl_of + o_nx;
    T.nz = nz; T.LY = (int)LY; T.LX = (int)LX; T.iz0 = iz0; T.iz1 = iz1;
    for(int d = 0; d < 3; ++d) { T.bmin[d] = bmin[d]; T.bmax[d] = bmax[d]; }
    T.domain_type = domain_type;
    const int64_t nsite = (int64_t)(iz1 - iz0) * LY * LX;
    if(nsite >= (int64_t(1) << 32)) throw std::runtime_error("lattice slab has more than 2^32 sites: use more slabs");
    flags.reserve((size_t)(nsite + 1) * 4);
    uint32_t * fl = flags.as<uint32_t>();
    const unsigned nb = (unsigned)((nsite + 255) / 256);
    int64_t n = 0;
    if(nsite > 0) {
      FSPH_LAUNCH((k_lattice<0>), nb, 256, s->stream, T, nsite, fl, nullptr, s->cols(), id_offset);
      exclusive_scan_u32(*s, fl, fl, nsite, nullptr);
      uint32_t tot = 0;
      FSPH_CUDA_OK(cudaMemcpyAsync(&tot, fl + nsite, 4, cudaMemcpyDeviceToHost, s->stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
      n = tot;
    }
    if(fsph_resize(s, n)) throw std::runtime_error(s->err);
    if(n > 0) {
      // the scan was exclusive: off[s] is the slot of an accepted site, accepted <=> off[s + 1] != off[s]; the emit pass
      // re-evaluates the (cheap) test instead
      FSPH_LAUNCH((k_lattice<1>), nb, 256, s->stream, T, nsite, nullptr, fl, s->cols(), id_offset);
      FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    FSPH_CUDA_OK(cudaGetLastError());
    dtab.release(); itab.release(); flags.release();
    if(n_out) *n_out = n;
    return 0;
  }
  catch(const std::exception & e) {
    s->err = e.what();
    return 1;
  }
}

// number of bodies with |x - c| < radius (the blast region of sedov/main.cc:283-295)
int32_t fsph_count_within(fsph_system * s, const double * center, double radius, int64_t * out) {
  if(!s) return 1;
  try {
    FSPH_CUDA_OK(cudaSetDevice(s->device));
    unsigned long long * d = reinterpret_cast<unsigned long long *>(s->b_status.as<uint32_t>() + 16);
    FSPH_CUDA_OK(cudaMemsetAsync(d, 0, 8, s->stream));
    if(s->n > 0) FSPH_LAUNCH(k_count_within, (unsigned)((s->n + 255) / 256), 256, s->stream, s->cols(), s->n, center[0], center[1], center[2], radius, d);
    unsigned long long h = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s->stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    *out = (int64_t)h;
    return 0;
  }
  catch(const std::exception & e) {
    s->err = e.what();
    return 1;
  }
}

// uniform fields of a lattice body set: f = {mass, h, rho, u, u_blast_add, blast_radius, gamma, dt, radial_velocity, rho_eos,
// pressure}; P = rho_eos u (gamma - 1) or the constant `pressure` (when not NaN), u += u_blast_add inside the blast radius,
// v = radial_velocity r / |r|
int32_t fsph_id_fill(fsph_system * s, const double * f11) {
  if(!s) return 1;
  try {
    FSPH_CUDA_OK(cudaSetDevice(s->device));
    IdFill f{f11[0], f11[1], f11[2], f11[3], f11[4], f11[5], f11[6], f11[7], f11[8], f11[9], f11[10]};
    if(s->n > 0) FSPH_LAUNCH(k_id_fill, (unsigned)((s->n + 255) / 256), 256, s->stream, s->cols(), s->n, f);
    FSPH_CUDA_OK(cudaStreamSynchronize(s->stream));
    FSPH_CUDA_OK(cudaGetLastError());
    s->tree_valid = false;
    s->search_valid = false;
    return 0;
  }
  catch(const std::exception & e) {
    s->err = e.what();
    return 1;
  }
}

}  // extern "C"
