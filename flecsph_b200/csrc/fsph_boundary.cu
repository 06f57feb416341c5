// fsph_boundary.cu -- periodic boundaries through WALL copies (compiled with -fmad=false: the copies' coordinates
// feed the range and the keys, which are bit-exact).
//
// Replaces boundary::pboundary_clean / pboundary_generate (include/physics/boundary.h:45-248) as called by
// body_system::update_iteration (mpisph/bodies_system.h:160-169):
//   1. drop the WALL copies of the previous iteration (stable: the other bodies keep their order);
//   2. halo = 2.5 x the largest h (bodies_system.h:164-165);
//   3. wrap bodies that left the box back into it (boundary.h:141-158);
//   4. every body within `halo` of a periodic face gets a WALL copy shifted by the box size along that axis, plus
//      copies for the edges xy, xz, zy and the corner xyz it is close to (boundary.h:160-210, corner coordinates by
//      the reference's own expression box[0] - |x + box[0]| / box[1] + |x + box[1]|, :73-80); the copies are appended
//      in the reference's order (body order, then x, y, z, xy, xz, zy, xyz) and numbered from the number of real bodies
//      upwards (:229-243), so ids agree with the reference.
// The copies are ordinary bodies from then on (they are swept and integrated like the rest, exactly as in the
// reference); the analysis reductions skip them by type (analysis.h:60).
#include <stdexcept>
#include <algorithm>
#include "fsph_internal.h"

namespace fsph {

constexpr int32_t TYPE_WALL = 1;  // particle_type_t::WALL, include/physics/body.h:35

struct PBox {
  double lo[3], hi[3], halo;
  int periodic[3];
};

struct CopyCols {
  uint64_t * p8[NF64 + 4];
  uint32_t * p4[4];
  int n8, n4;
};

__global__ void __launch_bounds__(256) k_not_wall(const int32_t * __restrict__ type, int64_t n, uint32_t * __restrict__ f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) f[i] = (type[i] & 0xff) == TYPE_WALL ? 0u : 1u;
}
__global__ void __launch_bounds__(256) k_compact_keep(const uint32_t * __restrict__ f, const uint32_t * __restrict__ ex, int64_t n,
  uint32_t * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && f[i]) out[ex[i]] = (uint32_t)i;
}

template<int D>
__global__ void __launch_bounds__(256) k_teleport(Cols c, int64_t n, PBox b) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    if(!b.periodic[d]) continue;
    double x = c.f[F_POS + d][i];
    if(x > b.hi[d]) c.f[F_POS + d][i] = x - (b.hi[d] - b.lo[d]);
    else if(x <= b.lo[d]) c.f[F_POS + d][i] = x + (b.hi[d] - b.lo[d]);
  }
}

// bit d of the result: the body is within the halo of a periodic face of axis d
template<int D>
__device__ __forceinline__ int edge_mask(const Cols & c, int64_t i, const PBox & b, double * x) {
  int m = 0;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    x[d] = c.f[F_POS + d][i];
    if(b.periodic[d] && (x[d] + b.halo > b.hi[d] || x[d] - b.halo <= b.lo[d])) m |= 1 << d;
  }
  return m;
}
__device__ __forceinline__ int copies_of(int m, int D) {
  int k = __popc(m);
  if(D == 2) k += (m == 3);
  if(D == 3) k += ((m & 3) == 3) + ((m & 5) == 5) + ((m & 6) == 6) + (m == 7);
  return k;
}

template<int D>
__global__ void __launch_bounds__(256) k_wall_count(Cols c, int64_t n, PBox b, uint32_t * __restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  double x[3];
  cnt[i] = (uint32_t)copies_of(edge_mask<D>(c, i, b, x), D);
}

template<int D>
__global__ void __launch_bounds__(128) k_wall_emit(Cols c, int64_t n, PBox b, const uint32_t * __restrict__ off, CopyCols cc, int64_t id_base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  double x[3] = {0, 0, 0};
  const int m = edge_mask<D>(c, i, b, x);
  if(!m) return;
  int64_t dst = n + off[i];
  auto emit = [&](const double * nx) {
    for(int k = 0; k < cc.n8; ++k) cc.p8[k][dst] = cc.p8[k][i];
    for(int k = 0; k < cc.n4; ++k) cc.p4[k][dst] = cc.p4[k][i];
#pragma unroll
    for(int d = 0; d < D; ++d) c.f[F_POS + d][dst] = nx[d];
    c.type[dst] = TYPE_WALL;
    c.id[dst] = id_base + (dst - n);
    ++dst;
  };
  // one copy per axis (boundary.h:172-189)
#pragma unroll
  for(int d = 0; d < D; ++d) {
    if(!((m >> d) & 1)) continue;
    double nx[3] = {x[0], x[1], x[2]};
    if(x[d] + b.halo > b.hi[d]) nx[d] = x[d] - (b.hi[d] - b.lo[d]);
    else nx[d] = x[d] + (b.hi[d] - b.lo[d]);
    emit(nx);
  }
  // edges and corner (add_corner_XX / add_corner_XXX, boundary.h:63-118)
  auto corner = [&](int dm) {
    if((m & dm) != dm) return;
    double nx[3] = {x[0], x[1], x[2]};
#pragma unroll
    for(int d = 0; d < D; ++d) {
      if(!((dm >> d) & 1)) continue;
      if(x[d] + b.halo > b.hi[d]) nx[d] = b.lo[d] - fabs(x[d] + b.lo[d]);
      else nx[d] = b.hi[d] + fabs(x[d] + b.hi[d]);
    }
    emit(nx);
  };
  if(D == 2) corner(3);
  if(D == 3) {
    corner(3);  // x, y
    corner(5);  // x, z
    corner(6);  // z, y
    corner(7);
  }
}

// grow every column to n_new bodies; the first n_keep bodies of the columns that carry state are preserved
static void resize_keep(System & s, int64_t n_new, int64_t n_keep) {
  std::vector<std::pair<DevBuf *, int>> cols;
  live_columns(s, cols);
  for(auto & c : cols) c.first->grow_keep((size_t)n_new * c.second, (size_t)n_keep * c.second, s.stream);
  for(int f = 0; f < NF64; ++f)
    if(s.col_used[f]) s.col[f].grow_keep((size_t)n_new * 8, 0, s.stream);  // the others only need the room
  s.col_nneigh.grow_keep((size_t)n_new * 4, 0, s.stream);
  s.scratch_col.reserve((size_t)n_new * 8);
}

template<int D>
static void periodic_walls_t(System & s) {
  // several ranks: every rank copies its own bodies (the copies are owned bodies that the partition step then
  // distributes like any other); halo from the global largest h, ids from the global count of real bodies plus the
  // copy counts of the lower ranks (boundary.h:124-126, 218-243)
  if(s.has_ghosts) dist_drop_ghosts(s);
  const unsigned nb0 = (unsigned)((s.n + 255) / 256);
  // 1. pboundary_clean
  {
    const int64_t n = s.n;
    s.b_cnt0.reserve((size_t)(n + 1) * 4);
    s.b_cnt1.reserve((size_t)(n + 1) * 4);
    s.b_perm.reserve((size_t)n * 4);
    uint32_t * f = s.b_cnt0.as<uint32_t>();
    uint32_t * ex = s.b_cnt1.as<uint32_t>();
    FSPH_LAUNCH(k_not_wall, nb0, 256, s.stream, s.col_type.as<int32_t>(), n, f);
    exclusive_scan_u32(s, f, ex, n, nullptr);
    uint32_t keep = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&keep, ex + n, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(keep == 0) throw std::runtime_error("pboundary_clean: only WALL bodies (boundary.h:51)");
    if((int64_t)keep != n) {
      FSPH_LAUNCH(k_compact_keep, nb0, 256, s.stream, f, ex, n, s.b_perm.as<uint32_t>());
      gather_bodies(s, s.b_perm.as<uint32_t>(), keep);
      s.n = keep;
    }
  }
  const int64_t n = s.n;
  const unsigned nb = (unsigned)((n + 255) / 256);
  // 2. + 3.
  PBox b;
  const double ext[3] = {s.par.box_length, s.par.box_width, s.par.box_height};
  const int per[3] = {s.par.periodic_boundary_x, s.par.periodic_boundary_y, s.par.periodic_boundary_z};
  for(int d = 0; d < 3; ++d) {
    b.lo[d] = -0.5 * ext[d];
    b.hi[d] = 0.5 * ext[d];
    b.periodic[d] = d < D ? per[d] : 0;
  }
  {
    double hm[3] = {max_h(s), 0, 0};
    if(s.comm) global_combine(s, 1 /* RED_MAX_H */, hm);
    b.halo = 2.5 * hm[0];
  }
  FSPH_LAUNCH((k_teleport<D>), nb, 256, s.stream, s.cols(), n, b);
  // 4. count, scan, emit
  s.b_cnt0.reserve((size_t)(n + 1) * 4);
  uint32_t * cnt = s.b_cnt0.as<uint32_t>();
  FSPH_LAUNCH((k_wall_count<D>), nb, 256, s.stream, s.cols(), n, b, cnt);
  exclusive_scan_u32(s, cnt, cnt, n, nullptr);
  uint32_t nwall = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&nwall, cnt + n, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(n + (int64_t)nwall >= (int64_t(1) << 31)) throw std::runtime_error("bodies + WALL copies exceed 2^31");
  int64_t id_base = n;
  if(s.comm) {
    const int P = s.comm->nranks, R = s.comm->rank;
    int64_t mine[2] = {n, (int64_t)nwall};
    std::vector<int64_t> all((size_t)2 * P);
    dist_all_gather_host(s, mine, all.data(), sizeof(mine));
    id_base = 0;
    for(int q = 0; q < P; ++q) id_base += all[(size_t)2 * q];
    for(int q = 0; q < R; ++q) id_base += all[(size_t)2 * q + 1];
  }
  if(nwall) {
    resize_keep(s, n + nwall, n);
    std::vector<std::pair<DevBuf *, int>> cols;
    live_columns(s, cols);
    CopyCols cc;
    cc.n8 = cc.n4 = 0;
    for(auto & c : cols) {
      if(c.second == 8) cc.p8[cc.n8++] = c.first->as<uint64_t>();
      else cc.p4[cc.n4++] = c.first->as<uint32_t>();
    }
    FSPH_LAUNCH((k_wall_emit<D>), (unsigned)((n + 127) / 128), 128, s.stream, s.cols(), n, b, cnt, cc, id_base);
    s.n = n + nwall;
  }
  s.n_owned = s.n;
  s.tree_valid = false;
  s.search_valid = false;
  FSPH_CUDA_OK(cudaGetLastError());
}

void periodic_walls(System & s) {
  if(s.dim == 1) throw std::runtime_error("periodic boundaries need 2 or 3 dimensions (boundary.h:128-137 indexes y in 1-D)");
  if(s.dim == 2) periodic_walls_t<2>(s);
  else periodic_walls_t<3>(s);
}

}  // namespace fsph
