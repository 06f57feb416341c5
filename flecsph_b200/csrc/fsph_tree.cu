// fsph_tree.cu -- range, Morton keys, sort + column gather, bottom-up octree, node summaries, work cells.
// Compiled with -fmad=false: every fp64 operation here is a separate IEEE operation in the reference's
// order, because keys, node boxes / centres / radii feed bit-exact decisions.
//
// Replaces (reference file:line):
//   body_system::mpi_compute_range        mpisph/bodies_system.h:231-268
//   morton_curve_u(range, p)              include/tree_topology/filling_curve.h:503-525
//   tree_topology::compute_keys           include/tree_topology/tree_topology.h:828-832
//   psort::psort (one rank)               mpisph/psort.h:227-322, comparator bodies_system.h:197-205
//   tree_topology::build_tree / finish_   include/tree_topology/tree_topology.h:878-997,1893-1918
//   physics::compute_cofm                 include/physics/default_physics.h:59-119
//   traversal_sph work-cell collection    include/tree_topology/tree_topology.h:221-235
//
// Tree layout: the reference inserts key-sorted particles one by one into an unordered_map; the result is
// the canonical octree (a node for every key prefix shared by >= 2 particles, every particle a leaf at the
// shallowest depth where it is alone).  That is a pure function of the longest-common-prefix depths of
// adjacent sorted keys, so it is built here without insertion: particle i opens the nodes of depths
// (lcp[i-1], lcp[i]] and one leaf at depth max(lcp[i-1], lcp[i]) + 1; a stable counting sort by depth puts
// all cells in level order, where the children of consecutive nodes are consecutive cells.
#include <stdexcept>
#include <cfloat>
#include "fsph_internal.h"

namespace fsph {

template<int D> struct KeyBits { static constexpr int B = 63 / D; };  // filling_curve.h:45-47

// ---- range --------------------------------------------------------------------------------------------
template<int D>
__global__ void __launch_bounds__(256) k_range_partial(Cols c, int64_t n, double * __restrict__ part) {
  double mn[D], mx[D];
#pragma unroll
  for(int d = 0; d < D; ++d) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
  for(int64_t i = c.own_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.own_hi; i += (int64_t)gridDim.x * blockDim.x) {
    double h = c.f[F_H][i];
#pragma unroll
    for(int d = 0; d < D; ++d) {
      double x = c.f[F_POS + d][i];
      mn[d] = fmin(mn[d], x - h);
      mx[d] = fmax(mx[d], x + h);
    }
  }
  __shared__ double sm[8][2 * D];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for(int d = 0; d < D; ++d) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if(lane == 0) { sm[w][d] = mn[d]; sm[w][D + d] = mx[d]; }
  }
  __syncthreads();
  if(threadIdx.x < 2 * D) {
    bool ismin = threadIdx.x < D;
    double v = sm[0][threadIdx.x];
    for(int ww = 1; ww < 8; ++ww) v = ismin ? fmin(v, sm[ww][threadIdx.x]) : fmax(v, sm[ww][threadIdx.x]);
    part[(size_t)blockIdx.x * 2 * D + threadIdx.x] = v;
  }
}
template<int D>
__global__ void k_range_final(const double * __restrict__ part, int nb, double * __restrict__ out) {
  if(threadIdx.x < 2 * D) {
    bool ismin = threadIdx.x < D;
    double v = part[threadIdx.x];
    for(int b = 1; b < nb; ++b) {
      double u = part[(size_t)b * 2 * D + threadIdx.x];
      v = ismin ? fmin(v, u) : fmax(v, u);
    }
    out[threadIdx.x] = v;
  }
}

double * reduce_scratch(System & s, size_t n) {
  s.b_red.reserve(n * sizeof(double));
  return s.b_red.as<double>();
}

template<int D>
static void compute_range_t(System & s) {
  int nb = (int)std::min<int64_t>((s.n + 255) / 256, 148 * 4);
  double * part = reduce_scratch(s, (size_t)(nb + 1) * 2 * D + 16);
  double * out = part + (size_t)nb * 2 * D;
  FSPH_LAUNCH((k_range_partial<D>), nb, 256, s.stream, s.cols(), s.n, part);
  FSPH_LAUNCH((k_range_final<D>), 1, 32, s.stream, part, nb, out);
  double h[2 * D];
  FSPH_CUDA_OK(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  bool same = true;
  for(int d = 0; d < D; ++d) {
    s.lo[d] = h[d];
    s.hi[d] = h[D + d];
    if(s.lo[d] != s.hi[d]) same = false;
  }
  if(same) throw std::runtime_error("Range are equals (bodies_system.h:175)");
}
void compute_range(System & s) {
  if(s.n <= 0) throw std::runtime_error("no bodies");
  if(s.dim == 1) compute_range_t<1>(s);
  else if(s.dim == 2) compute_range_t<2>(s);
  else compute_range_t<3>(s);
}

// ---- keys ---------------------------------------------------------------------------------------------
struct Range3 { double lo[3], hi[3]; };

template<int D>
__global__ void __launch_bounds__(256) k_keys(Cols c, int64_t n, Range3 r) {
  constexpr int B = KeyBits<D>::B;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  if(i < c.own_lo || i >= c.own_hi) {
    c.key[i] = ~0ull;  // last in key order: the next sort drops the remote copies of the previous iteration
    return;
  }
  const uint64_t maxv = (uint64_t(1) << B) - 1;
  const double two_b = (double)(uint64_t(1) << B);
  uint64_t cd[D];
#pragma unroll
  for(int d = 0; d < D; ++d) {
    double mn = r.lo[d];
    double scale = r.hi[d] - mn;
    double t = (c.f[F_POS + d][i] - mn) / scale * two_b;  // sub, div, mul: filling_curve.h:513-515
    uint64_t v = (uint64_t)t;                              // truncation (cvt.rzi.u64.f64)
    cd[d] = v < maxv ? v : maxv;
  }
  uint64_t key = uint64_t(1) << (D * B);
  for(int b = 0; b < B; ++b) {
#pragma unroll
    for(int d = 0; d < D; ++d) key |= ((cd[d] >> b) & 1ull) << (b * D + d);
  }
  c.key[i] = key;
}

void compute_keys(System & s) {
  Range3 r;
  for(int d = 0; d < 3; ++d) { r.lo[d] = s.lo[d]; r.hi[d] = s.hi[d]; }
  unsigned nb = (unsigned)((s.n + 255) / 256);
  if(s.dim == 1) FSPH_LAUNCH((k_keys<1>), nb, 256, s.stream, s.cols(), s.n, r);
  else if(s.dim == 2) FSPH_LAUNCH((k_keys<2>), nb, 256, s.stream, s.cols(), s.n, r);
  else FSPH_LAUNCH((k_keys<3>), nb, 256, s.stream, s.cols(), s.n, r);
  FSPH_CUDA_OK(cudaGetLastError());
}

// ---- sort + gather --------------------------------------------------------------------------------------
template<typename T>
__global__ void __launch_bounds__(256) k_gather(const T * __restrict__ in, T * __restrict__ out, const uint32_t * __restrict__ perm,
  int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = in[perm[i]];
}
__global__ void __launch_bounds__(256) k_is_identity(const uint32_t * __restrict__ perm, int64_t n, uint32_t * flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && perm[i] != (uint32_t)i) *flag = 1;
}

static void gather4(System & s, DevBuf & colbuf, const uint32_t * perm, int64_t n_new) {
  s.b_tmp[0].reserve_exact(std::max(colbuf.cap, (size_t)n_new * 4));
  unsigned nb = (unsigned)((n_new + 255) / 256);
  if(n_new > 0) FSPH_LAUNCH((k_gather<uint32_t>), nb, 256, s.stream, colbuf.as<uint32_t>(), s.b_tmp[0].as<uint32_t>(), perm, n_new);
  std::swap(colbuf.p, s.b_tmp[0].p);
  std::swap(colbuf.cap, s.b_tmp[0].cap);
}

// The columns that carry state across a sort / an exchange, with their element size.
void live_columns(System & s, std::vector<std::pair<DevBuf *, int>> & out) {
  out.clear();
  for(int f = 0; f < NF64; ++f) {
    if(f < F_MASS && (f % 3) >= s.dim) continue;  // unused vector components
    if(f == F_PRHO2) continue;                      // derived column
    if(!s.col_used[f]) continue;                    // optional column this physics does not allocate
    if(s.skip_dead_columns && s.dead_col[f]) continue;  // fsph_step: rewritten before it is next read
    out.emplace_back(&s.col[f], 8);
  }
  out.emplace_back(&s.col_id, 8);
  out.emplace_back(&s.col_key, 8);
  out.emplace_back(&s.col_type, 4);
  if(s.comm) {
    out.emplace_back(&s.col_gd, 4);
    out.emplace_back(&s.col_nneigh, 4);  // stays aligned with the owned bodies when the ghosts are dropped
  }
}

// Several 8-byte columns gathered in one pass: the permutation is read once per body for all of them.
struct GatherSet {
  const uint64_t * in[8];
  uint64_t * out[8];
  int n;
};
__global__ void __launch_bounds__(256) k_gather_set(GatherSet g, const uint32_t * __restrict__ perm, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint32_t p = perm[i];
  uint64_t v[8];
#pragma unroll
  for(int j = 0; j < 8; ++j)
    if(j < g.n) v[j] = g.in[j][p];
#pragma unroll
  for(int j = 0; j < 8; ++j)
    if(j < g.n) g.out[j][i] = v[j];
}

// The columns a remote copy needs when it ARRIVES (fsph_dist.cu, halo exchange inside fsph_step): what the tree, the
// neighbour search and the density sweep read from neighbours -- position, h, mass -- plus identity.  Everything else a
// sweep reads from a neighbour (rho, P, c_s, v, v_1/2, alpha, div v) is re-sent by reset_ghosts before that sweep.
void halo_columns(System & s, std::vector<std::pair<DevBuf *, int>> & out) {
  out.clear();
  for(int d = 0; d < s.dim; ++d) out.emplace_back(&s.col[F_POS + d], 8);
  out.emplace_back(&s.col[F_H], 8);
  out.emplace_back(&s.col[F_MASS], 8);
  out.emplace_back(&s.col_id, 8);
  out.emplace_back(&s.col_key, 8);
  out.emplace_back(&s.col_type, 4);
  out.emplace_back(&s.col_gd, 4);
}

// out[i] = in[perm[i]] for every live column; perm must not live in scratch_col / b_tmp[0].  The 8-byte columns go
// through a ring of spare buffers, eight columns per launch: column j's old storage becomes the spare of a later one.
void gather_bodies(System & s, const uint32_t * perm, int64_t n_new) {
  std::vector<std::pair<DevBuf *, int>> cols;
  live_columns(s, cols);
  std::vector<DevBuf *> c8;
  for(auto & c : cols) {
    if(c.second == 8) c8.push_back(c.first);
    else gather4(s, *c.first, perm, n_new);
  }
  const unsigned nb = (unsigned)((n_new + 255) / 256);
  size_t need = (size_t)n_new * 8;
  for(DevBuf * c : c8) need = std::max(need, c->cap);
  for(size_t b = 0; b < c8.size(); b += 8) {
    const int k = (int)std::min<size_t>(8, c8.size() - b);
    // spares for this batch: scratch_col and the gather pool
    DevBuf * spare[8];
    for(int j = 0; j < k; ++j) {
      spare[j] = j == 0 ? &s.scratch_col : &s.b_gather[j - 1];
      spare[j]->reserve_exact(need);
    }
    GatherSet g;
    g.n = k;
    for(int j = 0; j < 8; ++j) { g.in[j] = nullptr; g.out[j] = nullptr; }
    for(int j = 0; j < k; ++j) { g.in[j] = c8[b + j]->as<uint64_t>(); g.out[j] = spare[j]->as<uint64_t>(); }
    if(n_new > 0) FSPH_LAUNCH(k_gather_set, nb, 256, s.stream, g, perm, n_new);
    for(int j = 0; j < k; ++j) {
      std::swap(c8[b + j]->p, spare[j]->p);
      std::swap(c8[b + j]->cap, spare[j]->cap);
    }
  }
  FSPH_CUDA_OK(cudaGetLastError());
}

void sort_bodies(System & s, int64_t n_keep) {
  const int64_t n = s.n;
  if(n_keep < 0) n_keep = n;
  s.b_keys2.reserve((size_t)n * 8);
  s.b_perm.reserve((size_t)n * 4);
  s.b_perm2.reserve((size_t)n * 4);
  s.scratch_col.reserve((size_t)n * 8);
  // sort a copy of the keys so that the key column itself is gathered like every other column
  uint64_t * k0 = s.scratch_col.as<uint64_t>();
  uint64_t * k1 = s.b_keys2.as<uint64_t>();
  uint32_t * v0 = s.b_perm.as<uint32_t>();
  uint32_t * v1 = s.b_perm2.as<uint32_t>();
  FSPH_CUDA_OK(cudaMemcpyAsync(k0, s.col_key.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, s.stream));
  radix_sort_pairs(s, k0, v0, k1, v1, n, s.col_key.as<uint64_t>());
  // the sorted permutation must not live in scratch_col (the gathers overwrite it)
  const uint32_t * perm = v0;
  s.last_perm = perm;
  // skip the gathers when the bodies are already in key order
  uint32_t * flag = s.b_status.as<uint32_t>() + 1;
  FSPH_CUDA_OK(cudaMemsetAsync(flag, 0, 4, s.stream));
  unsigned nb = (unsigned)((n_keep + 255) / 256);
  if(n_keep > 0) FSPH_LAUNCH(k_is_identity, nb, 256, s.stream, perm, n_keep, flag);
  uint32_t moved = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&moved, flag, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(moved) gather_bodies(s, perm, n_keep);
  s.n = n_keep;
  FSPH_CUDA_OK(cudaGetLastError());
}

// ---- tree topology ------------------------------------------------------------------------------------
template<int D>
__global__ void __launch_bounds__(256) k_lcp(const uint64_t * __restrict__ key, int64_t n, int8_t * __restrict__ lcp, uint32_t * status) {
  constexpr int B = KeyBits<D>::B;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  int d = -1;
  if(i + 1 < n) {
    uint64_t x = key[i] ^ key[i + 1];
    if(x == 0) {
      atomicOr(status, (uint32_t)ERR_DUP_KEYS);
      d = B - 1;
    }
    else {
      int p = 63 - __clzll((long long)x);
      d = B - 1 - p / D;
    }
  }
  lcp[i] = (int8_t)d;
}

// Leaf depth of every body in the global tree: one below the deepest node it shares with a key neighbour (the root
// always exists, tree_topology.h:893).  On several ranks the neighbours of the first / last body live on the adjacent
// ranks (kb / ka), and for a tree that also holds ghosts the depth is given (computed by each body's owner).
template<int D>
__global__ void __launch_bounds__(256) k_leaf_depth(const int8_t * __restrict__ lcp, int64_t n, const uint64_t * __restrict__ key, int has_b,
  uint64_t kb, int has_a, uint64_t ka, int32_t * __restrict__ gd) {
  constexpr int B = KeyBits<D>::B;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  int dp = i > 0 ? lcp[i - 1] : -1, dc = i + 1 < n ? lcp[i] : -1;
  if(i == 0 && has_b) {
    uint64_t x = key[0] ^ kb;
    dp = x ? B - 1 - (63 - __clzll((long long)x)) / D : B - 1;
  }
  if(i == n - 1 && has_a) {
    uint64_t x = key[i] ^ ka;
    dc = x ? B - 1 - (63 - __clzll((long long)x)) / D : B - 1;
  }
  int d = (dp > dc ? dp : dc) + 1;
  gd[i] = d < 1 ? 1 : d;
}

// body i opens the nodes of depths (lcp[i-1], gd[i] - 1] and one leaf at depth gd[i]
__global__ void __launch_bounds__(256) k_cell_count(const int8_t * __restrict__ lcp, const int32_t * __restrict__ gd, int64_t n,
  uint32_t * __restrict__ cnt, uint32_t * status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int dp = i > 0 ? lcp[i - 1] : -1, dc = i + 1 < n ? lcp[i] : -1, g = gd[i];
  if(g - 1 < dp || g - 1 < dc || g > 63) {
    atomicOr(status, (uint32_t)ERR_DD_CAPACITY);  // inconsistent global leaf depth (cannot happen for a sub-set of one tree)
    cnt[i] = 1;
    return;
  }
  cnt[i] = (uint32_t)(g - dp);
}

__global__ void __launch_bounds__(256) k_cell_emit(const int8_t * __restrict__ lcp, const int32_t * __restrict__ gd, int64_t n,
  const uint32_t * __restrict__ off, uint64_t * __restrict__ desc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int dp = i > 0 ? lcp[i - 1] : -1, g = gd[i];
  if(g - 1 < dp) return;
  uint32_t o = off[i];
  uint64_t lowbits = (uint64_t)i << 1;
  for(int k = dp + 1; k <= g - 1; ++k) desc[o++] = ((uint64_t)k << 32) | lowbits;
  desc[o] = ((uint64_t)g << 32) | lowbits | 1ull;
}

template<int D>
__global__ void __launch_bounds__(256) k_cell_flags(const uint64_t * __restrict__ desc, int64_t nc, const uint64_t * __restrict__ key,
  uint64_t * __restrict__ flags) {
  constexpr int B = KeyBits<D>::B;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= nc) return;
  uint64_t dsc = desc[c];
  int depth = (int)(dsc >> 32);
  uint64_t isnode = (dsc & 1ull) ? 0 : 1;
  uint64_t head = 0;
  if(c > 0) {
    uint64_t prev = desc[c - 1];
    int pdepth = (int)(prev >> 32);
    if(pdepth != depth) head = 1;
    else {
      int sh = D * (B - (depth - 1));
      uint64_t a = key[(uint32_t)dsc >> 1] >> sh, b = key[(uint32_t)prev >> 1] >> sh;
      head = a != b;
    }
  }
  flags[c] = isnode | (head << 32);
}

__global__ void __launch_bounds__(256) k_cell_link(const uint64_t * __restrict__ desc, const uint64_t * __restrict__ flags,
  const uint64_t * __restrict__ ex, int64_t nc, Tree t, int32_t * __restrict__ level_node_begin) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= nc) return;
  uint64_t dsc = desc[c], fl = flags[c], e = ex[c];
  int depth = (int)(dsc >> 32);
  int start = (int)((uint32_t)dsc >> 1);
  bool isnode = fl & 1ull, head = (fl >> 32) & 1ull;
  int nr = (int)(uint32_t)e;
  int parent = (int)(uint32_t)(e >> 32) + (head ? 1 : 0) - 1;
  t.cell_parent[c] = parent;
  if(isnode) {
    t.cell[c] = ~nr;
    t.node_start[nr] = start;
    t.node_depth[nr] = depth;
    t.node_parent[nr] = parent;
  }
  else t.cell[c] = start;
  if(head) t.child_begin[parent] = (int)c;
  if(c > 0) {
    bool last = (c + 1 == nc) || ((flags[c + 1] >> 32) & 1ull);
    if(last) t.child_end[parent] = (int)c + 1;
  }
  if(c == 0 || (int)(desc[c - 1] >> 32) != depth) level_node_begin[depth] = nr;
}

template<int D>
__device__ __forceinline__ double dist_pts(const double * a, const double * b) {
  // flecsi::distance, space_vector.h:46-58
  if(D == 1) return fabs(a[0] - b[0]);
  double sum = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) sum = sum + (a[d] - b[d]) * (a[d] - b[d]);
  return sqrt(sum);
}

// physics::compute_cofm (default_physics.h:59-119) for the nodes [nb, ne) of one level: entity children first,
// then node children, each in child-index order (tree_topology.h:1954-1968).
template<int D>
__global__ void __launch_bounds__(128) k_cofm(Tree t, Cols c, int nb, int ne) {
  int nd = nb + blockIdx.x * blockDim.x + threadIdx.x;
  if(nd >= ne) return;
  const int cb = t.child_begin[nd], ce = t.child_end[nd];
  double X[D], mn[D], mx[D];
  double M = 0.0;
  int S = 0;
#pragma unroll
  for(int d = 0; d < D; ++d) { X[d] = 0.0; mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
  for(int cc = cb; cc < ce; ++cc) {
    int v = t.cell[cc];
    if(v < 0) continue;
    double m = c.f[F_MASS][v], h = c.f[F_H][v];
#pragma unroll
    for(int d = 0; d < D; ++d) {
      double x = c.f[F_POS + d][v];
      X[d] = X[d] + m * x;
      mn[d] = fmin(mn[d], x - h / 2.);
      mx[d] = fmax(mx[d], x + h / 2.);
    }
    M = M + m;
    ++S;
  }
  for(int cc = cb; cc < ce; ++cc) {
    int v = t.cell[cc];
    if(v >= 0) continue;
    int ch = ~v;
    double m = t.mass[ch];
#pragma unroll
    for(int d = 0; d < D; ++d) {
      X[d] = X[d] + m * t.com[d][ch];
      mn[d] = fmin(mn[d], t.bmin[d][ch]);
      mx[d] = fmax(mx[d], t.bmax[d][ch]);
    }
    M = M + m;
    S += t.sub[ch];
  }
#pragma unroll
  for(int d = 0; d < D; ++d) X[d] = X[d] / M;
  double radius = 0.0, lap = 0.0;
  for(int cc = cb; cc < ce; ++cc) {
    int v = t.cell[cc];
    if(v < 0) continue;
    double p[D];
#pragma unroll
    for(int d = 0; d < D; ++d) p[d] = c.f[F_POS + d][v];
    double dist = dist_pts<D>(X, p);
    radius = fmax(radius, dist);
    lap = fmax(lap, dist + c.f[F_H][v]);
  }
  for(int cc = cb; cc < ce; ++cc) {
    int v = t.cell[cc];
    if(v >= 0) continue;
    int ch = ~v;
    double p[D];
#pragma unroll
    for(int d = 0; d < D; ++d) p[d] = t.com[d][ch];
    double dist = dist_pts<D>(X, p);
    double rc = t.radius[ch];
    radius = fmax(radius, dist + rc);
    lap = fmax(lap, dist + rc + t.lap[ch]);
  }
  t.mass[nd] = M;
  t.sub[nd] = S;
  t.radius[nd] = radius;
  t.lap[nd] = lap;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    t.com[d][nd] = X[d];
    t.bmin[d][nd] = mn[d];
    t.bmax[d][nd] = mx[d];
  }
}

// packed walk records (fsph_search.cu) from the SoA tables
template<int D>
__global__ void __launch_bounds__(256) k_pack_nodes(Tree t, int nn) {
  int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if(nd >= nn) return;
  NodeRec r;
#pragma unroll
  for(int d = 0; d < 3; ++d) {
    r.bmin[d] = d < D ? t.bmin[d][nd] : 0.0;
    r.bmax[d] = d < D ? t.bmax[d][nd] : 0.0;
  }
  const int cb = t.child_begin[nd], ce = t.child_end[nd];
  uint32_t leafmask = 0;
  int first = -1;
  for(int cc = cb; cc < ce; ++cc) {
    const int v = t.cell[cc];
    if(v >= 0) leafmask |= 1u << (cc - cb);
    else if(first < 0) first = ~v;
  }
  r.child_begin = cb;
  r.first_child_node = first;
  r.nchild = (uint32_t)(ce - cb);
  r.leafmask = leafmask;
  t.noderec[nd] = r;
}
template<int D>
__global__ void __launch_bounds__(256) k_pack_leaves(Tree t, Cols c, int64_t nc) {
  int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(cc >= nc) return;
  const int v = t.cell[cc];
  if(v < 0) return;
  LeafRec r;
#pragma unroll
  for(int d = 0; d < 3; ++d) r.x[d] = d < D ? c.f[F_POS + d][v] : 0.0;
  r.h = c.f[F_H][v];
  t.leafrec[cc] = r;
}

template<int D>
static void build_tree_t(System & s, const TreeEdges & te) {
  const int64_t n = s.n;
  unsigned nbp = (unsigned)((n + 255) / 256);
  uint32_t * status = s.b_status.as<uint32_t>();
  s.b_tmp[0].reserve((size_t)n);                      // lcp (int8)
  s.b_cnt0.reserve((size_t)(n + 1) * sizeof(uint32_t));
  int8_t * lcp = s.b_tmp[0].as<int8_t>();
  uint32_t * cnt = s.b_cnt0.as<uint32_t>();
  FSPH_LAUNCH((k_lcp<D>), nbp, 256, s.stream, s.col_key.as<uint64_t>(), n, lcp, status);
  const int32_t * gd = te.gd_in;
  if(!gd) {
    int32_t * g = te.gd_out;
    if(!g) {
      s.b_tmp[1].reserve((size_t)n * 4);
      g = s.b_tmp[1].as<int32_t>();
    }
    FSPH_LAUNCH((k_leaf_depth<D>), nbp, 256, s.stream, lcp, n, s.col_key.as<uint64_t>(), te.has_before ? 1 : 0, te.key_before,
      te.has_after ? 1 : 0, te.key_after, g);
    gd = g;
  }
  FSPH_LAUNCH(k_cell_count, nbp, 256, s.stream, lcp, gd, n, cnt, status);
  exclusive_scan_u32(s, cnt, cnt, n, nullptr);
  uint32_t h2[2];
  FSPH_CUDA_OK(cudaMemcpyAsync(&h2[0], cnt + n, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaMemcpyAsync(&h2[1], status, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(h2[1] & ERR_DUP_KEYS) throw std::runtime_error("two bodies have the same key (the reference's build_tree breaks on this too)");
  if(h2[1] & ERR_DD_CAPACITY) throw std::runtime_error("inconsistent global leaf depths in the merged tree");
  const int64_t nc = h2[0];
  const int64_t nn = nc - n;
  if(nc >= (int64_t(1) << 31)) throw std::runtime_error("tree has more than 2^31 cells");
  s.b_cell_desc.reserve((size_t)nc * 8);
  s.b_cell_desc2.reserve((size_t)nc * 8);
  uint64_t * desc0 = s.b_cell_desc2.as<uint64_t>();
  uint64_t * desc = s.b_cell_desc.as<uint64_t>();
  FSPH_LAUNCH(k_cell_emit, nbp, 256, s.stream, lcp, gd, n, cnt, desc0);
  radix_pass(s, desc0, nullptr, desc, nullptr, nc, 32);  // stable counting sort by depth -> level order
  // flags + scan (node rank in the low word, number of first-children in the high word)
  uint64_t * flags = desc0;  // reuse
  s.b_cnt1.reserve((size_t)(nc + 1) * 8);
  uint64_t * ex = s.b_cnt1.as<uint64_t>();
  unsigned nbc = (unsigned)((nc + 255) / 256);
  FSPH_LAUNCH((k_cell_flags<D>), nbc, 256, s.stream, desc, nc, s.col_key.as<uint64_t>(), flags);
  exclusive_scan_u64(s, flags, ex, nc, nullptr);
  // allocate the tables
  s.b_cell.reserve((size_t)nc * 4);
  s.b_cell_parent.reserve((size_t)nc * 4);
  s.b_node_parent.reserve((size_t)std::max<int64_t>(nn, 1) * 4);
  size_t nn4 = (size_t)std::max<int64_t>(nn, 1) * 4, nn8 = (size_t)std::max<int64_t>(nn, 1) * 8;
  s.b_child_begin.reserve(nn4); s.b_child_end.reserve(nn4); s.b_node_start.reserve(nn4); s.b_node_depth.reserve(nn4); s.b_sub.reserve(nn4);
  s.b_nmass.reserve(nn8); s.b_nradius.reserve(nn8); s.b_nlap.reserve(nn8);
  for(int d = 0; d < 3; ++d) { s.b_ncom[d].reserve(nn8); s.b_nbmin[d].reserve(nn8); s.b_nbmax[d].reserve(nn8); }
  Tree & t = s.tree;
  t.cell = s.b_cell.as<int32_t>(); t.cell_parent = s.b_cell_parent.as<int32_t>(); t.cell_desc = desc;
  t.node_parent = s.b_node_parent.as<int32_t>();
  t.child_begin = s.b_child_begin.as<int32_t>(); t.child_end = s.b_child_end.as<int32_t>();
  t.node_start = s.b_node_start.as<int32_t>(); t.node_depth = s.b_node_depth.as<int32_t>(); t.sub = s.b_sub.as<int32_t>();
  t.mass = s.b_nmass.as<double>(); t.radius = s.b_nradius.as<double>(); t.lap = s.b_nlap.as<double>();
  for(int d = 0; d < 3; ++d) { t.com[d] = s.b_ncom[d].as<double>(); t.bmin[d] = s.b_nbmin[d].as<double>(); t.bmax[d] = s.b_nbmax[d].as<double>(); }
  t.ncells = (int)nc;
  t.nnodes = (int)nn;
  s.b_noderec.reserve((size_t)std::max<int64_t>(nn, 1) * sizeof(NodeRec));
  s.b_leafrec.reserve((size_t)nc * sizeof(LeafRec));
  t.noderec = s.b_noderec.as<NodeRec>();
  t.leafrec = s.b_leafrec.as<LeafRec>();
  s.b_levels.reserve(MAX_LEVELS * sizeof(int32_t));
  int32_t * lv = s.b_levels.as<int32_t>();
  FSPH_CUDA_OK(cudaMemsetAsync(lv, 0xff, MAX_LEVELS * sizeof(int32_t), s.stream));
  FSPH_LAUNCH(k_cell_link, nbc, 256, s.stream, desc, flags, ex, nc, t, lv);
  int32_t hl[MAX_LEVELS];
  FSPH_CUDA_OK(cudaMemcpyAsync(hl, lv, sizeof(hl), cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  int maxlev = 0;
  for(int k = 0; k < MAX_LEVELS; ++k)
    if(hl[k] >= 0) maxlev = k;
  s.nlevels = maxlev + 1;
  s.level_node_begin[s.nlevels] = (int)nn;
  for(int k = maxlev; k >= 0; --k) s.level_node_begin[k] = hl[k] >= 0 ? hl[k] : s.level_node_begin[k + 1];
  // upward pass, deepest level first
  for(int k = maxlev; k >= 0; --k) {
    int b = s.level_node_begin[k], e = s.level_node_begin[k + 1];
    if(e <= b) continue;
    FSPH_LAUNCH((k_cofm<D>), (e - b + 127) / 128, 128, s.stream, t, s.cols(), b, e);
  }
  FSPH_CUDA_OK(cudaGetLastError());
}

template<int D>
static void pack_tree_t(System & s) {
  Tree & t = s.tree;
  FSPH_LAUNCH((k_pack_nodes<D>), (unsigned)((t.nnodes + 255) / 256), 256, s.stream, t, t.nnodes);
  FSPH_LAUNCH((k_pack_leaves<D>), (unsigned)((t.ncells + 255) / 256), 256, s.stream, t, s.cols(), (int64_t)t.ncells);
  FSPH_CUDA_OK(cudaGetLastError());
}

void finish_tree(System & s) {
  if(s.dim == 1) pack_tree_t<1>(s);
  else if(s.dim == 2) pack_tree_t<2>(s);
  else pack_tree_t<3>(s);
  find_work_cells(s);
}

void build_tree_edges(System & s, const TreeEdges & e) {
  s.last_perm = nullptr;  // the tree build reuses the sort's scratch
  if(s.dim == 1) build_tree_t<1>(s, e);
  else if(s.dim == 2) build_tree_t<2>(s, e);
  else build_tree_t<3>(s, e);
  s.tree_valid = true;
  s.tree_dirty = false;
  s.search_valid = false;
  if(e.finish) finish_tree(s);
}

void build_tree(System & s) {
  TreeEdges e;
  build_tree_edges(s, e);
}

// ---- work cells ---------------------------------------------------------------------------------------
// traversal_sph, tree_topology.h:221-235: maximal nodes with sub_entities <= 128, plus lone leaves hanging
// under bigger nodes.  Work cells are disjoint particle ranges; listing them by first particle gives Morton order.
__global__ void __launch_bounds__(256) k_wc_mark(Tree t, int64_t nc, uint32_t * __restrict__ wc_at, const int32_t * __restrict__ type,
  int ghosts, int limit) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= nc) return;
  int v = t.cell[c];
  int p = t.cell_parent[c];
  bool small = v >= 0 || t.sub[~v] <= limit;
  bool parent_big = p < 0 || t.sub[p] > limit;
  if(small && parent_big) {
    int start = v >= 0 ? v : t.node_start[~v];
    // cells of remote copies are only sources: the rank that owns them runs their sweeps (tree_topology.h:229, !is_shared)
    if(ghosts && (type[start] & TYPE_GHOST)) return;
    wc_at[start] = (uint32_t)c + 1u;
  }
}
__global__ void __launch_bounds__(256) k_flag_nonzero(const uint32_t * __restrict__ in, uint32_t * __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = in[i] != 0;
}
__global__ void __launch_bounds__(256) k_wc_compact(const uint32_t * __restrict__ wc_at, const uint32_t * __restrict__ ex, int64_t n,
  int32_t * __restrict__ wc_cell) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && wc_at[i]) wc_cell[ex[i]] = (int32_t)(wc_at[i] - 1u);
}

void mark_maximal_cells(System & s, int limit, bool skip_ghosts, uint32_t * at) {
  const int64_t nc = s.tree.ncells;
  FSPH_CUDA_OK(cudaMemsetAsync(at, 0, (size_t)s.n * 4, s.stream));
  FSPH_LAUNCH(k_wc_mark, (unsigned)((nc + 255) / 256), 256, s.stream, s.tree, nc, at, s.col_type.as<int32_t>(),
    (skip_ghosts && s.has_ghosts) ? 1 : 0, limit);
}

void find_work_cells(System & s) {
  const int64_t n = s.n;
  s.b_cnt0.reserve((size_t)(n + 1) * 4);
  s.b_tmp[0].reserve((size_t)(n + 1) * 4);
  uint32_t * wc_at = s.b_cnt0.as<uint32_t>();
  uint32_t * ex = s.b_tmp[0].as<uint32_t>();
  mark_maximal_cells(s, SUB_ENTITIES, true, wc_at);
  unsigned nbp = (unsigned)((n + 255) / 256);
  FSPH_LAUNCH(k_flag_nonzero, nbp, 256, s.stream, wc_at, ex, n);
  exclusive_scan_u32(s, ex, ex, n, nullptr);
  uint32_t nwc = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&nwc, ex + n, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  s.b_wc_cell.reserve((size_t)std::max<uint32_t>(nwc, 1) * 4);
  s.search.wc_cell = s.b_wc_cell.as<int32_t>();
  s.search.nwc = (int)nwc;
  FSPH_LAUNCH(k_wc_compact, nbp, 256, s.stream, wc_at, ex, n, s.search.wc_cell);
  FSPH_CUDA_OK(cudaGetLastError());
}

}  // namespace fsph
