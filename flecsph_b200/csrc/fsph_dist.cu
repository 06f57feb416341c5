// fsph_dist.cu -- Morton-key domain decomposition of one body_system over several GPUs (compiled with -fmad=false).
//
// Replaces the multi-rank half of body_system::update_iteration (mpisph/bodies_system.h:152-229):
//   mpi_compute_range's MPI_Allreduce                      bodies_system.h:253-262
//   psort::psort (sample sort: splitters, MPI_Alltoallv)   mpisph/psort.h:227-322
//   exchange_boundaries_ / share_nodes_ / cofm_update_     include/tree_topology/tree_topology.h:1549-1744,1752,1976
//   the lazy request / reply of remote branches            tree_topology.h:1276-1542 (used by traversal_sph :255-410)
//   reset_ghosts (re-fetch of remote bodies' fields)       bodies_system.h:273, tree_topology.h:153-157
//
// Design.  Every rank owns one contiguous range of the global (key) order.
//   1. range: local min/max, all-gather, combine (min / max are exact, so every rank holds the 1-GPU range and keys).
//   2. partition: local radix sort, regular key samples all-gathered, weighted splitters; each splitter is then moved
//      down to the first key of the work cell (maximal sub-tree of <= 128 bodies, tree_topology.h:221-235) that
//      straddles it, found from the 128 keys either side of it on every rank -- so no work cell is ever split
//      between ranks and the sweeps of every body run exactly as on one GPU.  One all-to-all-v per column.
//   3. local tree = the global canonical octree restricted to the bodies present: a body's leaf depth in the global
//      tree depends only on its two key neighbours, which are on this rank or are the edge keys of the adjacent
//      ranks (one all-gather of first / last keys); with that depth attached to each body the bottom-up builder
//      (fsph_tree.cu) reproduces every global ancestor, also for ghosts whose key neighbours stayed remote.
//   4. halo = locally essential tree.  Each rank all-gathers an occupancy bitmap of its bodies (2^24 cells, indexed
//      by the leading key digits) and its largest h.  A sender walks its own tree top-down once, carrying a bitmask of
//      destination ranks: a node stays open for rank r while its box, grown by hmax_r, touches an occupied cell of r
//      (a necessary condition of intersects_sphere_box for any body of r, tree_topology.h:348); leaf children of open
//      nodes are sent when their sphere of radius max(h, hmax_r) touches an occupied cell (necessary for the pair
//      predicate :377-381).  Bodies travel with all their live columns; every open node travels as (key, box).
//   5. the receiver merges owned bodies and ghosts (radix sort), rebuilds the tree with the given leaf depths and
//      overrides the boxes of nodes that hold remote bodies with the records (element-wise min / max: exact in any
//      order); nodes that straddle a rank boundary are completed from an all-gather of every rank's boundary-path
//      nodes.  Hence every box a local work cell can test is the global box: neighbour sets are those of one GPU.
//   6. reset_ghosts: the owners re-send the columns the next sweep reads (one packed all-to-all-v).
// Ghost bodies carry TYPE_GHOST in the type column; they are sources only (no work cell, no element-wise update).
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "fsph_internal.h"

namespace fsph {

constexpr int DD_SAMPLES = 2048;   // regular key samples per rank (imbalance <= nranks / DD_SAMPLES)
constexpr int DD_WIN = SUB_ENTITIES;
constexpr int DD_MAX_PATH = 192;   // boundary-path node records per rank (<= 2 * 63 straddling nodes)
constexpr int DD_MAX_RANKS = 32;

struct NodeBoxRec {
  uint64_t key;   // key prefix of the node with its sentinel bit (depth = position of the sentinel / D)
  double bmin[3], bmax[3];
};

struct DDGeom {
  double lo[3], scale[3], two_b;
  int B, gb, D, P, rank;
  double hmax[DD_MAX_RANKS];
  double rlo[DD_MAX_RANKS][3], rhi[DD_MAX_RANKS][3];  // per rank: min(x - h), max(x + h)
  size_t bitmap_words;
};

struct LevelTable {
  int begin[MAX_LEVELS + 1];
  int nlevels;
};

struct Dist {
  std::vector<uint64_t> split;  // [P + 1] first key of every rank's range (split[0] = 0, split[P] = ~0)
  std::vector<uint64_t> split_raw;  // the same before the alignment to work cells (kept from step to step)
  std::vector<size_t> send_cnt, send_off, recv_cnt, recv_off;  // halo plan, in bodies
  int64_t nsend = 0, nghost = 0;
  DevBuf send_idx, ghost_pos, sendbuf, recvbuf, small, small2, bitmap, bitmap_all, nmask, emask, flags, ex;
  DevBuf recs_send, recs_recv, path, path_all, counters, sel_idx, sel_mask;
  std::vector<double> rlo, rhi;  // box of every rank's bodies before the exchange (range all-gather)
  DDGeom geom;
  System * fmm_sys = nullptr;  // the gathered global body set + tree of gravitation_fmm on several ranks
};

void dist_release(System & s) {
  if(!s.dist) return;
  Dist & d = *s.dist;
  DevBuf * all[] = {&d.send_idx, &d.ghost_pos, &d.sendbuf, &d.recvbuf, &d.small, &d.small2, &d.bitmap, &d.bitmap_all, &d.nmask, &d.emask,
    &d.flags, &d.ex, &d.recs_send, &d.recs_recv, &d.path, &d.path_all, &d.counters, &d.sel_idx, &d.sel_mask};
  for(DevBuf * b : all) b->release();
  if(d.fmm_sys) {
    release_buffers(*d.fmm_sys);
    delete d.fmm_sys;
  }
  delete s.dist;
  s.dist = nullptr;
}

static Dist & dist_of(System & s) {
  if(!s.dist) s.dist = new Dist;
  return *s.dist;
}

void dist_all_gather_host(System & s, const void * in, void * out, size_t bytes) {
  Dist & d = dist_of(s);
  const int P = s.comm->nranks;
  d.small.reserve(bytes + 256);
  d.small2.reserve(bytes * P + 256);
  FSPH_CUDA_OK(cudaMemcpyAsync(d.small.p, in, bytes, cudaMemcpyHostToDevice, s.stream));
  s.comm->all_gather(d.small.p, d.small2.p, bytes, s.stream);
  FSPH_CUDA_OK(cudaMemcpyAsync(out, d.small2.p, bytes * P, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
}

// ---- host planning (pure functions; exported for the CPU tests) -----------------------------------------------

static int common_digits(uint64_t a, uint64_t b, int D) {
  const int B = 63 / D;
  const uint64_t x = a ^ b;
  if(!x) return B;
  return B - 1 - (63 - __builtin_clzll(x)) / D;
}

// Weighted splitters from regular samples: sample j of rank r stands for counts[r] / nsamples bodies.
void dd_choose_splitters(int nranks, int nsamples, const uint64_t * keys, const int64_t * counts, uint64_t * out) {
  std::vector<std::pair<uint64_t, double>> sm;
  double total = 0;
  for(int r = 0; r < nranks; ++r) {
    total += (double)counts[r];
    if(counts[r] <= 0) continue;
    for(int j = 0; j < nsamples; ++j) sm.emplace_back(keys[(size_t)r * nsamples + j], (double)counts[r] / nsamples);
  }
  std::sort(sm.begin(), sm.end());
  size_t i = 0;
  double cum = 0;
  for(int k = 1; k < nranks; ++k) {
    const double target = total * k / nranks;
    while(i < sm.size() && cum + sm[i].second <= target) cum += sm[i++].second;
    out[k - 1] = i < sm.size() ? sm[i].first : ~0ull;
  }
}

// Move splitter s down to the first key of the work cell that straddles it (if any).  below / above: every key any
// rank holds in the DD_WIN positions below / at-or-above s (unsorted, duplicates impossible).
uint64_t dd_align_splitter(int D, std::vector<uint64_t> & below, std::vector<uint64_t> & above, uint64_t s) {
  if(below.empty() || above.empty()) return s;
  std::sort(below.begin(), below.end());
  std::sort(above.begin(), above.end());
  if(below.size() > (size_t)DD_WIN) below.erase(below.begin(), below.end() - DD_WIN);
  if(above.size() > (size_t)DD_WIN) above.resize(DD_WIN);
  const uint64_t a0 = above.front(), b0 = below.back();
  const int B = 63 / D, L = common_digits(a0, b0, D);
  for(int d = 0; d <= L; ++d) {
    const int sh = D * (B - d);
    const uint64_t p = a0 >> sh;
    size_t cnt = 0;
    for(uint64_t k : below) cnt += (k >> sh) == p;
    for(uint64_t k : above) cnt += (k >> sh) == p;
    if(cnt <= (size_t)SUB_ENTITIES) return p << sh;
  }
  return s;
}

// ---- small kernels ----------------------------------------------------------------------------------------
__global__ void k_sample_keys(const uint64_t * __restrict__ key, int64_t n, int S, uint64_t * __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= S) return;
  int64_t i = (int64_t)(((double)j + 0.5) * (double)n / (double)S);
  if(i >= n) i = n - 1;
  out[j] = n > 0 ? key[i] : ~0ull;
}

__device__ __forceinline__ int64_t lower_bound_key(const uint64_t * __restrict__ key, int64_t n, uint64_t v) {
  int64_t lo = 0, hi = n;
  while(lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if(key[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// one block per splitter: the DD_WIN keys below and at-or-above it (0 / ~0 where this rank has none)
__global__ void __launch_bounds__(2 * DD_WIN) k_split_windows(const uint64_t * __restrict__ key, int64_t n, const uint64_t * __restrict__ split,
  uint64_t * __restrict__ out) {
  __shared__ int64_t pos_sm;
  if(threadIdx.x == 0) pos_sm = lower_bound_key(key, n, split[blockIdx.x]);
  __syncthreads();
  const int64_t pos = pos_sm;
  const int t = threadIdx.x;
  uint64_t v;
  if(t < DD_WIN) {
    const int64_t i = pos - DD_WIN + t;
    v = i >= 0 ? key[i] : 0ull;
  }
  else {
    const int64_t i = pos + (t - DD_WIN);
    v = i < n ? key[i] : ~0ull;
  }
  out[(size_t)blockIdx.x * 2 * DD_WIN + t] = v;
}

__global__ void k_cuts(const uint64_t * __restrict__ key, int64_t n, const uint64_t * __restrict__ split, int ns, int64_t * __restrict__ cut) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k < ns) cut[k] = lower_bound_key(key, n, split[k]);
}

__global__ void __launch_bounds__(256) k_owned_flag(const int32_t * __restrict__ type, int64_t n, uint32_t * __restrict__ f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) f[i] = (type[i] & TYPE_GHOST) ? 0u : 1u;
}
__global__ void __launch_bounds__(256) k_bit_flag(const uint32_t * __restrict__ mask, int64_t n, int bit, uint32_t * __restrict__ f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) f[i] = (mask[i] >> bit) & 1u;
}
__global__ void __launch_bounds__(256) k_compact_idx(const uint32_t * __restrict__ f, const uint32_t * __restrict__ ex, int64_t n,
  uint32_t * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && f[i]) out[ex[i]] = (uint32_t)i;
}
template<typename T>
__global__ void __launch_bounds__(256) k_pack(const T * __restrict__ in, const uint32_t * __restrict__ idx, int64_t n, T * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = in[idx[i]];
}
struct FieldPtrs {
  uint64_t * p[16];
};
struct ColPtrs8 { uint64_t * p[NF64 + 4]; };
struct ColPtrs4 { uint32_t * p[4]; };
// halo traffic of one peer, all live columns at once: out[col][body] = col[idx[body]]   (blockIdx.y = column)
template<typename T, typename PT>
__global__ void __launch_bounds__(256) k_pack_cols(PT f, const uint32_t * __restrict__ idx, int64_t n, T * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[(size_t)blockIdx.y * n + i] = reinterpret_cast<const T *>(f.p[blockIdx.y])[idx[i]];
}
template<typename T, typename PT>
__global__ void __launch_bounds__(256) k_unpack_cols(PT f, int64_t dst, int64_t n, const T * __restrict__ in) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) reinterpret_cast<T *>(f.p[blockIdx.y])[dst + i] = in[(size_t)blockIdx.y * n + i];
}
__global__ void __launch_bounds__(256) k_nonzero_flag(const uint32_t * __restrict__ mask, int64_t n, uint32_t * __restrict__ f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) f[i] = mask[i] != 0u;
}
// selected[k] = (index, mask) of the k-th element with a non-zero destination mask
__global__ void __launch_bounds__(256) k_select(const uint32_t * __restrict__ mask, const uint32_t * __restrict__ f, const uint32_t * __restrict__ ex,
  int64_t n, uint32_t * __restrict__ sel_idx, uint32_t * __restrict__ sel_mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && f[i]) { sel_idx[ex[i]] = (uint32_t)i; sel_mask[ex[i]] = mask[i]; }
}
__global__ void __launch_bounds__(256) k_compact_sel(const uint32_t * __restrict__ f, const uint32_t * __restrict__ ex, int64_t n,
  const uint32_t * __restrict__ sel_idx, uint32_t * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && f[i]) out[ex[i]] = sel_idx[i];
}
// refresh traffic of one peer: out[field][body] = col[field][idx[body]]  (blockIdx.y = field)
__global__ void __launch_bounds__(256) k_pack_fields(FieldPtrs f, const uint32_t * __restrict__ idx, int64_t n, uint64_t * __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[(size_t)blockIdx.y * n + i] = f.p[blockIdx.y][idx[i]];
}
__global__ void __launch_bounds__(256) k_unpack_fields(FieldPtrs f, const uint32_t * __restrict__ pos, int64_t n, const uint64_t * __restrict__ in) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) f.p[blockIdx.y][pos[i]] = in[(size_t)blockIdx.y * n + i];
}
__global__ void __launch_bounds__(256) k_mark_ghost(int32_t * __restrict__ type, int64_t b, int64_t e) {
  int64_t i = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < e) type[i] |= TYPE_GHOST;
}
// merged order [ghosts of lower ranks | owned | ghosts of higher ranks] from storage order [owned | ghosts by sender rank]
__global__ void __launch_bounds__(256) k_merge_perm(uint32_t * __restrict__ perm, int64_t n, int64_t n_glow, int64_t n_owned) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  perm[i] = (uint32_t)(i < n_glow ? n_owned + i : (i < n_glow + n_owned ? i - n_glow : i));
}
__global__ void __launch_bounds__(256) k_ghost_pos(uint32_t * __restrict__ pos, int64_t ng, int64_t n_glow, int64_t n_owned) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(j < ng) pos[j] = (uint32_t)(j < n_glow ? j : n_owned + j);
}
__global__ void __launch_bounds__(256) k_shift_idx(uint32_t * __restrict__ idx, int64_t n, uint32_t by) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) idx[i] += by;
}
__global__ void __launch_bounds__(256) k_count_bits(const uint32_t * __restrict__ mask, int64_t n, int P, unsigned long long * __restrict__ cnt) {
  __shared__ unsigned int sm[DD_MAX_RANKS];
  if(threadIdx.x < DD_MAX_RANKS) sm[threadIdx.x] = 0;
  __syncthreads();
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t m = mask[i];
    while(m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      atomicAdd(&sm[b], 1u);
    }
  }
  __syncthreads();
  if(threadIdx.x < P && sm[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)sm[threadIdx.x]);
}

// ---- occupancy bitmap --------------------------------------------------------------------------------------
template<int D>
__device__ __forceinline__ size_t cell_index(const uint32_t c[3], int gb) {
  if(D == 1) return c[0];
  if(D == 2) return ((size_t)c[1] << gb) | c[0];
  return ((((size_t)c[2] << gb) | c[1]) << gb) | c[0];
}

template<int D>
__global__ void __launch_bounds__(256) k_occupancy(const uint64_t * __restrict__ key, int64_t n, int gb, uint32_t * __restrict__ bm) {
  constexpr int B = 63 / D;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint64_t k = key[i];
  uint32_t c[3] = {0, 0, 0};
  for(int b = 0; b < gb; ++b) {
#pragma unroll
    for(int d = 0; d < D; ++d) c[d] |= (uint32_t)((k >> ((B - 1 - b) * D + d)) & 1ull) << (gb - 1 - b);
  }
  const size_t idx = cell_index<D>(c, gb);
  const uint32_t bit = 1u << (idx & 31);
  uint32_t * w = bm + (idx >> 5);
  if(!(*w & bit)) atomicOr(w, bit);
}

// cell coordinate of x along axis d: the leading gb bits of the key coordinate (filling_curve.h:513-515), monotone in x
__device__ __forceinline__ uint32_t cell_coord(const DDGeom & g, int d, double x) {
  const double t = (x - g.lo[d]) / g.scale[d] * g.two_b;
  const uint64_t maxv = (uint64_t(1) << g.B) - 1;
  uint64_t v = t > 0.0 ? (t < g.two_b ? (uint64_t)t : maxv) : 0ull;
  if(v > maxv) v = maxv;
  return (uint32_t)(v >> (g.B - g.gb));
}

// does the axis-aligned box [lo, hi] touch an occupied cell of rank q?  (big boxes: "yes" without looking)
template<int D>
__device__ __forceinline__ bool box_touches(const DDGeom & g, const uint32_t * __restrict__ bm_all, int q, const double * lo, const double * hi) {
#pragma unroll
  for(int d = 0; d < D; ++d)
    if(hi[d] < g.rlo[q][d] || lo[d] > g.rhi[q][d]) return false;
  uint32_t c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
#pragma unroll
  for(int d = 0; d < D; ++d) { c0[d] = cell_coord(g, d, lo[d]); c1[d] = cell_coord(g, d, hi[d]); }
  const uint32_t w0 = c0[0] >> 5, w1 = c1[0] >> 5;
  const size_t rows = (size_t)(c1[1] - c0[1] + 1) * (size_t)(c1[2] - c0[2] + 1);
  if(rows * (w1 - w0 + 1) > 192) return true;
  const uint32_t * bm = bm_all + (size_t)q * g.bitmap_words;
  for(uint32_t z = c0[2]; z <= c1[2]; ++z)
    for(uint32_t y = c0[1]; y <= c1[1]; ++y) {
      const uint32_t cc[3] = {0, y, z};
      const size_t row = cell_index<D>(cc, g.gb) >> 5;
      for(uint32_t w = w0; w <= w1; ++w) {
        uint32_t m = bm[row + w];
        if(w == w0) m &= 0xffffffffu << (c0[0] & 31);
        if(w == w1) m &= 0xffffffffu >> (31 - (c1[0] & 31));
        if(m) return true;
      }
    }
  return false;
}

// ---- locally essential tree walk (sender side) -------------------------------------------------------------
template<int D>
__global__ void __launch_bounds__(128) k_let_nodes(Tree t, int nb, int ne, DDGeom g, const uint32_t * __restrict__ bm_all,
  uint32_t * __restrict__ nmask) {
  const int nd = nb + blockIdx.x * blockDim.x + threadIdx.x;
  if(nd >= ne) return;
  const int par = t.node_parent[nd];
  uint32_t pm = par < 0 ? (((g.P >= 32 ? 0u : (1u << g.P)) - 1u) & ~(1u << g.rank)) : nmask[par];
  uint32_t m = 0;
  double lo[D], hi[D];
  while(pm) {
    const int q = __ffs(pm) - 1;
    pm &= pm - 1;
#pragma unroll
    for(int d = 0; d < D; ++d) { lo[d] = t.bmin[d][nd] - g.hmax[q]; hi[d] = t.bmax[d][nd] + g.hmax[q]; }
    if(box_touches<D>(g, bm_all, q, lo, hi)) m |= 1u << q;
  }
  nmask[nd] = m;
}

template<int D>
__global__ void __launch_bounds__(128) k_let_leaves(Tree t, Cols c, int64_t nc, DDGeom g, const uint32_t * __restrict__ bm_all,
  const uint32_t * __restrict__ nmask, uint32_t * __restrict__ emask) {
  const int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(cc >= nc) return;
  const int v = t.cell[cc];
  if(v < 0) return;
  uint32_t pm = nmask[t.cell_parent[cc]];
  uint32_t m = 0;
  if(pm) {
    double x[D], lo[D], hi[D];
#pragma unroll
    for(int d = 0; d < D; ++d) x[d] = c.f[F_POS + d][v];
    const double h = c.f[F_H][v];
    while(pm) {
      const int q = __ffs(pm) - 1;
      pm &= pm - 1;
      const double r = fmax(h, g.hmax[q]);
#pragma unroll
      for(int d = 0; d < D; ++d) { lo[d] = x[d] - r; hi[d] = x[d] + r; }
      if(box_touches<D>(g, bm_all, q, lo, hi)) m |= 1u << q;
    }
  }
  emask[v] = m;
}

// ---- node box records ------------------------------------------------------------------------------------
template<int D>
__device__ __forceinline__ uint64_t node_key(const Tree & t, const uint64_t * __restrict__ key, int nd) {
  constexpr int B = 63 / D;
  return key[t.node_start[nd]] >> (D * (B - t.node_depth[nd]));
}
template<int D>
__device__ __forceinline__ void node_key_range(uint64_t nk, uint64_t & first, uint64_t & last) {
  constexpr int B = 63 / D;
  const int depth = (63 - __clzll((long long)nk)) / D;
  const int sh = D * (B - depth);
  first = nk << sh;
  last = first | ((uint64_t(1) << sh) - 1ull);
}
template<int D>
__device__ __forceinline__ NodeBoxRec make_rec(const Tree & t, const uint64_t * __restrict__ key, int nd) {
  NodeBoxRec r;
  r.key = node_key<D>(t, key, nd);
#pragma unroll
  for(int d = 0; d < 3; ++d) {
    r.bmin[d] = d < D ? t.bmin[d][nd] : 0.0;
    r.bmax[d] = d < D ? t.bmax[d][nd] : 0.0;
  }
  return r;
}

// nodes of the owned tree that reach outside [klo, khi_incl]: slot 0 of `out` holds the count in .key
template<int D>
__global__ void __launch_bounds__(256) k_path_records(Tree t, const uint64_t * __restrict__ key, int nn, uint64_t klo, uint64_t khi_incl,
  NodeBoxRec * __restrict__ out, unsigned long long * __restrict__ cnt, uint32_t * status) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if(nd >= nn) return;
  uint64_t a, b;
  node_key_range<D>(node_key<D>(t, key, nd), a, b);
  if(a >= klo && b <= khi_incl) return;
  const unsigned long long i = atomicAdd(cnt, 1ull);
  if(i >= (unsigned long long)DD_MAX_PATH) { atomicOr(status, (uint32_t)ERR_DD_CAPACITY); return; }
  out[1 + i] = make_rec<D>(t, key, nd);
}
__global__ void k_set_count(NodeBoxRec * out, const unsigned long long * cnt) {
  out[0].key = *cnt < (unsigned long long)DD_MAX_PATH ? *cnt : (unsigned long long)DD_MAX_PATH;
}

// records of the selected nodes sel_idx[k] (k < n) whose flag is set
template<int D>
__global__ void __launch_bounds__(256) k_emit_records(Tree t, const uint64_t * __restrict__ key, int64_t n, const uint32_t * __restrict__ sel_idx,
  const uint32_t * __restrict__ f, const uint32_t * __restrict__ ex, NodeBoxRec * __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k < n && f[k]) out[ex[k]] = make_rec<D>(t, key, (int)sel_idx[k]);
}

__device__ __forceinline__ void atomic_min_f64(double * p, double v) {
  unsigned long long old = __double_as_longlong(*p);
  while(__longlong_as_double(old) > v) {
    const unsigned long long seen = atomicCAS(reinterpret_cast<unsigned long long *>(p), old, (unsigned long long)__double_as_longlong(v));
    if(seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_f64(double * p, double v) {
  unsigned long long old = __double_as_longlong(*p);
  while(__longlong_as_double(old) < v) {
    const unsigned long long seen = atomicCAS(reinterpret_cast<unsigned long long *>(p), old, (unsigned long long)__double_as_longlong(v));
    if(seen == old) break;
    old = seen;
  }
}

// records -> boxes of the nodes with the same key (absent nodes are skipped).  `stride_blocks` > 0: the input is
// nblocks blocks of (1 + DD_MAX_PATH) records with the count in slot 0 (all-gathered boundary paths).
template<int D>
__global__ void __launch_bounds__(256) k_apply_records(Tree t, const uint64_t * __restrict__ key, LevelTable lv, const NodeBoxRec * __restrict__ recs,
  int64_t nrec, int path_blocks) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= nrec) return;
  if(path_blocks) {
    const int64_t blk = i / (1 + DD_MAX_PATH), j = i % (1 + DD_MAX_PATH);
    if(j == 0 || j > (int64_t)recs[blk * (1 + DD_MAX_PATH)].key) return;
  }
  const NodeBoxRec r = recs[i];
  const int depth = (63 - __clzll((long long)r.key)) / D;
  if(depth >= lv.nlevels) return;
  int lo = lv.begin[depth], hi = lv.begin[depth + 1];
  while(lo < hi) {
    const int mid = (lo + hi) >> 1;
    if(node_key<D>(t, key, mid) < r.key) lo = mid + 1;
    else hi = mid;
  }
  if(lo >= lv.begin[depth + 1] || node_key<D>(t, key, lo) != r.key) return;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    atomic_min_f64(&t.bmin[d][lo], r.bmin[d]);
    atomic_max_f64(&t.bmax[d][lo], r.bmax[d]);
  }
}

// Nodes that straddle a rank boundary hold more than 128 bodies globally (the splitters were aligned to work
// cells), whatever part of them is present here: they are never work cells (tree_topology.h:225 descends shared nodes).
template<int D>
__global__ void __launch_bounds__(256) k_fix_straddling(Tree t, const uint64_t * __restrict__ key, int nn, uint64_t klo, uint64_t khi_incl) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if(nd >= nn) return;
  uint64_t a, b;
  node_key_range<D>(node_key<D>(t, key, nd), a, b);
  const bool owned = a >= klo && b <= khi_incl, remote = b < klo || a > khi_incl;
  if(!owned && !remote) t.sub[nd] = 1 << 30;  // neither a work cell nor a super cell of the two-level walk
}

// ---- host side -----------------------------------------------------------------------------------------------
static unsigned blocks(int64_t n, int nt = 256) { return (unsigned)std::max<int64_t>((n + nt - 1) / nt, 1); }

void dist_drop_ghosts(System & s) {
  if(!s.has_ghosts) return;
  Dist & d = dist_of(s);
  const int64_t n = s.n;
  d.flags.reserve((size_t)(n + 1) * 4);
  d.ex.reserve((size_t)(n + 1) * 4);
  s.b_perm.reserve((size_t)std::max<int64_t>(n, 1) * 4);
  uint32_t * f = d.flags.as<uint32_t>();
  uint32_t * ex = d.ex.as<uint32_t>();
  FSPH_LAUNCH(k_owned_flag, blocks(n), 256, s.stream, s.col_type.as<int32_t>(), n, f);
  exclusive_scan_u32(s, f, ex, n, nullptr);
  FSPH_LAUNCH(k_compact_idx, blocks(n), 256, s.stream, f, ex, n, s.b_perm.as<uint32_t>());
  gather_bodies(s, s.b_perm.as<uint32_t>(), s.n_owned);
  s.n = s.n_owned;
  s.has_ghosts = false;
  s.tree_valid = false;
  s.search_valid = false;
  s.last_perm = nullptr;
}

template<int D>
static void apply_records(System & s, const NodeBoxRec * recs, int64_t nrec, int path_blocks) {
  if(nrec <= 0) return;
  LevelTable lv;
  for(int k = 0; k <= MAX_LEVELS; ++k) lv.begin[k] = k <= s.nlevels ? s.level_node_begin[k] : s.level_node_begin[s.nlevels];
  lv.nlevels = s.nlevels;
  FSPH_LAUNCH((k_apply_records<D>), blocks(nrec), 256, s.stream, s.tree, s.col_key.as<uint64_t>(), lv, recs, nrec, path_blocks);
}

// one column-wise all-to-all-v of bodies: block q of every live column = bodies [soff[q], soff[q] + scnt[q])
static void exchange_bodies(System & s, const std::vector<size_t> & soff, const std::vector<size_t> & scnt, const std::vector<size_t> & rcnt) {
  const int P = s.comm->nranks;
  std::vector<size_t> roff(P);
  size_t n_new = 0;
  for(int q = 0; q < P; ++q) { roff[q] = n_new; n_new += rcnt[q]; }
  std::vector<std::pair<DevBuf *, int>> cols;
  live_columns(s, cols);
  std::vector<size_t> so(P), sc(P), ro(P), rc(P);
  // head-room for the halo that is appended later; buffers only ever grow (cudaMalloc / cudaFree cost ~1 ms each)
  const size_t room = n_new + n_new / 3 + 4096;
  for(auto & c : cols) {
    const size_t e = (size_t)c.second;
    DevBuf & scratch = e == 8 ? s.scratch_col : s.b_tmp[0];
    if(scratch.cap < std::max(n_new * e, c.first->cap)) scratch.reserve_exact(std::max(room * e, c.first->cap));
    for(int q = 0; q < P; ++q) { so[q] = soff[q] * e; sc[q] = scnt[q] * e; ro[q] = roff[q] * e; rc[q] = rcnt[q] * e; }
    s.comm->all_to_all_v(c.first->p, so.data(), sc.data(), scratch.p, ro.data(), rc.data(), s.stream);
    std::swap(c.first->p, scratch.p);
    std::swap(c.first->cap, scratch.cap);
  }
  s.n = (int64_t)n_new;
}

static void ensure_capacity(System & s, int64_t n_total) {
  // columns that are not carried (dead inside fsph_step, derived, neighbour counts) only need the room
  for(int f = 0; f < NF64; ++f)
    if(s.col_used[f]) s.col[f].grow_keep((size_t)n_total * 8, 0, s.stream);
  s.col_nneigh.grow_keep((size_t)n_total * 4, 0, s.stream);
  s.scratch_col.reserve((size_t)n_total * 8);
}

template<int D>
static void dist_update_iteration_t(System & s) {
  Comm & cm = *s.comm;
  const int P = cm.nranks, R = cm.rank;
  if(P > DD_MAX_RANKS) throw std::runtime_error("more than 32 ranks are not supported");
  Dist & d = dist_of(s);
  constexpr int B = 63 / D;
  std::unique_ptr<PhaseScope> ph;
  ph.reset(new PhaseScope(s, "dd:range_keys_sort"));
  // the remote copies of the previous iteration are still interleaved with the owned bodies: they take no part in the
  // range, get the largest key and fall off the end of the next sort (no separate compaction pass)
  const int64_t n_keep = s.n_owned_or_n();
  if(n_keep <= 0) throw std::runtime_error("a rank without bodies cannot take part in the decomposition");
  s.col_gd.grow_keep((size_t)s.n * 4, 0, s.stream);

  // 1. global range (bodies_system.h:231-268): min / max are exact, every rank gets the single-rank result
  compute_range(s);
  d.rlo.assign((size_t)3 * P, 0.0);
  d.rhi.assign((size_t)3 * P, 0.0);
  {
    double mine[6] = {s.lo[0], s.lo[1], s.lo[2], s.hi[0], s.hi[1], s.hi[2]};
    std::vector<double> all((size_t)6 * P);
    dist_all_gather_host(s, mine, all.data(), sizeof(mine));
    for(int q = 0; q < P; ++q)
      for(int k = 0; k < 3; ++k) {
        d.rlo[(size_t)3 * q + k] = all[(size_t)6 * q + k];
        d.rhi[(size_t)3 * q + k] = all[(size_t)6 * q + 3 + k];
      }
    for(int k = 0; k < D; ++k) {
      double lo = DBL_MAX, hi = -DBL_MAX;
      for(int q = 0; q < P; ++q) { lo = std::min(lo, d.rlo[(size_t)3 * q + k]); hi = std::max(hi, d.rhi[(size_t)3 * q + k]); }
      s.lo[k] = lo;
      s.hi[k] = hi;
    }
  }
  compute_keys(s);
  sort_bodies(s, n_keep);
  s.has_ghosts = false;
  s.n_owned = s.n;
  s.tree_valid = false;
  s.search_valid = false;

  ph.reset(new PhaseScope(s, "dd:splitters"));
  // 2. splitters (psort.h:258-301).  The raw splitters of the previous iteration are kept while they balance the ranks
  //    to within 3 %: bodies move little per step, so in steady state (almost) nothing changes rank.  They are
  //    re-aligned to the work cells of the current bodies every time.
  bool moved_any = false;
  std::vector<size_t> soff(P), scnt(P), rcnt(P);
  auto sample_splitters = [&]() {
    d.split_raw.assign(P + 1, 0);
    d.split_raw[P] = ~0ull;
    d.small.reserve((size_t)(DD_SAMPLES + 1) * 8 + 256);
    d.small2.reserve((size_t)(DD_SAMPLES + 1) * 8 * P + 256);
    uint64_t * smp = d.small.as<uint64_t>();
    FSPH_LAUNCH(k_sample_keys, blocks(DD_SAMPLES), 256, s.stream, s.col_key.as<uint64_t>(), s.n, DD_SAMPLES, smp);
    const uint64_t cnt = (uint64_t)s.n;
    FSPH_CUDA_OK(cudaMemcpyAsync(smp + DD_SAMPLES, &cnt, 8, cudaMemcpyHostToDevice, s.stream));
    cm.all_gather(smp, d.small2.p, (size_t)(DD_SAMPLES + 1) * 8, s.stream);
    std::vector<uint64_t> all((size_t)(DD_SAMPLES + 1) * P);
    FSPH_CUDA_OK(cudaMemcpyAsync(all.data(), d.small2.p, all.size() * 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    std::vector<uint64_t> keys((size_t)DD_SAMPLES * P);
    std::vector<int64_t> counts(P);
    for(int q = 0; q < P; ++q) {
      std::copy(all.begin() + (size_t)q * (DD_SAMPLES + 1), all.begin() + (size_t)q * (DD_SAMPLES + 1) + DD_SAMPLES, keys.begin() + (size_t)q * DD_SAMPLES);
      counts[q] = (int64_t)all[(size_t)q * (DD_SAMPLES + 1) + DD_SAMPLES];
    }
    dd_choose_splitters(P, DD_SAMPLES, keys.data(), counts.data(), d.split_raw.data() + 1);
  };
  // align the raw splitters to work cells, cut the local order, all-gather the counts; returns "balanced"
  auto plan = [&]() -> bool {
    d.split = d.split_raw;
    std::vector<int64_t> cut(P + 1, 0);
    cut[P] = s.n;
    if(P > 1) {
      const int ns = P - 1;
      const size_t wbytes = (size_t)ns * 2 * DD_WIN * 8;
      d.small.reserve(wbytes + (size_t)ns * 16 + 256);
      d.small2.reserve(wbytes * P + 256);
      uint64_t * dsplit = reinterpret_cast<uint64_t *>(d.small.as<char>() + wbytes);
      int64_t * dcut = reinterpret_cast<int64_t *>(dsplit + ns);
      FSPH_CUDA_OK(cudaMemcpyAsync(dsplit, d.split.data() + 1, (size_t)ns * 8, cudaMemcpyHostToDevice, s.stream));
      FSPH_LAUNCH(k_split_windows, ns, 2 * DD_WIN, s.stream, s.col_key.as<uint64_t>(), s.n, dsplit, d.small.as<uint64_t>());
      cm.all_gather(d.small.p, d.small2.p, wbytes, s.stream);
      std::vector<uint64_t> all((size_t)ns * 2 * DD_WIN * P);
      FSPH_CUDA_OK(cudaMemcpyAsync(all.data(), d.small2.p, all.size() * 8, cudaMemcpyDeviceToHost, s.stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
      std::vector<uint64_t> below, above;
      for(int k = 0; k < ns; ++k) {
        below.clear();
        above.clear();
        for(int q = 0; q < P; ++q) {
          const uint64_t * w = all.data() + ((size_t)q * ns + k) * 2 * DD_WIN;
          for(int t = 0; t < DD_WIN; ++t) {
            if(w[t] != 0ull) below.push_back(w[t]);
            if(w[DD_WIN + t] != ~0ull) above.push_back(w[DD_WIN + t]);
          }
        }
        d.split[k + 1] = dd_align_splitter(D, below, above, d.split[k + 1]);
      }
      for(int k = 1; k < P; ++k) d.split[k] = std::max(d.split[k], d.split[k - 1]);
      // cut the local order at the splitters (psort.h:303-308)
      FSPH_CUDA_OK(cudaMemcpyAsync(dsplit, d.split.data() + 1, (size_t)ns * 8, cudaMemcpyHostToDevice, s.stream));
      FSPH_LAUNCH(k_cuts, 1, 32, s.stream, s.col_key.as<uint64_t>(), s.n, dsplit, ns, dcut);
      FSPH_CUDA_OK(cudaMemcpyAsync(cut.data() + 1, dcut, (size_t)ns * 8, cudaMemcpyDeviceToHost, s.stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    }
    std::vector<int64_t> mine(P), all((size_t)P * P);
    for(int q = 0; q < P; ++q) { soff[q] = (size_t)cut[q]; scnt[q] = (size_t)(cut[q + 1] - cut[q]); mine[q] = (int64_t)scnt[q]; }
    dist_all_gather_host(s, mine.data(), all.data(), (size_t)P * 8);
    moved_any = false;
    int64_t total = 0, worst = 0;
    for(int r = 0; r < P; ++r) {
      int64_t got = 0;
      for(int q = 0; q < P; ++q) {
        got += all[(size_t)q * P + r];
        if(q != r && all[(size_t)q * P + r] != 0) moved_any = true;
      }
      total += got;
      worst = std::max(worst, got);
    }
    for(int q = 0; q < P; ++q) rcnt[q] = (size_t)all[(size_t)q * P + R];
    return (double)worst <= 1.03 * (double)total / P + 256.0;
  };
  const bool had_splitters = (int)d.split_raw.size() == P + 1;
  if(!had_splitters) sample_splitters();
  if(!plan() && had_splitters) {
    sample_splitters();
    plan();
  }

  ph.reset(new PhaseScope(s, "dd:exchange"));
  // 3. the bodies change rank (psort.h:303-308); in steady state often none does: nothing to move, nothing to merge
  if(moved_any) exchange_bodies(s, soff, scnt, rcnt);
  if(s.n <= 0) throw std::runtime_error("too few bodies for this many ranks (a rank received none)");
  s.n_owned = s.n;
  ph.reset(new PhaseScope(s, "dd:sort_owned"));
  if(moved_any) sort_bodies(s);  // merge of the received runs

  ph.reset(new PhaseScope(s, "dd:tree_owned"));
  // 4. owned tree with the edge keys of the neighbouring ranks (exchange_boundaries_, tree_topology.h:1976)
  const uint64_t klo = d.split[R], khi_incl = R + 1 < P ? d.split[R + 1] - 1ull : ~0ull;
  TreeEdges te;
  {
    uint64_t edge[2];
    FSPH_CUDA_OK(cudaMemcpyAsync(&edge[0], s.col_key.as<uint64_t>(), 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&edge[1], s.col_key.as<uint64_t>() + (s.n - 1), 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    std::vector<uint64_t> all((size_t)2 * P);
    dist_all_gather_host(s, edge, all.data(), 16);
    if(R > 0) { te.has_before = true; te.key_before = all[(size_t)2 * (R - 1) + 1]; }
    if(R + 1 < P) { te.has_after = true; te.key_after = all[(size_t)2 * (R + 1)]; }
  }
  te.gd_out = s.col_gd.as<int32_t>();
  te.finish = false;
  build_tree_edges(s, te);

  // boundary-path nodes of every rank -> exact boxes of the nodes that straddle rank boundaries
  uint32_t * status = s.b_status.as<uint32_t>();
  d.counters.reserve(2 * DD_MAX_RANKS * 8 + 64);
  unsigned long long * ctr = d.counters.as<unsigned long long>();
  const size_t path_bytes = (size_t)(1 + DD_MAX_PATH) * sizeof(NodeBoxRec);
  d.path.reserve(path_bytes);
  d.path_all.reserve(path_bytes * P);
  FSPH_CUDA_OK(cudaMemsetAsync(ctr, 0, 2 * DD_MAX_RANKS * 8 + 64, s.stream));
  FSPH_LAUNCH((k_path_records<D>), blocks(s.tree.nnodes), 256, s.stream, s.tree, s.col_key.as<uint64_t>(), s.tree.nnodes, klo, khi_incl,
    d.path.as<NodeBoxRec>(), ctr + 2 * DD_MAX_RANKS, status);
  FSPH_LAUNCH(k_set_count, 1, 1, s.stream, d.path.as<NodeBoxRec>(), ctr + 2 * DD_MAX_RANKS);
  cm.all_gather(d.path.p, d.path_all.p, path_bytes, s.stream);
  apply_records<D>(s, d.path_all.as<NodeBoxRec>(), (int64_t)(1 + DD_MAX_PATH) * P, P);

  ph.reset(new PhaseScope(s, "dd:let_walk"));
  // 5. occupancy bitmaps + largest h of every rank
  DDGeom & g = d.geom;
  g.B = B; g.D = D; g.P = P; g.rank = R;
  g.gb = D == 3 ? 8 : (D == 2 ? 12 : 20);
  g.two_b = (double)(uint64_t(1) << B);
  g.bitmap_words = (size_t(1) << (D * g.gb)) / 32;
  for(int k = 0; k < 3; ++k) { g.lo[k] = k < D ? s.lo[k] : 0.0; g.scale[k] = k < D ? s.hi[k] - s.lo[k] : 1.0; }
  {
    // box (min(x - h), max(x + h)) and largest h of the bodies every rank owns NOW: coarse pre-filter of the walk
    const double glo[3] = {s.lo[0], s.lo[1], s.lo[2]}, ghi[3] = {s.hi[0], s.hi[1], s.hi[2]};
    compute_range(s);
    double mine[7] = {s.lo[0], s.lo[1], s.lo[2], s.hi[0], s.hi[1], s.hi[2], max_h(s)};
    for(int k = 0; k < 3; ++k) { s.lo[k] = glo[k]; s.hi[k] = ghi[k]; }
    std::vector<double> all((size_t)7 * P);
    dist_all_gather_host(s, mine, all.data(), sizeof(mine));
    for(int q = 0; q < P; ++q) {
      g.hmax[q] = all[(size_t)7 * q + 6];
      for(int k = 0; k < 3; ++k) { g.rlo[q][k] = all[(size_t)7 * q + k]; g.rhi[q][k] = all[(size_t)7 * q + 3 + k]; }
    }
  }
  ph.reset(new PhaseScope(s, "dd:let_bitmap"));
  d.bitmap.reserve(g.bitmap_words * 4);
  d.bitmap_all.reserve(g.bitmap_words * 4 * P);
  FSPH_CUDA_OK(cudaMemsetAsync(d.bitmap.p, 0, g.bitmap_words * 4, s.stream));
  FSPH_LAUNCH((k_occupancy<D>), blocks(s.n), 256, s.stream, s.col_key.as<uint64_t>(), s.n, g.gb, d.bitmap.as<uint32_t>());
  cm.all_gather(d.bitmap.p, d.bitmap_all.p, g.bitmap_words * 4, s.stream);

  ph.reset(new PhaseScope(s, "dd:let_nodes"));
  // 6. sender-side walk: destination masks of nodes (top-down, level by level) and of bodies
  const int nn = s.tree.nnodes;
  const int64_t nc = s.tree.ncells, n = s.n;
  d.nmask.reserve((size_t)std::max(nn, 1) * 4);
  d.emask.reserve((size_t)n * 4);
  uint32_t * nmask = d.nmask.as<uint32_t>();
  uint32_t * emask = d.emask.as<uint32_t>();
  for(int k = 0; k < s.nlevels; ++k) {
    const int b = s.level_node_begin[k], e = s.level_node_begin[k + 1];
    if(e > b) FSPH_LAUNCH((k_let_nodes<D>), blocks(e - b, 128), 128, s.stream, s.tree, b, e, g, d.bitmap_all.as<uint32_t>(), nmask);
  }
  ph.reset(new PhaseScope(s, "dd:let_leaves"));
  FSPH_LAUNCH((k_let_leaves<D>), blocks(nc, 128), 128, s.stream, s.tree, s.cols(), nc, g, d.bitmap_all.as<uint32_t>(), nmask, emask);
  ph.reset(new PhaseScope(s, "dd:let_select"));
  FSPH_LAUNCH(k_count_bits, (unsigned)std::min<int64_t>(blocks(n), 148 * 8), 256, s.stream, emask, n, P, ctr);
  FSPH_LAUNCH(k_count_bits, (unsigned)std::min<int64_t>(blocks(nn), 148 * 8), 256, s.stream, nmask, (int64_t)nn, P, ctr + DD_MAX_RANKS);
  const bool dd_debug = getenv("FSPH_DEBUG_DD") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tdbg[8] = {now(), 0, 0, 0, 0, 0, 0, 0};
  std::vector<unsigned long long> hc(2 * DD_MAX_RANKS);
  uint32_t st = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(hc.data(), ctr, hc.size() * 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(st & ERR_DD_CAPACITY) throw std::runtime_error("more boundary-path nodes than DD_MAX_PATH");

  tdbg[1] = now();
  d.send_cnt.assign(P, 0); d.send_off.assign(P, 0); d.recv_cnt.assign(P, 0); d.recv_off.assign(P, 0);
  std::vector<size_t> rs_cnt(P, 0), rs_off(P, 0), rr_cnt(P, 0), rr_off(P, 0);
  size_t nsend = 0, nrs = 0;
  for(int q = 0; q < P; ++q) {
    d.send_cnt[q] = (size_t)hc[q]; d.send_off[q] = nsend; nsend += d.send_cnt[q];
    rs_cnt[q] = (size_t)hc[DD_MAX_RANKS + q]; rs_off[q] = nrs; nrs += rs_cnt[q];
  }
  d.nsend = (int64_t)nsend;
  d.send_idx.reserve_roomy(std::max<size_t>(nsend, 1) * 4);
  d.recs_send.reserve_roomy(std::max<size_t>(nrs, 1) * sizeof(NodeBoxRec));
  {
    // the bodies / nodes with any destination are few: select them once, then split the short lists per destination
    const int64_t m = std::max<int64_t>(n, nn);
    d.flags.reserve((size_t)(m + 1) * 4);
    d.ex.reserve((size_t)(m + 1) * 4);
    uint32_t * f = d.flags.as<uint32_t>();
    uint32_t * ex = d.ex.as<uint32_t>();
    auto select = [&](const uint32_t * mask, int64_t cnt, DevBuf & bidx, DevBuf & bmask) -> int64_t {
      FSPH_LAUNCH(k_nonzero_flag, blocks(cnt), 256, s.stream, mask, cnt, f);
      exclusive_scan_u32(s, f, ex, cnt, nullptr);
      uint32_t nsel = 0;
      FSPH_CUDA_OK(cudaMemcpyAsync(&nsel, ex + cnt, 4, cudaMemcpyDeviceToHost, s.stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
      bidx.reserve_roomy(std::max<size_t>(nsel, 1) * 4);
      bmask.reserve_roomy(std::max<size_t>(nsel, 1) * 4);
      if(nsel) FSPH_LAUNCH(k_select, blocks(cnt), 256, s.stream, mask, f, ex, cnt, bidx.as<uint32_t>(), bmask.as<uint32_t>());
      return (int64_t)nsel;
    };
    tdbg[2] = now();
    const int64_t nsel_b = select(emask, n, d.sel_idx, d.sel_mask);
    tdbg[3] = now();
    for(int q = 0; q < P; ++q) {
      if(!d.send_cnt[q]) continue;
      FSPH_LAUNCH(k_bit_flag, blocks(nsel_b), 256, s.stream, d.sel_mask.as<uint32_t>(), nsel_b, q, f);
      exclusive_scan_u32(s, f, ex, nsel_b, nullptr);
      FSPH_LAUNCH(k_compact_sel, blocks(nsel_b), 256, s.stream, f, ex, nsel_b, d.sel_idx.as<uint32_t>(), d.send_idx.as<uint32_t>() + d.send_off[q]);
    }
    tdbg[4] = now();
    const int64_t nsel_n = select(nmask, nn, d.sel_idx, d.sel_mask);
    tdbg[5] = now();
    for(int q = 0; q < P; ++q) {
      if(!rs_cnt[q]) continue;
      FSPH_LAUNCH(k_bit_flag, blocks(nsel_n), 256, s.stream, d.sel_mask.as<uint32_t>(), nsel_n, q, f);
      exclusive_scan_u32(s, f, ex, nsel_n, nullptr);
      FSPH_LAUNCH((k_emit_records<D>), blocks(nsel_n), 256, s.stream, s.tree, s.col_key.as<uint64_t>(), nsel_n, d.sel_idx.as<uint32_t>(), f, ex,
        d.recs_send.as<NodeBoxRec>() + rs_off[q]);
    }
  }

  if(dd_debug) {
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    tdbg[6] = now();
    fprintf(stderr, "[dd rank %d] n %lld nn %d nsend %zu nrecs %zu | host ms: counts+sync %.2f reserve %.2f select_b %.2f loop_b %.2f select_n %.2f loop_n+sync %.2f\n",
      R, (long long)n, nn, nsend, nrs, tdbg[1] - tdbg[0], tdbg[2] - tdbg[1], tdbg[3] - tdbg[2], tdbg[4] - tdbg[3], tdbg[5] - tdbg[4], tdbg[6] - tdbg[5]);
  }
  ph.reset(new PhaseScope(s, "dd:halo_exchange"));
  // 7. halo exchange: counts, then every live column straight into the tail of the column, then the node records
  {
    std::vector<int64_t> mine((size_t)2 * P), all((size_t)2 * P * P);
    for(int q = 0; q < P; ++q) { mine[q] = (int64_t)d.send_cnt[q]; mine[P + q] = (int64_t)rs_cnt[q]; }
    dist_all_gather_host(s, mine.data(), all.data(), (size_t)2 * P * 8);
    size_t ng = 0, nrr = 0;
    for(int q = 0; q < P; ++q) {
      d.recv_cnt[q] = (size_t)all[(size_t)q * 2 * P + R]; d.recv_off[q] = ng; ng += d.recv_cnt[q];
      rr_cnt[q] = (size_t)all[(size_t)q * 2 * P + P + R]; rr_off[q] = nrr; nrr += rr_cnt[q];
    }
    d.nghost = (int64_t)ng;
    const int64_t n_total = s.n_owned + d.nghost;
    if(n_total >= (int64_t(1) << 31)) throw std::runtime_error("owned + ghost bodies exceed 2^31 on this rank");
    // the columns of the bodies a peer needs travel in ONE all-to-all-v: [peer][column][body].  Inside fsph_step only
    // what the tree, the search and the density sweep read from a neighbour (the rest is re-sent by reset_ghosts before
    // the sweep that reads it); outside it every live column (the caller may run any sweep next).
    std::vector<std::pair<DevBuf *, int>> cols, all_cols;
    live_columns(s, all_cols);
    if(s.in_step) halo_columns(s, cols);
    else cols = all_cols;
    for(auto & c : all_cols) c.first->grow_keep((size_t)n_total * c.second, (size_t)s.n_owned * c.second, s.stream);
    std::vector<size_t> so(P), sc(P), ro(P), rc(P);
    ColPtrs8 p8;
    ColPtrs4 p4;
    int n8 = 0, n4 = 0;
    for(auto & c : cols) {
      c.first->grow_keep((size_t)n_total * c.second, (size_t)s.n_owned * c.second, s.stream);
      if(c.second == 8) p8.p[n8++] = c.first->as<uint64_t>();
      else p4.p[n4++] = c.first->as<uint32_t>();
    }
    const size_t row = ((size_t)8 * n8 + (size_t)4 * n4 + 7) & ~(size_t)7;   // bytes per body (blocks stay 8-byte aligned)
    d.sendbuf.reserve_roomy(std::max<size_t>(nsend, 1) * row);
    d.recvbuf.reserve_roomy(std::max<size_t>(ng, 1) * row);
    for(int q = 0; q < P; ++q) {
      const size_t cnt = d.send_cnt[q];
      so[q] = d.send_off[q] * row; sc[q] = cnt * row; ro[q] = d.recv_off[q] * row; rc[q] = d.recv_cnt[q] * row;
      if(!cnt) continue;
      char * base = d.sendbuf.as<char>() + so[q];
      const uint32_t * idx = d.send_idx.as<uint32_t>() + d.send_off[q];
      FSPH_LAUNCH((k_pack_cols<uint64_t, ColPtrs8>), dim3(blocks(cnt), (unsigned)n8), 256, s.stream, p8, idx, (int64_t)cnt, reinterpret_cast<uint64_t *>(base));
      FSPH_LAUNCH((k_pack_cols<uint32_t, ColPtrs4>), dim3(blocks(cnt), (unsigned)n4), 256, s.stream, p4, idx, (int64_t)cnt,
        reinterpret_cast<uint32_t *>(base + cnt * 8 * n8));
    }
    cm.all_to_all_v(d.sendbuf.p, so.data(), sc.data(), d.recvbuf.p, ro.data(), rc.data(), s.stream);
    for(int q = 0; q < P; ++q) {
      const size_t cnt = d.recv_cnt[q];
      if(!cnt) continue;
      const char * base = d.recvbuf.as<char>() + ro[q];
      const int64_t dst = s.n_owned + (int64_t)d.recv_off[q];
      FSPH_LAUNCH((k_unpack_cols<uint64_t, ColPtrs8>), dim3(blocks(cnt), (unsigned)n8), 256, s.stream, p8, dst, (int64_t)cnt, reinterpret_cast<const uint64_t *>(base));
      FSPH_LAUNCH((k_unpack_cols<uint32_t, ColPtrs4>), dim3(blocks(cnt), (unsigned)n4), 256, s.stream, p4, dst, (int64_t)cnt,
        reinterpret_cast<const uint32_t *>(base + cnt * 8 * n8));
    }
    d.recs_recv.reserve_roomy(std::max<size_t>(nrr, 1) * sizeof(NodeBoxRec));
    for(int q = 0; q < P; ++q) {
      so[q] = rs_off[q] * sizeof(NodeBoxRec); sc[q] = rs_cnt[q] * sizeof(NodeBoxRec);
      ro[q] = rr_off[q] * sizeof(NodeBoxRec); rc[q] = rr_cnt[q] * sizeof(NodeBoxRec);
    }
    cm.all_to_all_v(d.recs_send.p, so.data(), sc.data(), d.recs_recv.p, ro.data(), rc.data(), s.stream);
    if(d.nghost) FSPH_LAUNCH(k_mark_ghost, blocks(d.nghost), 256, s.stream, s.col_type.as<int32_t>(), s.n_owned, n_total);
    ensure_capacity(s, n_total);

    ph.reset(new PhaseScope(s, "dd:merge_sort"));
    // 8. merge owned + ghosts, remember where every ghost went and where the bodies we serve to peers went
    //    Ranks own increasing key ranges and every sender lists its bodies in key order, so the merged key order is
    //    simply [ghosts of lower ranks | owned | ghosts of higher ranks]: a known permutation, no sort.
    s.n = n_total;
    s.has_ghosts = d.nghost > 0;
    int64_t n_glow = 0;
    for(int q = 0; q < R; ++q) n_glow += (int64_t)d.recv_cnt[q];
    s.own_begin = n_glow;
    d.ghost_pos.reserve_roomy(std::max<size_t>(ng, 1) * 4);
    if(ng) FSPH_LAUNCH(k_ghost_pos, blocks((int64_t)ng), 256, s.stream, d.ghost_pos.as<uint32_t>(), (int64_t)ng, n_glow, s.n_owned);
    if(n_glow > 0) {
      s.b_perm.reserve((size_t)n_total * 4);
      FSPH_LAUNCH(k_merge_perm, blocks(n_total), 256, s.stream, s.b_perm.as<uint32_t>(), n_total, n_glow, s.n_owned);
      gather_bodies(s, s.b_perm.as<uint32_t>(), n_total);
      if(nsend) FSPH_LAUNCH(k_shift_idx, blocks(nsend), 256, s.stream, d.send_idx.as<uint32_t>(), (int64_t)nsend, (uint32_t)n_glow);
    }

    ph.reset(new PhaseScope(s, "dd:tree_merged"));
    // 9. merged tree with the owners' leaf depths; global boxes for every node that holds remote bodies
    TreeEdges tm;
    tm.gd_in = s.col_gd.as<int32_t>();
    tm.finish = false;
    build_tree_edges(s, tm);
    apply_records<D>(s, d.path_all.as<NodeBoxRec>(), (int64_t)(1 + DD_MAX_PATH) * P, P);
    apply_records<D>(s, d.recs_recv.as<NodeBoxRec>(), (int64_t)nrr, 0);
    FSPH_LAUNCH((k_fix_straddling<D>), blocks(s.tree.nnodes), 256, s.stream, s.tree, s.col_key.as<uint64_t>(), s.tree.nnodes, klo, khi_incl);
    finish_tree(s);
  }
  ph.reset();
  FSPH_CUDA_OK(cudaGetLastError());
}

void dist_update_iteration(System & s) {
  if(s.dim == 1) dist_update_iteration_t<1>(s);
  else if(s.dim == 2) dist_update_iteration_t<2>(s);
  else dist_update_iteration_t<3>(s);
}

// ---- FMM gravity on several ranks ------------------------------------------------------------------------------
// Replaces the remote-branch traffic of traversal_fmm (tree_topology.h:583-612, 1384-1463: REQUEST_SUBTREE).  Gravity
// needs of a body only (x, m) and of a node only (mass, centre, radius): every rank all-gathers the owned bodies' keys,
// positions, masses and h (40 B per body over NVLink), builds the GLOBAL tree -- the ranks' ranges concatenate in key
// order, so no sort -- and runs the dual walk restricted to the sink cells that hold its own bodies ("iam_owner" of the
// reference).  Sink nodes that straddle ranks are evaluated by each rank that shares them (same sources, same result).
// The tree and hence the interaction set are those of a single rank: parity with one rank by construction.
__global__ void __launch_bounds__(256) k_add_range(double * __restrict__ dst, const double * __restrict__ src, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) dst[i] += src[i];
}

void dist_gravitation_fmm(System & s) {
  if(s.dim != 3) throw std::runtime_error("fmm needs 3 dimensions (bodies_system.h:283)");
  if(!s.tree_valid) throw std::runtime_error("gravitation_fmm before update_iteration");
  Dist & d = dist_of(s);
  Comm & cm = *s.comm;
  const int P = cm.nranks, R = cm.rank;
  const int64_t n_own = s.n_owned_or_n(), own_begin = s.has_ghosts ? s.own_begin : 0;
  std::vector<int64_t> cnt(P);
  dist_all_gather_host(s, &n_own, cnt.data(), 8);
  std::vector<int64_t> goff(P + 1, 0);
  for(int q = 0; q < P; ++q) goff[q + 1] = goff[q] + cnt[q];
  const int64_t N = goff[P];
  if(N >= (int64_t(1) << 31)) throw std::runtime_error("more than 2^31 bodies in the gathered gravity tree");
  if(!d.fmm_sys) {
    d.fmm_sys = new System;
    d.fmm_sys->b_status.reserve(128);
    FSPH_CUDA_OK(cudaMemsetAsync(d.fmm_sys->b_status.p, 0, 128, s.stream));
  }
  System & g = *d.fmm_sys;
  g.par = s.par;
  g.dim = s.dim;
  g.device = s.device;
  g.stream = s.stream;
  g.n = g.n_owned = N;
  const int fields[] = {F_POS, F_POS + 1, F_POS + 2, F_MASS, F_H};
  std::vector<size_t> so(P), sc(P), ro(P), rc(P);
  for(int q = 0; q < P; ++q) { so[q] = 0; sc[q] = (size_t)n_own * 8; ro[q] = (size_t)goff[q] * 8; rc[q] = (size_t)cnt[q] * 8; }
  for(int f : fields) {
    g.col[f].reserve((size_t)N * 8);
    cm.all_to_all_v(s.col[f].as<char>() + (size_t)own_begin * 8, so.data(), sc.data(), g.col[f].p, ro.data(), rc.data(), s.stream);
  }
  g.col_key.reserve((size_t)N * 8);
  cm.all_to_all_v(s.col_key.as<char>() + (size_t)own_begin * 8, so.data(), sc.data(), g.col_key.p, ro.data(), rc.data(), s.stream);
  g.col_type.reserve((size_t)N * 4);
  FSPH_CUDA_OK(cudaMemsetAsync(g.col_type.p, 0, (size_t)N * 4, s.stream));
  const int acc[] = {F_GACC, F_GACC + 1, F_GACC + 2, F_GPOT};
  for(int f : acc) {
    g.col[f].reserve((size_t)N * 8);
    FSPH_CUDA_OK(cudaMemsetAsync(g.col[f].p, 0, (size_t)N * 8, s.stream));
  }
  TreeEdges te;
  te.finish = false;
  build_tree_edges(g, te);
  gravitation_fmm_range(g, goff[R], goff[R + 1]);
  for(int f : acc)
    FSPH_LAUNCH(k_add_range, blocks(n_own), 256, s.stream, s.col[f].as<double>() + own_begin, g.col[f].as<double>() + goff[R], n_own);
  s.fmm_pairs = g.fmm_pairs;
  FSPH_CUDA_OK(cudaGetLastError());
}

// reset_ghosts: the owners re-send the columns the sweeps read from neighbours and that change between sweeps
// (density, pressure, both sound speeds, velocity, half-step velocity); positions, h and masses are unchanged.
void dist_refresh_ghosts(System & s) {
  if(!s.comm) return;
  Dist & d = dist_of(s);
  if(!s.tree_valid) throw std::runtime_error("reset_ghosts before update_iteration");
  const int P = s.comm->nranks;
  std::vector<int> fields = {F_RHO, F_P, F_CS, F_CSOLD};
  if(s.par.sph_viscosity == FSPH_VISC_CULLEN) { fields.push_back(F_ALPHA); fields.push_back(F_DIVV); }   // read from neighbours
  for(int k = 0; k < s.dim; ++k) { fields.push_back(F_VEL + k); fields.push_back(F_VELH + k); }
  const size_t nf = fields.size();
  const size_t nsend = (size_t)d.nsend, ng = (size_t)d.nghost;
  d.sendbuf.reserve_roomy(std::max<size_t>(nsend * nf, 1) * 8);
  d.recvbuf.reserve_roomy(std::max<size_t>(ng * nf, 1) * 8);
  // layout: [peer][field][body of that peer's list]
  if(nf > 16) throw std::runtime_error("too many refresh fields");
  FieldPtrs fp;
  for(size_t j = 0; j < nf; ++j) fp.p[j] = s.col[fields[j]].as<uint64_t>();
  for(int q = 0; q < P; ++q) {
    const size_t cnt = d.send_cnt[q];
    if(!cnt) continue;
    FSPH_LAUNCH(k_pack_fields, dim3(blocks(cnt), (unsigned)nf), 256, s.stream, fp, d.send_idx.as<uint32_t>() + d.send_off[q], (int64_t)cnt,
      d.sendbuf.as<uint64_t>() + d.send_off[q] * nf);
  }
  std::vector<size_t> so(P), sc(P), ro(P), rc(P);
  for(int q = 0; q < P; ++q) {
    so[q] = d.send_off[q] * nf * 8; sc[q] = d.send_cnt[q] * nf * 8;
    ro[q] = d.recv_off[q] * nf * 8; rc[q] = d.recv_cnt[q] * nf * 8;
  }
  s.comm->all_to_all_v(d.sendbuf.p, so.data(), sc.data(), d.recvbuf.p, ro.data(), rc.data(), s.stream);
  for(int q = 0; q < P; ++q) {
    const size_t cnt = d.recv_cnt[q];
    if(!cnt) continue;
    FSPH_LAUNCH(k_unpack_fields, dim3(blocks(cnt), (unsigned)nf), 256, s.stream, fp, d.ghost_pos.as<uint32_t>() + d.recv_off[q], (int64_t)cnt,
      d.recvbuf.as<uint64_t>() + d.recv_off[q] * nf);
  }
  FSPH_CUDA_OK(cudaGetLastError());
}

}  // namespace fsph
