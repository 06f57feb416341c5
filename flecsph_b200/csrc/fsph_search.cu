// fsph_search.cu -- smoothing-length neighbour search (compiled with -fmad=false: bit-exact predicates).
//
// Replaces tree_topology::traversal_sph (include/tree_topology/tree_topology.h:212-442) with the predicates of
// tree_geometry (include/tree_topology/tree_geometry.h: within_distance2 :60,161,265; intersects_box_box
// :118,198,306; intersects_sphere_box :107,219,329).
//
// What the reference computes.  For a work cell G (maximal sub-tree of <= 128 bodies, :221-235) a tree node C is
// accepted iff its box overlaps G's box and some body sphere of G touches it (:327-335); the leaves under accepted
// nodes that pass |x_e - X_G| <= max(h_e, lap_G) + radius_G (:365-370) are G's candidates; body k of G gets the
// candidates within max(h_k, h_e) (:372-378).  Both node predicates are monotone in the box and boxes are nested, so
// "C is accepted" is a function of (C, G) alone: the candidates of G are the leaf children, passing the prefilter, of
// the nodes that pass the two tests -- however those nodes are enumerated.  (The tests are NOT conservative for
// unequal h, SURVEY.md 7-1, which is why they are evaluated here exactly as the reference does, per work cell.)
//
// How it is computed here.
//   1. Super cells (maximal sub-trees of <= SUPER_ENTITIES bodies) walk the tree once each (k_super_walk) with a
//      conservative test -- node box within the largest h of the super cell's tight position box -- and keep the flat
//      list of passing nodes that have leaf children.
//   2. Blocks: runs of consecutive work cells of one super cell with <= 256 bodies and <= 16 cells.  One CTA of 256
//      threads per block, one body per thread (k_block_search):
//        a. the super cell's list is filtered with the same conservative test against the block's tight box;
//        b. every surviving node is tested exactly against every work cell of the block -> 16-bit cell mask;
//        c. the leaf children of accepted nodes are tested against those cells (prefilter + a conservative
//           tight-box cull that drops only leaves no body of the cell can pair with) -> the block's union candidate
//           list, each candidate with the mask of the cells it belongs to;
//        d. candidates are staged in shared memory in tiles of 512 (x, h^2 from the packed leaf records) and every
//           body tests the candidates of its own cell with the exact fp64 pair predicate
//           (dx^2+dy^2)+dz^2 <= max(h_a,h_b)^2.
//   3. Output (reused by every sweep of the step: positions and h do not change between them): per block the
//      candidate indices, and per (tile, warp) a neighbour list stored as rows of 32 x uint16 -- row 0 the lanes'
//      counts, row 1 + j the tile-local candidate position of each lane's j-th neighbour -- so that the sweeps read
//      their lists with coalesced 64-byte loads and run a uniform loop without decoding bit masks.
// Candidate, directory and list storage is bump-allocated (one atomicAdd per block / tile) against a capacity taken
// from the previous search; when it does not fit, the cursors still count and the search is repeated once with the
// exact sizes.  Per-block results do not depend on where they were placed, so the sums of the sweeps are
// reproducible bit for bit.
#include <stdexcept>
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include "fsph_internal.h"

namespace fsph {

constexpr int SW_QCAP = 16384;  // node frontier capacity per level of the super-cell walk (per-CTA strip of global memory)
constexpr int SW_NT = 256;      // threads per CTA of the super-cell walk
constexpr int SR_NT = BLK_BODIES;
constexpr int SR_NW = SR_NT / 32;
#ifndef FSPH_SR_LROWS
#define FSPH_SR_LROWS 48
#endif
#ifndef FSPH_SR_OCC
#define FSPH_SR_OCC 3
#endif
constexpr int SR_LROWS = FSPH_SR_LROWS;   // list rows staged per warp and tile in shared memory (longer lists: slow path)
constexpr double GEO_TOL = 2.220446049250313e-16 * 10.;  // tree_geometry.h:57

template<int D>
__device__ __forceinline__ bool sphere_box(const double * mn, const double * mx, const double * c, double r) {
  if(D == 3) {
    // tree_geometry.h:329-341
    double s = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      double x = c[d] < mx[d] ? c[d] : mx[d];
      x = x < mn[d] ? mn[d] : x;
      double t = (x - c[d]) * (x - c[d]);
      s = d == 0 ? t : s + t;
    }
    return s <= r * r;
  }
  else if(D == 2) {
    // tree_geometry.h:219-227
    double s = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      double x = fmax(mn[d], fmin(c[d], mx[d]));
      s = s + (x - c[d]) * (x - c[d]);
    }
    return sqrt(s) - r <= GEO_TOL;
  }
  else {
    // tree_geometry.h:107-114
    double x = fmax(mn[0], fmin(c[0], mx[0]));
    return fabs(x - c[0]) - r <= GEO_TOL;
  }
}

template<int NT>
__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t & total, uint32_t * sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if(lane >= o) x += y;
  }
  if(lane == 31) sm[w] = x;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for(int i = 0; i < NT / 32; ++i) {
    if(i < w) base += sm[i];
    tot += sm[i];
  }
  total = tot;
  __syncthreads();
  return base + x - v;
}

// max / min of finite doubles: one compare + select (fmax / fmin also order NaNs, which costs several more instructions)
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

// squared distance between two axis-aligned boxes (0 when they overlap)
template<int D>
__device__ __forceinline__ double box_gap2(const double * amn, const double * amx, const double * bmn, const double * bmx) {
  double dd = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) {
    const double tt = dmax(0.0, dmax(amn[d] - bmx[d], bmn[d] - amx[d]));
    dd = dd + tt * tt;
  }
  return dd;
}

// ---- super cells: one flat node list each -------------------------------------------------------------------
// A super cell S walks the tree once with a test that every node any work cell G inside S can accept must pass:
// G accepts C only if some body sphere of G touches C's box (intersects_sphere_box, tree_topology.h:348), hence the
// distance between C's box and the tight box of S's positions is at most the largest h in S.  The nodes that pass
// and have leaf children are S's list (in walk order: fixed).  MODE 0 counts, MODE 1 fills (offsets: prefix sum).
// MODE 2 does it in one pass when the previous search left an estimate of the total: the list is staged in a per-CTA strip
// of global memory, the super cell reserves its share of the arrays with one atomicAdd when its walk is done and copies
// the strip over (lists land in completion order; sl_off / sl_len say where).  A list longer than the strip or a total
// beyond the capacity raises ERR_SEARCH_CAPACITY and the host falls back to the two passes.
constexpr int SW_STAGE = 16384;   // entries of a CTA's staging strip
template<int D, int MODE>
__global__ void __launch_bounds__(SW_NT) k_super_walk(Tree t, Cols c, Search sr, const int32_t * __restrict__ super_first_wc, int32_t * __restrict__ gqueue,
  uint32_t * __restrict__ status, int32_t * __restrict__ gstage, unsigned long long * __restrict__ sl_total, unsigned long long sl_cap) {
  __shared__ unsigned long long base_sm;
  int32_t * const stage = MODE == 2 ? gstage + (size_t)blockIdx.x * SW_STAGE : nullptr;
  constexpr int NT = SW_NT;
  __shared__ uint32_t scan_sm[NT / 32];
  __shared__ double red_sm[NT / 32][2 * D + 1];
  // the node frontier lives in a per-CTA strip of global memory (two levels of SW_QCAP nodes + one result byte per node):
  // its size depends on the data (bodies per leaf-parent node, kernel width), and the walk is a small part of the search
  int32_t * const queue0 = gqueue + (size_t)blockIdx.x * (2 * SW_QCAP + SW_QCAP / 4);
  uint8_t * const qres = reinterpret_cast<uint8_t *>(queue0 + 2 * SW_QCAP);
  const int tid = threadIdx.x;
  for(int sidx = blockIdx.x; sidx < sr.nsuper; sidx += gridDim.x) {
    __syncthreads();
    if(super_first_wc[sidx] < 0) {
      // no work cell of this rank inside (remote copies only): no list
      if(MODE == 0 && tid == 0) sr.sl_off[sidx] = 0;
      if(MODE == 2 && tid == 0) { sr.sl_off[sidx] = 0; sr.sl_len[sidx] = 0; }
      continue;
    }
    int start, P;
    {
      const int v = t.cell[sr.super_cell[sidx]];
      if(v >= 0) { start = v; P = 1; }
      else { start = t.node_start[~v]; P = t.sub[~v]; }
    }
    double tmn[D], tmx[D], hmax = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) { tmn[d] = 1.79769313486231570e308; tmx[d] = -1.79769313486231570e308; }
    for(int k = tid; k < P; k += NT) {
#pragma unroll
      for(int d = 0; d < D; ++d) {
        const double x = c.f[F_POS + d][start + k];
        tmn[d] = fmin(tmn[d], x);
        tmx[d] = fmax(tmx[d], x);
      }
      hmax = fmax(hmax, c.f[F_H][start + k]);
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for(int d = 0; d < D; ++d) {
        tmn[d] = fmin(tmn[d], __shfl_xor_sync(0xffffffffu, tmn[d], o));
        tmx[d] = fmax(tmx[d], __shfl_xor_sync(0xffffffffu, tmx[d], o));
      }
      hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    }
    if((tid & 31) == 0) {
#pragma unroll
      for(int d = 0; d < D; ++d) { red_sm[tid >> 5][d] = tmn[d]; red_sm[tid >> 5][D + d] = tmx[d]; }
      red_sm[tid >> 5][2 * D] = hmax;
    }
    if(tid == 0) queue0[0] = 0;  // the root node
    __syncthreads();
#pragma unroll
    for(int ww = 0; ww < NT / 32; ++ww) {
#pragma unroll
      for(int d = 0; d < D; ++d) {
        tmn[d] = fmin(tmn[d], red_sm[ww][d]);
        tmx[d] = fmax(tmx[d], red_sm[ww][D + d]);
      }
      hmax = fmax(hmax, red_sm[ww][2 * D]);
    }
    const double hmax2 = hmax * hmax * (1.0 + 1e-9);
    int32_t * out = MODE == 1 ? sr.sl_nodes + sr.sl_off[sidx] : nullptr;
    SlRec * out_rec = MODE == 1 ? sr.sl_rec + sr.sl_off[sidx] : nullptr;
    int ncur = 1, cur = 0;
    uint32_t nlist = 0;
    while(ncur > 0) {
      const int32_t * qc = queue0 + cur * SW_QCAP;
      int32_t * qn = queue0 + (cur ^ 1) * SW_QCAP;
      uint32_t my_nodes = 0, my_list = 0;
      for(int j = tid; j < ncur; j += NT) {
        const NodeRec rec = t.noderec[qc[j]];
        uint32_t r = 0;
        if(box_gap2<D>(rec.bmin, rec.bmax, tmn, tmx) <= hmax2) {
          const uint32_t node_mask = ((1u << rec.nchild) - 1u) & ~rec.leafmask;
          r = (uint32_t)__popc(node_mask) | (rec.leafmask ? 0x80u : 0u);
        }
        qres[j] = (uint8_t)r;
        my_nodes += r & 0x3fu;
        my_list += (r >> 7) & 1u;
      }
      uint32_t nnext_raw, nlist_add;
      uint32_t qpos = block_scan_u32<NT>(my_nodes, nnext_raw, scan_sm);
      uint32_t lpos = nlist + block_scan_u32<NT>(my_list, nlist_add, scan_sm);
      for(int j = tid; j < ncur; j += NT) {
        const uint32_t r = qres[j];
        if(r == 0) continue;
        const int nd = qc[j];
        if(r & 0x3fu) {
          const NodeRec rec = t.noderec[nd];
          uint32_t node_mask = ((1u << rec.nchild) - 1u) & ~rec.leafmask;
          while(node_mask) {
            const int b = __ffs(node_mask) - 1;
            node_mask &= node_mask - 1;
            // node children of a node have consecutive node indices (level order)
            if(qpos < SW_QCAP) qn[qpos] = rec.first_child_node + __popc(~rec.leafmask & ((1u << b) - 1u));
            ++qpos;
          }
        }
        if(r & 0x80u) {
          if(MODE == 1) {
            out[lpos] = nd;
            const NodeRec rec = t.noderec[nd];
            SlRec sl;
#pragma unroll
            for(int d = 0; d < 3; ++d) {
              sl.mn[d] = d < D ? __double2float_rd(rec.bmin[d]) : 0.f;
              sl.mx[d] = d < D ? __double2float_ru(rec.bmax[d]) : 0.f;
            }
            sl.node = nd;
            sl.pad = 0;
            out_rec[lpos] = sl;
          }
          if(MODE == 2 && lpos < (uint32_t)SW_STAGE) stage[lpos] = nd;
          ++lpos;
        }
      }
      nlist += nlist_add;
      uint32_t nnext = nnext_raw;
      if(nnext > SW_QCAP) {
        if(tid == 0) atomicOr(status, (uint32_t)ERR_QUEUE_OVERFLOW);
        nnext = SW_QCAP;
      }
      __syncthreads();
      ncur = (int)nnext;
      cur ^= 1;
    }
    if(MODE == 0 && tid == 0) sr.sl_off[sidx] = nlist;
    if(MODE == 2) {
      if(tid == 0) base_sm = atomicAdd(sl_total, (unsigned long long)nlist);
      __syncthreads();
      const unsigned long long base = base_sm;
      const bool fits = nlist <= (uint32_t)SW_STAGE && base + nlist <= sl_cap;
      if(tid == 0) {
        sr.sl_off[sidx] = base;
        sr.sl_len[sidx] = fits ? (int32_t)nlist : 0;
        if(!fits) atomicOr(status, (uint32_t)ERR_SEARCH_CAPACITY);
      }
      if(fits)
        for(uint32_t i = tid; i < nlist; i += NT) {
          const int nd = stage[i];
          const NodeRec rec = t.noderec[nd];
          SlRec sl;
#pragma unroll
          for(int d = 0; d < 3; ++d) {
            sl.mn[d] = d < D ? __double2float_rd(rec.bmin[d]) : 0.f;
            sl.mx[d] = d < D ? __double2float_ru(rec.bmax[d]) : 0.f;
          }
          sl.node = nd;
          sl.pad = 0;
          sr.sl_nodes[base + i] = nd;
          sr.sl_rec[base + i] = sl;
        }
    }
  }
}

__global__ void __launch_bounds__(256) k_sl_len(const uint64_t * __restrict__ off, int32_t * __restrict__ len, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) len[i] = (int32_t)(off[i + 1] - off[i]);
}

__global__ void __launch_bounds__(256) k_super_compact(const uint32_t * __restrict__ at, const uint32_t * __restrict__ ex, int64_t n,
  int32_t * __restrict__ super_cell) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n && at[i]) super_cell[ex[i]] = (int32_t)(at[i] - 1u);
}
__global__ void __launch_bounds__(256) k_flag_set(const uint32_t * __restrict__ in, uint32_t * __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = in[i] != 0;
}
// super cell of a work cell = the one whose body range holds the work cell's first body; the first work cell of each
// super cell is recorded for the block packer
__global__ void __launch_bounds__(256) k_wc_super(Tree t, Search sr, const uint32_t * __restrict__ at, const uint32_t * __restrict__ ex,
  int32_t * __restrict__ super_first_wc) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= sr.nwc) return;
  const int v = t.cell[sr.wc_cell[g]];
  const int start = v >= 0 ? v : t.node_start[~v];
  const int su = (int32_t)(ex[start] + (at[start] ? 1u : 0u)) - 1;
  sr.wc_super[g] = su;
  // work cells are in Morton order, so those of one super cell are consecutive: the first one has a start inside the
  // super cell and either is work cell 0 or follows a cell of another super cell
  bool first = g == 0;
  if(!first) {
    const int vp = t.cell[sr.wc_cell[g - 1]];
    const int sp = vp >= 0 ? vp : t.node_start[~vp];
    first = ((int32_t)(ex[sp] + (at[sp] ? 1u : 0u)) - 1) != su;
  }
  if(first) super_first_wc[su] = g;
}

// ---- blocks: greedy packing of the consecutive work cells of a super cell --------------------------------------
// MODE 0: count the blocks of each super cell; MODE 1: write their first work cells (offsets: prefix sum of the counts)
template<int MODE>
__global__ void __launch_bounds__(128) k_block_pack(Tree t, Search sr, const int32_t * __restrict__ super_first_wc, uint32_t * __restrict__ cnt,
  const uint32_t * __restrict__ off) {
  const int su = blockIdx.x * blockDim.x + threadIdx.x;
  if(su >= sr.nsuper) return;
  int g = super_first_wc[su];
  uint32_t nb = 0;
  if(g >= 0) {
    int bodies = 0, cells = 0, next_body = -1;
    for(; g < sr.nwc && sr.wc_super[g] == su; ++g) {
      const int v = t.cell[sr.wc_cell[g]];
      const int start = v >= 0 ? v : t.node_start[~v];
      const int P = v >= 0 ? 1 : t.sub[~v];
      // a block is a run of body-contiguous cells with <= BLK_BODIES bodies and <= BLK_CELLS cells
      if(cells == 0 || bodies + P > BLK_BODIES || cells == BLK_CELLS || start != next_body) {
        if(MODE == 1) sr.blk_wc[off[su] + nb] = g;
        ++nb;
        bodies = 0;
        cells = 0;
      }
      bodies += P;
      ++cells;
      next_body = start + P;
    }
  }
  if(MODE == 0) cnt[su] = nb;
}

template<int D>
static void build_super_lists(System & s) {
  Search & sr = s.search;
  const int64_t n = s.n;
  uint32_t * status = s.b_status.as<uint32_t>();
  s.b_cnt0.reserve((size_t)(n + 1) * 4);
  s.b_tmp[0].reserve((size_t)(n + 1) * 4);
  uint32_t * at = s.b_cnt0.as<uint32_t>();
  uint32_t * ex = s.b_tmp[0].as<uint32_t>();
  mark_maximal_cells(s, SUPER_ENTITIES, false, at);
  unsigned nbp = (unsigned)((n + 255) / 256);
  FSPH_LAUNCH(k_flag_set, nbp, 256, s.stream, at, ex, n);
  exclusive_scan_u32(s, ex, ex, n, nullptr);
  uint32_t nsuper = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&nsuper, ex + n, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(nsuper == 0) throw std::runtime_error("no super cells");
  s.b_super_cell.reserve((size_t)nsuper * 4);
  s.b_sl_off.reserve((size_t)(nsuper + 1) * 8);
  s.b_wc_super.reserve((size_t)std::max(sr.nwc, 1) * 4);
  s.b_tmp[1].reserve((size_t)nsuper * 4);
  s.b_tmp[2].reserve((size_t)(nsuper + 1) * 4);
  sr.super_cell = s.b_super_cell.as<int32_t>();
  sr.sl_off = s.b_sl_off.as<uint64_t>();
  sr.wc_super = s.b_wc_super.as<int32_t>();
  sr.nsuper = (int)nsuper;
  int32_t * first_wc = s.b_tmp[1].as<int32_t>();
  uint32_t * bcnt = s.b_tmp[2].as<uint32_t>();
  FSPH_CUDA_OK(cudaMemsetAsync(first_wc, 0xff, (size_t)nsuper * 4, s.stream));
  FSPH_LAUNCH(k_super_compact, nbp, 256, s.stream, at, ex, n, sr.super_cell);
  FSPH_LAUNCH(k_wc_super, (unsigned)((sr.nwc + 255) / 256), 256, s.stream, s.tree, sr, at, ex, first_wc);
  // blocks
  FSPH_LAUNCH((k_block_pack<0>), (nsuper + 127) / 128, 128, s.stream, s.tree, sr, first_wc, bcnt, nullptr);
  exclusive_scan_u32(s, bcnt, bcnt, nsuper, nullptr);
  uint32_t st = 0, nblk = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&nblk, bcnt + nsuper, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(nblk == 0) throw std::runtime_error("no blocks");
  s.b_blk_wc.reserve((size_t)(nblk + 1) * 4);
  sr.blk_wc = s.b_blk_wc.as<int32_t>();
  sr.nblk = (int)nblk;
  FSPH_LAUNCH((k_block_pack<1>), (nsuper + 127) / 128, 128, s.stream, s.tree, sr, first_wc, nullptr, bcnt);
  // node lists
  const int sw_grid = (int)std::min<uint32_t>(nsuper, 148 * 6);
  s.b_super_queue.reserve((size_t)sw_grid * (2 * SW_QCAP + SW_QCAP / 4) * 4);
  s.b_sl_len.reserve((size_t)nsuper * 4);
  sr.sl_len = s.b_sl_len.as<int32_t>();
  int32_t * gq = s.b_super_queue.as<int32_t>();
  unsigned long long * sl_total = reinterpret_cast<unsigned long long *>(status + 24);
  if(s.est_sl > 0 && s.est_n > 0 && !getenv("FSPH_SL_TWO_PASS")) {
    // one pass, capacity from the previous search
    const double r = std::max(1.0, (double)s.n / (double)s.est_n);
    const unsigned long long cap = (unsigned long long)(s.est_sl * r * 1.25) + 65536;
    s.b_sl_nodes.reserve((size_t)cap * 4);
    s.b_sl_rec.reserve((size_t)cap * sizeof(SlRec));
    s.b_sl_stage.reserve((size_t)sw_grid * SW_STAGE * 4);
    sr.sl_nodes = s.b_sl_nodes.as<int32_t>();
    sr.sl_rec = s.b_sl_rec.as<SlRec>();
    FSPH_CUDA_OK(cudaMemsetAsync(sl_total, 0, 8, s.stream));
    FSPH_LAUNCH((k_super_walk<D, 2>), sw_grid, SW_NT, s.stream, s.tree, s.cols(), sr, first_wc, gq, status, s.b_sl_stage.as<int32_t>(), sl_total, cap);
    unsigned long long tot = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&tot, sl_total, 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(st & ERR_QUEUE_OVERFLOW) throw std::runtime_error("neighbour search: the node frontier of a super cell exceeds 16384 nodes per level");
    if(!(st & ERR_SEARCH_CAPACITY)) {
      s.est_sl = tot;
      FSPH_CUDA_OK(cudaGetLastError());
      return;
    }
    const uint32_t cleared = st & ~(uint32_t)ERR_SEARCH_CAPACITY;
    FSPH_CUDA_OK(cudaMemcpyAsync(status, &cleared, 4, cudaMemcpyHostToDevice, s.stream));
  }
  // two passes: count, offsets, fill
  FSPH_LAUNCH((k_super_walk<D, 0>), sw_grid, SW_NT, s.stream, s.tree, s.cols(), sr, first_wc, gq, status, nullptr, nullptr, 0ull);
  exclusive_scan_u64(s, sr.sl_off, sr.sl_off, nsuper, nullptr);
  uint64_t tot = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&tot, sr.sl_off + nsuper, 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(st & ERR_QUEUE_OVERFLOW) throw std::runtime_error("neighbour search: the node frontier of a super cell exceeds 16384 nodes per level");
  s.b_sl_nodes.reserve((size_t)std::max<uint64_t>(tot, 1) * 4);
  sr.sl_nodes = s.b_sl_nodes.as<int32_t>();
  s.b_sl_rec.reserve((size_t)std::max<uint64_t>(tot, 1) * sizeof(SlRec));
  sr.sl_rec = s.b_sl_rec.as<SlRec>();
  FSPH_LAUNCH(k_sl_len, (nsuper + 255) / 256, 256, s.stream, sr.sl_off, sr.sl_len, (int)nsuper);
  FSPH_LAUNCH((k_super_walk<D, 1>), sw_grid, SW_NT, s.stream, s.tree, s.cols(), sr, first_wc, gq, status, nullptr, nullptr, 0ull);
  s.est_sl = tot;
  FSPH_CUDA_OK(cudaGetLastError());
}

// ---- work-cell records -------------------------------------------------------------------------------------------
// Everything the walk needs of a work cell G, computed once per search by k_wc_info (16 lanes per cell).
template<int D> struct __align__(16) WcInfo {
  double gmn[D], gmx[D];   // node box (intersects_box_box)
  double gcom[D], glap, grad;
  double tmn[D], tmx[D];   // tight box of the cell's positions
  double hmax, hmax2;      // largest h of the cell, and its square with the safety margin of the quick reject
  int32_t first, P;        // first body / body count
  int32_t has_node;        // 0: lone leaf under a bigger node (no box test, no prefilter: tree_topology.h:326,364)
  uint8_t rep[1 << D];     // per corner of the tight box: the nearest body of the cell (offset from `first`)
};

struct SearchCursors {
  unsigned long long cand, dir, rows;
};

template<int D>
__global__ void __launch_bounds__(256) k_wc_info(Tree t, Cols c, Search sr, WcInfo<D> * __restrict__ out, LeafRec * __restrict__ reps) {
  constexpr int NCH = 1 << D;
  const int g = (int)((blockIdx.x * 256u + threadIdx.x) >> 4), l16 = threadIdx.x & 15;
  const bool live = g < sr.nwc;
  int v = 0, start = 0, P = 0;
  if(live) {
    v = t.cell[sr.wc_cell[g]];
    start = v >= 0 ? v : t.node_start[~v];
    P = v >= 0 ? 1 : t.sub[~v];
  }
  double tmn[D], tmx[D], hm = 0.0;
#pragma unroll
  for(int d = 0; d < D; ++d) { tmn[d] = 1.79769313486231570e308; tmx[d] = -1.79769313486231570e308; }
  for(int k = l16; k < P; k += 16) {
#pragma unroll
    for(int d = 0; d < D; ++d) {
      const double x = c.f[F_POS + d][start + k];
      tmn[d] = fmin(tmn[d], x);
      tmx[d] = fmax(tmx[d], x);
    }
    hm = fmax(hm, c.f[F_H][start + k]);
  }
#pragma unroll
  for(int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for(int d = 0; d < D; ++d) {
      tmn[d] = fmin(tmn[d], __shfl_xor_sync(0xffffffffu, tmn[d], o));
      tmx[d] = fmax(tmx[d], __shfl_xor_sync(0xffffffffu, tmx[d], o));
    }
    hm = fmax(hm, __shfl_xor_sync(0xffffffffu, hm, o));
  }
  // representative bodies: for each corner of the tight box the body of the cell nearest to it.  The "does any body
  // sphere of G touch this node box" loop tries the representative on the node's side first (pure ordering heuristic:
  // the outcome is the exact any-of-all-bodies result either way).
  uint8_t rep[NCH];
  {
    // one pass over the cell's bodies, all corners at once
    unsigned long long key[NCH];
#pragma unroll
    for(int q = 0; q < NCH; ++q) key[q] = ~0ull;
    for(int k = l16; k < P; k += 16) {
      double x[D];
#pragma unroll
      for(int d = 0; d < D; ++d) x[d] = c.f[F_POS + d][start + k];
#pragma unroll
      for(int q = 0; q < NCH; ++q) {
        double d2 = 0.0;
#pragma unroll
        for(int d = 0; d < D; ++d) {
          const double df = x[d] - (((q >> d) & 1) ? tmx[d] : tmn[d]);
          d2 += df * df;
        }
        const unsigned long long kk = ((unsigned long long)__double_as_longlong(d2) & ~0xffull) | (unsigned long long)k;
        key[q] = kk < key[q] ? kk : key[q];
      }
    }
#pragma unroll
    for(int q = 0; q < NCH; ++q) {
#pragma unroll
      for(int o = 8; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[q], o);
        key[q] = other < key[q] ? other : key[q];
      }
      rep[q] = (uint8_t)(key[q] & 0xffull);
    }
  }
  if(live && l16 == 0) {
    WcInfo<D> w;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      w.tmn[d] = tmn[d];
      w.tmx[d] = tmx[d];
      w.gmn[d] = w.gmx[d] = w.gcom[d] = 0.0;
    }
    w.glap = w.grad = 0.0;
    w.hmax = hm;
    w.hmax2 = hm * hm * (1.0 + 1e-9);
    w.first = start;
    w.P = P;
    w.has_node = v < 0;
    if(v < 0) {
      const int nd = ~v;
#pragma unroll
      for(int d = 0; d < D; ++d) { w.gmn[d] = t.bmin[d][nd]; w.gmx[d] = t.bmax[d][nd]; w.gcom[d] = t.com[d][nd]; }
      w.glap = t.lap[nd];
      w.grad = t.radius[nd];
    }
#pragma unroll
    for(int q = 0; q < NCH; ++q) w.rep[q] = rep[q];
    out[g] = w;
  }
  // position and h of the representatives, packed (one 32-byte load in the walk)
  if(live && l16 < NCH) {
    const int k = start + rep[l16 < NCH ? l16 : 0];
    LeafRec r;
    r.x[0] = r.x[1] = r.x[2] = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) r.x[d] = c.f[F_POS + d][k];
    r.h = c.f[F_H][k];
    reps[(size_t)g * NCH + l16] = r;
  }
}

// ---- the walk: one warp per block ------------------------------------------------------------------------------
// (a) the super cell's node list is filtered with the conservative test against the block's tight box, 32 nodes per
//     round, survivors queued in shared memory (per warp);
// (b) 32 queued nodes at a time, one per lane: the exact tests against every work cell of the block -> cell mask;
// (c) the leaf children of the accepted nodes, one per lane: per cell a conservative reach test against the cell's tight
//     box, then (accepting cells only) the prefilter -> (leaf cell, mask of the cells it is a candidate of, flag: within
//     reach of a cell that does not take it) appended to the warp's strip in walk order.
// The warps of a CTA take consecutive blocks (same super cell: its list is shared through L1); no block barriers.
constexpr int WK_WARPS = 4;
constexpr int WK_NT = WK_WARPS * 32;
#ifndef FSPH_WK_Q
#define FSPH_WK_Q 256
#endif
#ifndef FSPH_WK_OCC
#define FSPH_WK_OCC 8
#endif
constexpr int WK_Q = FSPH_WK_Q;  // queued nodes per round of (b) + (c)
constexpr int WK_ITEMS = 32 * 8; // leaf children of 32 queue entries
template<int D> struct WalkSmem {
  WcInfo<D> cell[BLK_CELLS];
  float fbox[BLK_CELLS][8];   // per cell: tight box relative to the block centre (rounded outwards), largest h (rounded up),
                              // largest |coordinate| of the box
  int32_t q_node[WK_Q];
  int32_t it_cell[WK_ITEMS];
  uint16_t it_mask[WK_ITEMS];
};

template<int D>
__global__ void __launch_bounds__(WK_NT, FSPH_WK_OCC) k_block_walk(Tree t, Cols c, Search sr, const WcInfo<D> * __restrict__ wci, const LeafRec * __restrict__ reps, int64_t * __restrict__ scratch,
  int scratch_cap, SearchCursors * __restrict__ cur, unsigned long long cand_cap, unsigned long long dir_cap, int64_t * __restrict__ wcand,
  uint32_t * __restrict__ status) {
  __shared__ WalkSmem<D> sm_all[WK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  WalkSmem<D> & sm = sm_all[warp];
  int64_t * const strip = scratch + (size_t)(blockIdx.x * WK_WARPS + warp) * scratch_cap;
  for(int blk = blockIdx.x * WK_WARPS + warp; blk < sr.nblk; blk += gridDim.x * WK_WARPS) {
    const int g0 = sr.blk_wc[blk], K = sr.blk_wc[blk + 1] - g0;
    __syncwarp();
    {
      const uint4 * src = reinterpret_cast<const uint4 *>(wci + g0);
      uint4 * dst = reinterpret_cast<uint4 *>(sm.cell);
      const int n16 = K * (int)(sizeof(WcInfo<D>) / 16);
      for(int i = lane; i < n16; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    // the block's own tight box and largest h
    double bmn[D], bmx[D], bh = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) { bmn[d] = 1.79769313486231570e308; bmx[d] = -1.79769313486231570e308; }
    if(lane < K) {
#pragma unroll
      for(int d = 0; d < D; ++d) { bmn[d] = sm.cell[lane].tmn[d]; bmx[d] = sm.cell[lane].tmx[d]; }
      bh = sm.cell[lane].hmax;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for(int d = 0; d < D; ++d) {
        bmn[d] = fmin(bmn[d], __shfl_xor_sync(0xffffffffu, bmn[d], o));
        bmx[d] = fmax(bmx[d], __shfl_xor_sync(0xffffffffu, bmx[d], o));
      }
      bh = fmax(bh, __shfl_xor_sync(0xffffffffu, bh, o));
    }
    // the block box in fp32 (absolute coordinates, rounded outwards) for the list filter
    float fbmn[D], fbmx[D], fbabs = 0.f;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      fbmn[d] = __double2float_rd(bmn[d]);
      fbmx[d] = __double2float_ru(bmx[d]);
      fbabs = fmaxf(fbabs, fmaxf(fabsf(fbmn[d]), fabsf(fbmx[d])));
    }
    const float fbh = __double2float_ru(bh);
    double bctr[D];
#pragma unroll
    for(int d = 0; d < D; ++d) bctr[d] = 0.5 * (bmn[d] + bmx[d]);
    if(lane < K) {
#pragma unroll
      for(int d = 0; d < D; ++d) {
        sm.fbox[lane][d] = __double2float_rd(sm.cell[lane].tmn[d] - bctr[d]);
        sm.fbox[lane][3 + d] = __double2float_ru(sm.cell[lane].tmx[d] - bctr[d]);
      }
      sm.fbox[lane][6] = __double2float_ru(sm.cell[lane].hmax);
      float bext = 0.f;
#pragma unroll
      for(int d = 0; d < D; ++d) bext = fmaxf(bext, fmaxf(fabsf(sm.fbox[lane][d]), fabsf(sm.fbox[lane][3 + d])));
      sm.fbox[lane][7] = bext;
    }
    __syncwarp();
    const int body0 = sm.cell[0].first;
    const int PB = sm.cell[K - 1].first + sm.cell[K - 1].P - body0;
    const int sidx = sr.wc_super[g0];
    const uint64_t sl_lo = sr.sl_off[sidx];
    const int sl_n = sr.sl_len[sidx];
    uint32_t ncands = 0;
    int nq = 0;
    double hcand = 0.0;   // largest h among the candidates
    for(int lbase = 0; lbase < sl_n || nq > 0; lbase += 32) {
      // (a) conservative filter: node box within the block's largest h of the block's tight box (necessary for
      //     intersects_sphere_box with any body of the block), in fp32 on boxes rounded outwards; queued in list order
      if(lbase < sl_n) {
        const int e = lbase + lane;
        int nd = -1;
        bool ok = false;
        if(e < sl_n) {
          const float4 * rp = reinterpret_cast<const float4 *>(sr.sl_rec + sl_lo + e);
          const float4 ra = rp[0], rb = rp[1];   // mn.xyz mx.x | mx.yz node pad
          nd = __float_as_int(rb.z);
          const float nmn[3] = {ra.x, ra.y, ra.z}, nmx[3] = {ra.w, rb.x, rb.y};
          float g2 = 0.f, nabs = 0.f;
#pragma unroll
          for(int d = 0; d < D; ++d) {
            const float tt = fmaxf(0.f, fmaxf(nmn[d] - fbmx[d], fbmn[d] - nmx[d]));
            g2 = fmaf(tt, tt, g2);
            nabs = fmaxf(nabs, fmaxf(fabsf(nmn[d]), fabsf(nmx[d])));
          }
          const float hm = fbh + 1e-6f * (fbabs + nabs);
          ok = g2 <= hm * hm * 1.00001f;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if(ok) sm.q_node[nq + __popc(bal & lt)] = nd;
        nq += __popc(bal);
        if(nq <= WK_Q - 32 && lbase + 32 < sl_n) continue;
      }
      __syncwarp();
      for(int eb = 0; eb < nq; eb += 32) {
        // (b) the tests of 32 queued nodes against every cell of the block: intersects_box_box + any-body
        //     intersects_sphere_box.  First one node per lane: in fp32, relative to the block centre and with the node box
        //     rounded outwards, the cells farther from it than their largest h are dropped (the exact quick reject below
        //     would drop them as well); then the surviving (node, cell) pairs are dealt to the lanes for the exact tests.
        const int e = eb + lane;
        uint32_t pm = 0, lmask = 0;
        int cb = 0;
        if(e < nq) {
          const NodeRec & rec = t.noderec[sm.q_node[e]];
          cb = rec.child_begin;
          lmask = rec.leafmask;
          float fmn[D], fmx[D], nabs = 0.f;
#pragma unroll
          for(int d = 0; d < D; ++d) {
            fmn[d] = __double2float_rd(rec.bmin[d] - bctr[d]);
            fmx[d] = __double2float_ru(rec.bmax[d] - bctr[d]);
            nabs = fmaxf(nabs, fmaxf(fabsf(fmn[d]), fabsf(fmx[d])));
          }
          for(int ci = 0; ci < K; ++ci) {
            const float * fb = sm.fbox[ci];
            float g2 = 0.f;
#pragma unroll
            for(int d = 0; d < D; ++d) {
              const float tt = fmaxf(0.f, fmaxf(fmn[d] - fb[3 + d], fb[d] - fmx[d]));
              g2 = fmaf(tt, tt, g2);
            }
            const float hm = fb[6] + 1e-6f * (fb[7] + nabs);
            if(g2 <= hm * hm * 1.00001f) pm |= 1u << ci;
          }
        }
        uint16_t * const pair_sm = reinterpret_cast<uint16_t *>(sm.it_cell);   // (lane of the node) << 8 | cell
        uint32_t * const mask_sm = reinterpret_cast<uint32_t *>(sm.it_mask);   // per node: accepting cells
        int npair;
        {
          uint32_t x = __popc(pm);
#pragma unroll
          for(int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if(lane >= o) x += y;
          }
          npair = (int)__shfl_sync(0xffffffffu, x, 31);
          uint32_t off = x - __popc(pm);
          for(uint32_t rest = pm; rest; rest &= rest - 1) pair_sm[off++] = (uint16_t)((lane << 8) | (__ffs(rest) - 1));
          mask_sm[lane] = 0u;
        }
        __syncwarp();
        for(int pi = lane; pi < npair; pi += 32) {
          const int slot = pair_sm[pi] >> 8, ci = pair_sm[pi] & 255;
          const NodeRec & rec = t.noderec[sm.q_node[eb + slot]];
          double nmn[D], nmx[D];
#pragma unroll
          for(int d = 0; d < D; ++d) { nmn[d] = rec.bmin[d]; nmx[d] = rec.bmax[d]; }
          const WcInfo<D> & G = sm.cell[ci];
          bool ok = true;
          if(G.has_node) {
            bool sep = false;
#pragma unroll
            for(int d = 0; d < D; ++d) sep = sep || (nmx[d] < G.gmn[d]) || (G.gmx[d] < nmn[d]);
            ok = !sep;
          }
          // quick reject (conservative, exact outcome): the node box is farther than hmax from every position of G
          if(ok) ok = box_gap2<D>(nmn, nmx, G.tmn, G.tmx) <= G.hmax2;
          if(ok) {
            int q = 0;
#pragma unroll
            for(int d = 0; d < D; ++d) q |= (int)((nmn[d] + nmx[d]) > (G.tmn[d] + G.tmx[d])) << d;
            const LeafRec rr = reps[(size_t)(g0 + ci) * (1 << D) + q];
            double p[D];
#pragma unroll
            for(int d = 0; d < D; ++d) p[d] = rr.x[d];
            bool acc = sphere_box<D>(nmn, nmx, p, rr.h);
            for(int k = G.first; k < G.first + G.P && !acc; ++k) {
#pragma unroll
              for(int d = 0; d < D; ++d) p[d] = c.f[F_POS + d][k];
              acc = sphere_box<D>(nmn, nmx, p, c.f[F_H][k]);
            }
            ok = acc;
          }
          if(ok) atomicOr(&mask_sm[slot], 1u << ci);
        }
        __syncwarp();
        const uint32_t mask = mask_sm[lane];
        if(!mask) lmask = 0;
        __syncwarp();
        // (c) the leaf children of the 32 entries, flattened over the lanes
        uint32_t x = __popc(lmask);
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
          if(lane >= o) x += y;
        }
        const int nitem = (int)__shfl_sync(0xffffffffu, x, 31);
        uint32_t off = x - __popc(lmask);
        for(uint32_t rest = lmask; rest; rest &= rest - 1) {
          sm.it_cell[off] = cb + (__ffs(rest) - 1);
          sm.it_mask[off] = (uint16_t)mask;
          ++off;
        }
        __syncwarp();
        for(int ib = 0; ib < nitem; ib += 32) {
          const int it = ib + lane;
          uint32_t m = 0;
          int cidx = 0;
          bool excl = false;
          if(it < nitem) {
            cidx = sm.it_cell[it];
            const LeafRec lr = t.leafrec[cidx];
            const uint32_t nmask = sm.it_mask[it];
            // reach test in fp32 on coordinates relative to the block centre.  It only has to be conservative: the cell's
            // box is rounded outwards, h upwards, and `slack` covers the roundings of the leaf's coordinates and of the
            // differences (each below 2^-23 of the larger operand)
            float lx[D], lmax = 0.f;
#pragma unroll
            for(int d = 0; d < D; ++d) {
              lx[d] = (float)(lr.x[d] - bctr[d]);
              lmax = fmaxf(lmax, fabsf(lx[d]));
            }
            const float lh = __double2float_ru(lr.h);
            for(int ci = 0; ci < K; ++ci) {
              const float * fb = sm.fbox[ci];
              // reach (conservative): some body of G may lie within max(h_e, hmax_G) of the leaf
              float dd = 0.f;
#pragma unroll
              for(int d = 0; d < D; ++d) {
                const float tt = fmaxf(0.f, fmaxf(fb[d] - lx[d], lx[d] - fb[3 + d]));
                dd = fmaf(tt, tt, dd);
              }
              const float hm = fmaxf(lh, fb[6]) + 1e-6f * (lmax + fb[7]);
              if(!(dd <= hm * hm * 1.00001f)) continue;
              const WcInfo<D> & G = sm.cell[ci];
              bool incl = (nmask >> ci) & 1u;
              if(incl && G.has_node) {
                // |x_e - X_G|^2 <= (max(h_e, lap_G) + radius_G)^2, tree_topology.h:365-370
                const double ext = dmax(lr.h, G.glap) + G.grad;
                double s2 = 0.0;
#pragma unroll
                for(int d = 0; d < D; ++d) {
                  const double df = lr.x[d] - G.gcom[d];
                  s2 = d == 0 ? df * df : s2 + df * df;
                }
                incl = s2 <= ext * ext;
              }
              // within reach of G but not one of G's candidates in the reference: the pair kernel has to consult the mask
              if(incl) m |= 1u << ci;
              else excl = true;
            }
            if(m) hcand = dmax(hcand, lr.h);
            if(m && excl) m |= 1u << 16;
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, m != 0u);
          const uint32_t pos = ncands + __popc(bal & lt);
          if(m && pos < (uint32_t)scratch_cap) strip[pos] = ((int64_t)m << 32) | (uint32_t)cidx;
          ncands += __popc(bal);
        }
        __syncwarp();
      }
      nq = 0;
    }
    // ---- the block's share of the candidate array and of the tile directory -------------------------------------
    const int ntile = nl_ntiles((int)ncands);
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) hcand = fmax(hcand, __shfl_xor_sync(0xffffffffu, hcand, o));
    unsigned long long co = 0, dir0 = 0;
    if(lane == 0) {
      co = atomicAdd(&cur->cand, (unsigned long long)ncands);
      dir0 = atomicAdd(&cur->dir, (unsigned long long)ntile);
    }
    co = __shfl_sync(0xffffffffu, co, 0);
    dir0 = __shfl_sync(0xffffffffu, dir0, 0);
    const bool strip_over = ncands > (uint32_t)scratch_cap;
    const bool overflow = strip_over || co + ncands > cand_cap || dir0 + ntile > dir_cap;
    if(lane == 0) {
      if(strip_over) atomicOr(status, (uint32_t)ERR_QUEUE_OVERFLOW);
      else if(overflow) atomicOr(status, (uint32_t)ERR_SEARCH_CAPACITY);
      sr.blk_cand_off[blk] = co;
      sr.blk_dir_off[blk] = dir0;
      sr.blk_ncand[blk] = overflow ? -1 : (int32_t)ncands;
      sr.blk_body[blk] = body0;
      sr.blk_nbody[blk] = PB;
      sr.blk_hmax[blk] = __double2float_ru(hcand);
    }
    __syncwarp();
    if(!overflow)
      for(uint32_t i = lane; i < ncands; i += 32) wcand[co + i] = strip[i];
  }
}

// ---- the pair kernel -----------------------------------------------------------------------------------------------
// One CTA of 256 threads per block, one body per thread.  The block's candidates (walk order, from k_block_walk) are
// processed in chunks of NL_CHUNK (almost always one).  A chunk is staged in shared memory as 16-byte records -- fp32
// coordinates relative to the centre of the block's tight box, fp32 h^2 -- SORTED by the cell of a small uniform grid
// laid over the block's neighbourhood (cell size ~ half the largest h; rows along x are contiguous in the sorted
// order) and dealt to the chunk's tiles round-robin by sorted position, so that every tile is an even sample of the
// whole neighbourhood and a row of the grid is a contiguous run of slots in every tile.  After that the warps are
// independent (no block barrier inside the tile loop): every lane walks, tile by tile, the rows of the grid that its
// own sphere of reach touches -- about (2 h + cell)^3 candidates instead of the union over a group of bodies, without any
// compaction pass -- appends the tile slots of its neighbours to its list in the warp's staging area, and the warp
// reserves its rows of the tile with one atomicAdd and copies the staged lists out.
//
// The pair predicate within_distance2(x_k, x_e, max(h_k, h_e)) (tree_geometry.h:265-275) is decided in fp32 with a
// rigorous error bound, and re-evaluated exactly in fp64 only where fp32 cannot decide:
//   u = 2^-24; M bounds |coordinate - centre| of the block's bodies and candidates (half extent of the tight box +
//   Hm: a candidate lies within max(h_e, hmax_G) of a cell's tight box), Hm = the largest h among them.  Every fp32
//   coordinate difference d32 is within e1 = 4.01 u M of the true difference (two conversions + one subtraction), so
//   for any pair with sqrt(s) <= Rm = 1.001 Hm + 2 e1 the fp32 sum s32 of the three squares (mul, fma, fma: 3 roundings
//   of positive terms) satisfies |s32 - s| <= e1 (2 sqrt(3) Rm + 3 e1) + 3.1 u Rm^2 < EB - 4.9 u Rm^2,
//   EB = e1 (2 sqrt(3) Rm + 3 e1) + 8 u Rm^2.  With t32 = max(fl32(h_k^2), fl32(h_e^2)):
//     s32 - t32 < -EB  =>  s <= max(h_k^2, h_e^2)   (accepted; the remaining 1.9 u Rm^2 absorb the roundings of t32,
//     s32 - t32 >  EB  =>  s >  max(h_k^2, h_e^2)   (rejected)   of the difference and of the fp64 evaluation itself)
//   and the band in between (relative width ~ 30 u M / Hm: one pair in ~10^5) goes to the exact fp64 test, evaluated
//   operation by operation as the reference does.
// A candidate that lies within reach of a work cell of the block without being one of that cell's candidates in the
// reference (possible with unequal h: the reference's node tests are not conservative) carries a flag -- the sign of
// its h^2 -- and only for those the cell mask is consulted.
struct __align__(16) CandRec {
  float x[3], h2;
};
constexpr int SR_CAP = NL_CHUNK;              // candidate records in shared memory (+ the sentinel)
#ifndef FSPH_SR_UNROLL
#define FSPH_SR_UNROLL 8
#endif
#ifndef FSPH_SR_NCELL
#define FSPH_SR_NCELL 2048
#endif
constexpr int SR_NCELL = FSPH_SR_NCELL;        // cells of the grid (at most)
constexpr int SR_CST = SR_NCELL + 16;         // entries of a tile's cell table
constexpr int SM_CM = (SR_CAP + 1) * 16;
constexpr int SM_CS = SM_CM + (SR_CAP + 8) * 2;
constexpr int SM_LBUF = SM_CS + NL_CHUNK_TILES * SR_CST * 2;
constexpr int SM_TOTAL = SM_LBUF + SR_NW * (SR_LROWS + 1) * 64;
// during the sort the list staging area holds the cell counters / cursors and the sorted walk positions
constexpr int SM_CNT = SM_LBUF;
constexpr int SM_WIDX = SM_CNT + SR_CST * 4;
constexpr int SM_CCELL = SM_WIDX + SR_CAP * 2;
static_assert(SM_CCELL + SR_CAP * 2 <= SM_TOTAL, "sort scratch does not fit the list staging area");
constexpr int block_pairs_smem() { return SM_TOTAL; }

// the grid of a chunk and the division by the chunk's tile count
struct PairGrid {
  float org[3], invc;     // origin (relative to the block centre), 1 / cell size
  int nx, ny, nz;
};

template<int D>
__device__ __forceinline__ int grid_cell(const PairGrid & g, const float * x) {
  int cidx = 0;
#pragma unroll
  for(int d = D - 1; d >= 0; --d) {
    const int n = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
    int i = (int)floorf((x[d] - g.org[d]) * g.invc);
    i = min(max(i, 0), n - 1);
    cidx = cidx * n + i;
  }
  return cidx;
}

// The pair tests of one lane on a tile of the staged chunk: the rows (iy, iz) of the grid within reach, and in each row
// the slots of the tile that belong to the cells [cx0, cx1] (cst: first slot of each cell in this tile).  All lanes of
// the warp run the loop together, each on its own rows: a lane that has used up its row takes the next one in the same
// iteration (predicated), a lane whose row is empty sits the iteration out on the sentinel record.  Neighbours' slots
// go to row 1 + j of the lane's column of the warp's staging area (rows beyond SR_LROWS are counted, not stored).
// Optimistic: a pair is taken when the fp32 predicate does not exclude it; the return value tells whether any of the
// lane's decisions needs the careful version below (a pair inside the error band, or a flagged candidate accepted).
// shared-memory accesses of the hot loop through 32-bit shared-window addresses (the generic-pointer forms make the
// compiler rebuild the window base inside the loop)
__device__ __forceinline__ float4 lds_f4(const uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(const uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u16(const uint32_t a, const uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v));
}

// shared -> global copy of `bytes` (a multiple of 16, both sides 16-byte aligned) by the TMA engine; returns once the
// source may be overwritten.  The fence makes the generic-proxy writes of the staged rows visible to the async proxy.
__device__ __forceinline__ void bulk_store_rows(void * dst, const void * src_shared, const uint32_t bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(src_shared);
  asm volatile("fence.proxy.async.shared::cta;\n\t"
               "cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t"
               "cp.async.bulk.commit_group;\n\t"
               "cp.async.bulk.wait_group.read 0;" :: "l"(dst), "r"(sa), "r"(bytes) : "memory");
}

template<int D>
__device__ __forceinline__ bool lane_pair_tests_fast(const CandRec * __restrict__ c4, const int sentinel, const uint16_t * __restrict__ cst,
  const PairGrid & g, const int cx0, const int cx1, const int cy0, const int cy1, const int cz0, const int cz1, const float * af, const float h0f,
  const float EB, const bool active, uint16_t * const lcol, int & cnt) {
  const uint32_t rec0 = (uint32_t)__cvta_generic_to_shared(c4);
  const uint32_t sent = rec0 + (uint32_t)sentinel * 16u;
  const uint32_t hioff = (uint32_t)(cx1 + 1 - cx0) * 2u;                // from a row's first cell of the lane to one past its last
  const uint32_t rowstep = (uint32_t)g.nx * 2u;
  const uint32_t wrapstep = (uint32_t)((g.ny - (cy1 - cy0)) * g.nx) * 2u;   // from the last row of a z-slab to the first of the next
  uint32_t row = (uint32_t)__cvta_generic_to_shared(cst) + (uint32_t)((cz0 * g.ny + cy0 - 1) * g.nx + cx0) * 2u;   // the row before the first
  int yleft = 0, zleft = cz1 - cz0 + 1;       // rows left in the current z-slab, z-slabs left (including the current one)
  const int ny_lane = cy1 - cy0 + 1;
  uint32_t cur = 0, end = 0;
  bool done = !active;
  const uint32_t l0 = (uint32_t)__cvta_generic_to_shared(lcol), lmax = l0 + SR_LROWS * 64u;
  uint32_t lp = l0;
  float amin = 3.0e38f;
  int fb = 0;
  // before the first row: yleft == 0 makes the first advance open the first z-slab
  row -= wrapstep - rowstep;
  // a lane that has run out of rows parks on an empty row (both ends read from the table's first entry)
  const uint32_t park = (uint32_t)__cvta_generic_to_shared(cst);
  if(done) zleft = 0;
  for(;;) {
#pragma unroll
    for(int rep = 0; rep < FSPH_SR_UNROLL; ++rep) {
      if(cur >= end) {
        const bool wrap = yleft == 0;
        row += wrap ? wrapstep : rowstep;
        yleft = wrap ? ny_lane - 1 : yleft - 1;
        zleft -= wrap ? 1 : 0;
        const bool out = zleft < 0;
        row = out ? park : row;
        cur = lds_u16(row);
        end = lds_u16(row + (out ? 0u : hioff));
      }
      const bool valid = cur < end;
      const float4 r = lds_f4(valid ? rec0 + cur * 16u : sent);
      const float dx = af[0] - r.x;
      float s = dx * dx;
      if(D > 1) { const float dy = af[1] - r.y; s = fmaf(dy, dy, s); }
      if(D > 2) { const float dz = af[2] - r.z; s = fmaf(dz, dz, s); }
      const float dd = s - fmaxf(h0f, fabsf(r.w));
      amin = fminf(amin, fabsf(dd));
      const bool in = dd <= EB;
      if(in) fb |= __float_as_int(r.w);
      if(in && lp < lmax) sts_u16(lp, cur);
      lp += in ? 64u : 0u;
      cur += valid ? 1u : 0u;
    }
    if(__all_sync(0xffffffffu, zleft < 0)) break;
  }
  cnt = (int)((lp - l0) >> 6);
  return amin <= EB || fb < 0;
}

// The careful version: flagged candidates (within reach of a work cell that does not take them in the reference) are
// checked against the body's cell, pairs inside the error band of the fp32 predicate are re-evaluated exactly in fp64.
// Lanes with `active` false only keep the warp together.  DIRECT: the slots go straight to the warp's rows in global
// memory (lists longer than the staging area).  cm: the cell masks, cidx: the body indices of the tile's candidates.
template<int D, bool DIRECT>
__device__ __forceinline__ void lane_pair_tests_slow(const CandRec * __restrict__ c4, const uint16_t * __restrict__ cm, const uint16_t * __restrict__ cst,
  const PairGrid & g, const int cx0, const int cx1, const int cy0, const int cy1, const int cz0, const int cz1, const float * af, const float h0f,
  const float EB, const double * a0, const double h0, const Cols & c, const int32_t * __restrict__ cidx, const uint32_t my_bit, const bool active,
  const int lane, uint16_t (* lrows)[32], uint16_t * __restrict__ dst, int & cnt) {
  int iy = cy0 - 1, iz = cz0;
  int cur = 0, end = 0;
  bool done = !active;
  for(;;) {
    if(cur >= end && !done) {
      if(++iy > cy1) { iy = cy0; ++iz; }
      if(iz > cz1) done = true;
      else {
        const int base = (iz * g.ny + iy) * g.nx;
        cur = cst[base + cx0];
        end = cst[base + cx1 + 1];
      }
    }
    if(__all_sync(0xffffffffu, done)) break;
    if(cur < end) {
      const float4 r = *reinterpret_cast<const float4 *>(&c4[cur]);
      const float dx = af[0] - r.x;
      float s = dx * dx;
      if(D > 1) { const float dy = af[1] - r.y; s = fmaf(dy, dy, s); }
      if(D > 2) { const float dz = af[2] - r.z; s = fmaf(dz, dz, s); }
      const float dd = s - fmaxf(h0f, fabsf(r.w));
      if(dd <= EB) {
        bool in = true;
        if(r.w < 0.f) in = (cm[cur] & my_bit) != 0u;
        if(in && dd >= -EB) {
          // (dx^2 + dy^2) + dz^2 <= max(h_k, h_e)^2 in fp64 (max(h_k,h_e)^2 == max(h_k^2,h_e^2) exactly: rounding is monotonic)
          const int e = cidx[cur];
          double s0 = (a0[0] - c.f[F_POS][e]) * (a0[0] - c.f[F_POS][e]);
          if(D > 1) s0 = s0 + (a0[1] - c.f[F_POS + 1][e]) * (a0[1] - c.f[F_POS + 1][e]);
          if(D > 2) s0 = s0 + (a0[2] - c.f[F_POS + 2][e]) * (a0[2] - c.f[F_POS + 2][e]);
          const double he = c.f[F_H][e];
          in = s0 <= h0 || s0 <= he * he;
        }
        if(in) {
          if(DIRECT) dst[(1 + cnt) * 32 + lane] = (uint16_t)cur;
          else if(cnt < SR_LROWS) lrows[1 + cnt][lane] = (uint16_t)cur;
          ++cnt;
        }
      }
      ++cur;
    }
  }
}

template<int D>
__global__ void __launch_bounds__(SR_NT, FSPH_SR_OCC) k_block_pairs(Tree t, Cols c, Search sr, const WcInfo<D> * __restrict__ wci,
  const int64_t * __restrict__ wcand, SearchCursors * __restrict__ cur, unsigned long long rows_cap, uint32_t * __restrict__ status) {
  constexpr int NT = SR_NT;
  __shared__ int32_t cfirst[BLK_CELLS];
  __shared__ double cbox[BLK_CELLS][2 * D + 1];
  __shared__ uint32_t scan_sm[SR_NW];
  extern __shared__ __align__(16) unsigned char ubuf[];   // block_pairs_smem() bytes
  CandRec * const c4 = reinterpret_cast<CandRec *>(ubuf);                                    // [SR_CAP + 1] x, y, z, h^2
  uint16_t * const cm = reinterpret_cast<uint16_t *>(ubuf + SM_CM);                          // [SR_CAP] cell masks
  uint16_t (* const cst)[SR_CST] = reinterpret_cast<uint16_t (*)[SR_CST]>(ubuf + SM_CS);     // per tile: [cells + 1] first slot of each cell
  uint16_t (* const lbuf)[SR_LROWS + 1][32] = reinterpret_cast<uint16_t (*)[SR_LROWS + 1][32]>(ubuf + SM_LBUF);
  uint32_t * const cnt_sm = reinterpret_cast<uint32_t *>(ubuf + SM_CNT);                     // sort: per cell counter / cursor
  uint16_t * const widx = reinterpret_cast<uint16_t *>(ubuf + SM_WIDX);                      // sort: walk position by sorted position
  uint16_t * const ccell = reinterpret_cast<uint16_t *>(ubuf + SM_CCELL);                    // sort: grid cell by walk position

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if(tid == 0) {
    // the sentinel record lanes without a candidate test: accepted by nobody (s = inf)
    CandRec r;
    r.x[0] = r.x[1] = r.x[2] = 1e30f;
    r.h2 = 1.f;
    c4[SR_CAP] = r;
  }
  for(int blk = blockIdx.x; blk < sr.nblk; blk += gridDim.x) {
    const int ncand_s = sr.blk_ncand[blk];
    if(ncand_s < 0) continue;   // did not fit (the search is repeated)
    const int ncands = ncand_s;
    const int g0 = sr.blk_wc[blk], K = sr.blk_wc[blk + 1] - g0;
    const int body0 = sr.blk_body[blk], PB = sr.blk_nbody[blk];
    const unsigned long long co = sr.blk_cand_off[blk], dir0 = sr.blk_dir_off[blk];
    const int64_t * const strip = wcand + co;
    __syncthreads();   // the previous block is done with shared memory
    if(tid < K) {
      const WcInfo<D> & w = wci[g0 + tid];
      cfirst[tid] = w.first - body0;
#pragma unroll
      for(int d = 0; d < D; ++d) { cbox[tid][d] = w.tmn[d]; cbox[tid][D + d] = w.tmx[d]; }
      cbox[tid][2 * D] = w.hmax;
    }
    const bool active = tid < PB;
    double a0[D], ha = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) a0[d] = 0.0;
    if(active) {
#pragma unroll
      for(int d = 0; d < D; ++d) a0[d] = c.f[F_POS + d][body0 + tid];
      ha = c.f[F_H][body0 + tid];
    }
    __syncthreads();
    // the block's tight box and largest h, and the cell of each body
    double bmn[D], bmx[D], bh = cbox[0][2 * D];
#pragma unroll
    for(int d = 0; d < D; ++d) { bmn[d] = cbox[0][d]; bmx[d] = cbox[0][D + d]; }
    int my_slot = 0;
    for(int k = 1; k < K; ++k) {
#pragma unroll
      for(int d = 0; d < D; ++d) { bmn[d] = fmin(bmn[d], cbox[k][d]); bmx[d] = fmax(bmx[d], cbox[k][D + d]); }
      bh = fmax(bh, cbox[k][2 * D]);
      my_slot += cfirst[k] <= tid;
    }
    double bctr[D], hext = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      bctr[d] = 0.5 * (bmn[d] + bmx[d]);
      hext = dmax(hext, bmx[d] - bctr[d]);
      hext = dmax(hext, bctr[d] - bmn[d]);
    }
    float af[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for(int d = 0; d < D; ++d) af[d] = active ? (float)(a0[d] - bctr[d]) : 0.f;
    const double h0 = ha * ha;
    const float h0f = (float)h0;
    const int nwarp_act = (PB + 31) >> 5;
    const uint32_t my_bit = 1u << my_slot;
    // error bound of the fp32 predicate, reach and grid (header comment)
    const double Hm = dmax(bh, (double)sr.blk_hmax[blk]) * (1.0 + 1e-6);
    float EB, reach;
    PairGrid g;
    {
      const double u = 5.9604644775390625e-8;   // 2^-24
      const double M = hext + 1.0001 * Hm;
      const double e1 = 4.01 * u * M;
      const double Rm = 1.001 * Hm + 2.0 * e1;
      EB = __double2float_ru((e1 * (3.4641016151377553 * Rm + 3.0 * e1) + 8.0 * u * Rm * Rm) * 1.0001);
      const double R = 1.0001 * Hm + 8.0 * e1;   // no pair farther apart than this along any axis
      reach = __double2float_ru(R);
      // cells of half the reach, enlarged until the grid has at most SR_NCELL cells
      double cell = 0.5 * R;
      double ext[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for(int d = 0; d < D; ++d) ext[d] = (bmx[d] - bmn[d]) + 2.0 * R;
      int n[3] = {1, 1, 1};
      for(int pass = 0; pass < 8; ++pass) {
        double tot = 1.0;
#pragma unroll
        for(int d = 0; d < D; ++d) {
          const double q = ceil(ext[d] / cell);
          n[d] = (int)dmin(q, 1024.0);
          if(n[d] < 1) n[d] = 1;
          tot *= q;
        }
        if(tot <= (double)SR_NCELL && n[0] * n[1] * n[2] <= SR_NCELL) break;
        cell *= (D == 1 ? tot / SR_NCELL : (D == 2 ? sqrt(tot / SR_NCELL) : cbrt(tot / SR_NCELL))) * 1.02;
      }
      if(n[0] * n[1] * n[2] > SR_NCELL) n[0] = n[1] = n[2] = 1;   // (never in practice) one cell: every lane tests every candidate
      g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
      g.invc = (float)(1.0 / cell);
#pragma unroll
      for(int d = 0; d < 3; ++d) g.org[d] = d < D ? __double2float_rd((bmn[d] - bctr[d]) - R) : 0.f;
    }
    const int ncell = g.nx * g.ny * g.nz;
    // the lane's cell ranges (the margin of 1e-3 cells covers the roundings of both sides' cell coordinates)
    int cl[3] = {0, 0, 0}, ch[3] = {0, 0, 0};
#pragma unroll
    for(int d = 0; d < D; ++d) {
      const int n = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
      const float gx = (af[d] - g.org[d]) * g.invc, rc = reach * g.invc + 1e-3f;
      cl[d] = min(max((int)floorf(gx - rc), 0), n - 1);
      ch[d] = min(max((int)floorf(gx + rc), 0), n - 1);
    }
    bool overflow = false;
    int nn = 0;
    const int nchunk = (ncands + NL_CHUNK - 1) / NL_CHUNK;
    for(int chk = 0; chk < nchunk; ++chk) {
      const int cbase = chk * NL_CHUNK;
      const int nc = min(NL_CHUNK, ncands - cbase);
      const int nt = (nc + NL_TILE - 1) / NL_TILE;
      if(chk > 0) __syncthreads();   // the previous chunk's records and lists are no longer used
      // ---- sort the chunk by grid cell (ties in walk order) and stage it ----------------------------------------
      for(int i = tid; i <= ncell; i += NT) cnt_sm[i] = 0u;
      __syncthreads();
      for(int i = tid; i < nc; i += NT) {
        const LeafRec lr = t.leafrec[(uint32_t)strip[cbase + i]];
        float x[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for(int d = 0; d < D; ++d) x[d] = (float)(lr.x[d] - bctr[d]);
        const int gc = grid_cell<D>(g, x);
        ccell[i] = (uint16_t)gc;
        atomicAdd(&cnt_sm[gc], 1u);
      }
      __syncthreads();
      {
        // exclusive scan of the cell counters: 8 consecutive cells per thread
        constexpr int PER = SR_NCELL / NT + 1;   // cells 0 .. ncell (the end marker) in runs of PER per thread
        uint32_t v[PER], sum = 0;
#pragma unroll
        for(int k = 0; k < PER; ++k) {
          const int i = tid * PER + k;
          v[k] = i < ncell ? cnt_sm[i] : 0u;
          sum += v[k];
        }
        uint32_t total;
        uint32_t run = block_scan_u32<NT>(sum, total, scan_sm);
#pragma unroll
        for(int k = 0; k < PER; ++k) {
          const int i = tid * PER + k;
          if(i <= ncell) {
            // sorted positions [run, ...) of cell i; those of tile tj are p = tj (mod nt): first slot ceil((run - tj) / nt)
            const uint32_t pos = i < ncell ? run : (uint32_t)nc;
            if(i < ncell) cnt_sm[i] = run;
            for(int tj = 0; tj < nt; ++tj) cst[tj][i] = (uint16_t)((pos + (uint32_t)(nt - 1 - tj)) / (uint32_t)nt);
          }
          run += v[k];
        }
      }
      __syncthreads();
      for(int i = tid; i < nc; i += NT) widx[atomicAdd(&cnt_sm[ccell[i]], 1u)] = (uint16_t)i;
      __syncthreads();
      // the order inside a cell is that of the atomics: put every cell in walk order (cells hold ~1 candidate)
      for(int i = tid; i < ncell; i += NT) {
        const int lo = i > 0 ? (int)cnt_sm[i - 1] : 0, hi = (int)cnt_sm[i];   // the cursors now stand at the cells' ends
        for(int a = lo + 1; a < hi; ++a) {
          const uint16_t key = widx[a];
          int b = a - 1;
          while(b >= lo && widx[b] > key) { widx[b + 1] = widx[b]; --b; }
          widx[b + 1] = key;
        }
      }
      __syncthreads();
      // sorted position p -> tile p % nt, slot p / nt: record index tb(tile) + slot (the chunk's tiles back to back)
      const int tq = nc / nt, tr = nc % nt;
      for(int p = tid; p < nc; p += NT) {
        const int64_t sv = strip[cbase + widx[p]];
        const LeafRec lr = t.leafrec[(uint32_t)sv];
        const int tj = p % nt, slot = p / nt;
        const int j = tj * tq + min(tj, tr) + slot;
        CandRec r;
        r.x[0] = r.x[1] = r.x[2] = 0.f;
#pragma unroll
        for(int d = 0; d < D; ++d) r.x[d] = (float)(lr.x[d] - bctr[d]);
        const float h2 = (float)(lr.h * lr.h);
        r.h2 = ((sv >> 48) & 1) ? -h2 : h2;
        c4[j] = r;
        cm[j] = (uint16_t)(sv >> 32);
        sr.cand[co + cbase + j] = t.cell[(uint32_t)sv];
      }
      __syncthreads();   // the sort scratch becomes the list staging area
      // ---- the warp's tiles of this chunk: no block barrier from here on ----------------------------------------
      if(warp < nwarp_act) {
        for(int tj = 0; tj < nt; ++tj) {
          const int ti = chk * NL_CHUNK_TILES + tj;
          const int tb = tj * tq + min(tj, tr);
          const CandRec * const tc4 = c4 + tb;
          const uint16_t * const tcm = cm + tb;
          const int32_t * const cidx = sr.cand + co + cbase + tb;
          const int sentinel = SR_CAP - tb;
          int cnt = 0;
          const bool redo = lane_pair_tests_fast<D>(tc4, sentinel, cst[tj], g, cl[0], ch[0], cl[1], ch[1], cl[2], ch[2], af, h0f, EB, active,
            &lbuf[warp][1][lane], cnt);
          if(__any_sync(0xffffffffu, redo)) {
            int cnt2 = 0;
            lane_pair_tests_slow<D, false>(tc4, tcm, cst[tj], g, cl[0], ch[0], cl[1], ch[1], cl[2], ch[2], af, h0f, EB, a0, h0, c, cidx, my_bit,
              redo, lane, lbuf[warp], nullptr, cnt2);
            if(redo) cnt = cnt2;
          }
          lbuf[warp][0][lane] = (uint16_t)cnt;
          nn += cnt;
          const int cmax = __reduce_max_sync(0xffffffffu, cnt);
          // the warp's rows of this tile: 1 + cmax (row 0: the lanes' counts)
          unsigned long long rb = 0;
          if(lane == 0) rb = atomicAdd(&cur->rows, (unsigned long long)(1 + cmax));
          rb = __shfl_sync(0xffffffffu, rb, 0);
          if(rb + 1 + cmax > rows_cap) overflow = true;
          __syncwarp();
          if(!overflow) {
            if(lane == 0) { sr.dir[dir0 + ti].row[warp] = (uint32_t)rb; sr.dir[dir0 + ti].cmax[warp] = (uint32_t)cmax; }
            uint16_t * dst = sr.nl + rb * 32;
            if(cmax <= SR_LROWS) {
              // rows past a lane's own count are zeroed (the sweeps read whole rows); then the warp's 1 + cmax staged rows,
              // contiguous in shared and in global memory, leave with one TMA bulk store
              for(int j = cnt + 1; j <= cmax; ++j) lbuf[warp][j][lane] = (uint16_t)0;
              __syncwarp();
              if(lane == 0) bulk_store_rows(dst, &lbuf[warp][0][0], (uint32_t)(1 + cmax) * 64u);
            }
            else {
              // a lane has more neighbours in this tile than the staging area holds (rare): the tests are repeated,
              // writing straight to the rows
              dst[lane] = (uint16_t)cnt;
              for(int j = cnt + 1; j <= cmax; ++j) dst[j * 32 + lane] = (uint16_t)0;
              int k = 0;
              lane_pair_tests_slow<D, true>(tc4, tcm, cst[tj], g, cl[0], ch[0], cl[1], ch[1], cl[2], ch[2], af, h0f, EB, a0, h0, c, cidx, my_bit, active,
                lane, nullptr, dst, k);
            }
          }
          __syncwarp();
        }
      }
      else {
        // warps without bodies own no rows
        for(int tj = lane; tj < nt; tj += 32) { sr.dir[dir0 + chk * NL_CHUNK_TILES + tj].row[warp] = 0u; sr.dir[dir0 + chk * NL_CHUNK_TILES + tj].cmax[warp] = 0u; }
      }
    }
    if(active && !overflow) c.nneigh[body0 + tid] = nn;
    if(overflow && lane == 0) atomicOr(status, (uint32_t)ERR_SEARCH_CAPACITY);
  }
}

template<int D>
static void neighbor_search_t(System & s) {
  Search & sr = s.search;
  if(sr.nwc <= 0) throw std::runtime_error("no work cells");
  uint32_t * status = s.b_status.as<uint32_t>();
  {
    PhaseScope ph(s, "search:super_lists");
    build_super_lists<D>(s);
  }
  const int nblk = sr.nblk;
  // sentinel of the block table
  FSPH_CUDA_OK(cudaMemcpyAsync(sr.blk_wc + nblk, &sr.nwc, 4, cudaMemcpyHostToDevice, s.stream));
  s.b_blk_meta.reserve((size_t)nblk * 36);
  sr.blk_cand_off = s.b_blk_meta.as<uint64_t>();
  sr.blk_dir_off = sr.blk_cand_off + nblk;
  sr.blk_ncand = reinterpret_cast<int32_t *>(sr.blk_dir_off + nblk);
  sr.blk_body = sr.blk_ncand + nblk;
  sr.blk_nbody = sr.blk_body + nblk;
  sr.blk_hmax = reinterpret_cast<float *>(sr.blk_nbody + nblk);
  SearchCursors * cursors = reinterpret_cast<SearchCursors *>(s.b_status.as<uint32_t>() + 8);
  // per work cell: boxes, largest h, corner representatives
  s.b_wcinfo.reserve((size_t)sr.nwc * sizeof(WcInfo<D>));
  s.b_wcrep.reserve((size_t)sr.nwc * (1 << D) * sizeof(LeafRec));
  WcInfo<D> * wci = s.b_wcinfo.as<WcInfo<D>>();
  LeafRec * reps = s.b_wcrep.as<LeafRec>();
  {
    PhaseScope ph(s, "search:cells");
    FSPH_LAUNCH((k_wc_info<D>), (unsigned)((sr.nwc + 15) / 16), 256, s.stream, s.tree, s.cols(), sr, wci, reps);
  }
  const int walk_grid = std::min((nblk + WK_WARPS - 1) / WK_WARPS, 148 * FSPH_WK_OCC);
  const int pair_grid = std::min(nblk, 148 * FSPH_SR_OCC * 2);
  int scratch_cap = s.search_scratch_cap;
  // capacities: from the previous search of this system (scaled when the body count drifts, e.g. with ghosts), else a
  // first guess; a search that does not fit still counts, and is repeated with the exact sizes
  unsigned long long cand_cap, dir_cap, rows_cap;
  if(s.est_cand > 0 && s.est_n > 0) {
    const double r = std::max(1.0, (double)s.n / (double)s.est_n);
    cand_cap = (unsigned long long)(s.est_cand * r * 1.15) + 65536;
    dir_cap = (unsigned long long)(s.est_dir * r * 1.15) + 4096;
    rows_cap = (unsigned long long)(s.est_rows * r * 1.15) + 65536;
  }
  else {
    cand_cap = (unsigned long long)s.n * 8 + 65536;
    dir_cap = (unsigned long long)nblk * 4 + 4096;
    rows_cap = (unsigned long long)s.n * 3 + 65536;
  }
  if(const char * e = getenv("FSPH_SEARCH_TINY_CAP")) {   // test hook: force the repeat-with-exact-sizes path
    if(atoi(e) && !s.search_tiny_done) { cand_cap = 1024; dir_cap = 16; rows_cap = 1024; s.search_tiny_done = true; }
  }
  for(int attempt = 0;; ++attempt) {
    s.b_cand.reserve((size_t)cand_cap * 4);
    s.b_wcand.reserve((size_t)cand_cap * 8);
    s.b_dir.reserve((size_t)dir_cap * sizeof(TileDir));
    s.b_nl.reserve((size_t)rows_cap * 64);
    s.b_search_scratch.reserve((size_t)walk_grid * WK_WARPS * scratch_cap * 8);
    sr.cand = s.b_cand.as<int32_t>();
    sr.dir = s.b_dir.as<TileDir>();
    sr.nl = s.b_nl.as<uint16_t>();
    FSPH_CUDA_OK(cudaMemsetAsync(cursors, 0, sizeof(SearchCursors), s.stream));
    {
      PhaseScope ph(s, "search:walk");
      FSPH_LAUNCH((k_block_walk<D>), walk_grid, WK_NT, s.stream, s.tree, s.cols(), sr, wci, reps, s.b_search_scratch.as<int64_t>(), scratch_cap, cursors,
        cand_cap, dir_cap, s.b_wcand.as<int64_t>(), status);
    }
    {
      PhaseScope ph(s, "search:pairs");
      FSPH_LAUNCH_SMEM((k_block_pairs<D>), pair_grid, SR_NT, block_pairs_smem(), s.stream, s.tree, s.cols(), sr, wci, s.b_wcand.as<int64_t>(), cursors,
        rows_cap, status);
    }
    SearchCursors tot;
    uint32_t st = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&tot, cursors, sizeof(tot), cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    FSPH_CUDA_OK(cudaGetLastError());
    if(!(st & (ERR_SEARCH_CAPACITY | ERR_QUEUE_OVERFLOW))) {
      s.est_cand = tot.cand;
      s.est_dir = tot.dir;
      s.est_rows = tot.rows;
      s.est_n = s.n;
      sr.nrows = tot.rows;
      if(getenv("FSPH_DEBUG_SEARCH")) {
        // structure statistics (diagnostic): block fill, candidates per block, list-row padding
        std::vector<int32_t> nb(nblk), nc(nblk);
        FSPH_CUDA_OK(cudaMemcpy(nb.data(), sr.blk_nbody, (size_t)nblk * 4, cudaMemcpyDeviceToHost));
        FSPH_CUDA_OK(cudaMemcpy(nc.data(), sr.blk_ncand, (size_t)nblk * 4, cudaMemcpyDeviceToHost));
        std::vector<int32_t> sl(sr.nsuper);
        FSPH_CUDA_OK(cudaMemcpy(sl.data(), sr.sl_len, (size_t)sr.nsuper * 4, cudaMemcpyDeviceToHost));
        double sl_sum = 0;
        for(int i = 0; i < sr.nsuper; ++i) sl_sum += sl[i];
        double sb = 0, sc = 0, sw = 0;
        int full = 0;
        for(int b = 0; b < nblk; ++b) { sb += nb[b]; sc += nc[b]; sw += (nb[b] + 31) / 32; full += nb[b] > 224; }
        std::vector<int32_t> nn(s.n);
        FSPH_CUDA_OK(cudaMemcpy(nn.data(), s.col_nneigh.p, (size_t)s.n * 4, cudaMemcpyDeviceToHost));
        double sn = 0;
        for(int64_t i = 0; i < s.n; ++i) sn += nn[i];
        fprintf(stderr, "[search] n %lld nwc %d nsuper %d (list %.0f nodes) nblk %d bodies/block %.1f (>224: %.0f%%) warps/block %.2f cands/block %.0f "
                        "tiles %llu rows %llu pairs %.4g slot efficiency %.3f\n", (long long)s.n, sr.nwc, sr.nsuper, sl_sum / sr.nsuper,
          nblk, sb / nblk, 100.0 * full / nblk, sw / nblk, sc / nblk, (unsigned long long)tot.dir, (unsigned long long)tot.rows, sn,
          sn / ((double)tot.rows * 32));
      }
      return;
    }
    const uint32_t cleared = st & ~(uint32_t)(ERR_SEARCH_CAPACITY | ERR_QUEUE_OVERFLOW);
    FSPH_CUDA_OK(cudaMemcpyAsync(status, &cleared, 4, cudaMemcpyHostToDevice, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(attempt >= 8) throw std::runtime_error("neighbour search: capacity retry did not converge");
    if(st & ERR_QUEUE_OVERFLOW) {
      // a block has more candidates than a walk strip holds (very wide kernels): grow the strips
      if(scratch_cap >= (1 << 20)) throw std::runtime_error("neighbour search: a block has more than 2^20 candidates");
      scratch_cap *= 4;
      s.search_scratch_cap = scratch_cap;
      continue;
    }
    // the cursors counted every candidate and tile; the rows only of the blocks whose candidates fitted
    const bool walk_fitted = tot.cand <= cand_cap && tot.dir <= dir_cap;
    cand_cap = tot.cand + 64;
    dir_cap = tot.dir + 64;
    if(walk_fitted) rows_cap = tot.rows + 64;
  }
}

void neighbor_search(System & s) {
  if(s.dim == 1) neighbor_search_t<1>(s);
  else if(s.dim == 2) neighbor_search_t<2>(s);
  else neighbor_search_t<3>(s);
  s.search_valid = true;
  s.search_epoch++;
}

// ---- CSR export for the parity tests --------------------------------------------------------------------
__global__ void __launch_bounds__(SR_NT) k_export(Cols c, Search sr, const uint64_t * __restrict__ off, int64_t * __restrict__ ids) {
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int PB = sr.blk_nbody[blk];
  if(tid >= PB) return;
  const int C = sr.blk_ncand[blk], ntile = nl_ntiles(C);
  const uint64_t co = sr.blk_cand_off[blk];
  uint64_t o = off[sr.blk_body[blk] + tid];
  for(int ti = 0; ti < ntile; ++ti) {
    const uint16_t * rows = sr.nl + (uint64_t)sr.dir[sr.blk_dir_off[blk] + ti].row[warp] * 32;
    const int cnt = rows[lane];
    int tb0, tc;
    nl_tile_range(C, ti, tb0, tc);
    const uint64_t tb = co + (uint64_t)tb0;
    for(int j = 0; j < cnt; ++j) ids[o++] = c.id[sr.cand[tb + rows[(1 + j) * 32 + lane]]];
  }
}
__global__ void __launch_bounds__(256) k_i32_to_u64(const int32_t * __restrict__ in, uint64_t * __restrict__ out, int64_t n,
  const int32_t * __restrict__ type, int ghosts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = (ghosts && (type[i] & TYPE_GHOST)) ? 0ull : (uint64_t)in[i];
}

void export_neighbors(System & s, std::vector<int64_t> & off, std::vector<int64_t> * ids) {
  const int64_t n = s.n;
  s.b_cnt1.reserve((size_t)(n + 1) * 8);
  uint64_t * o = s.b_cnt1.as<uint64_t>();
  FSPH_LAUNCH(k_i32_to_u64, (unsigned)((n + 255) / 256), 256, s.stream, s.col_nneigh.as<int32_t>(), o, n, s.col_type.as<int32_t>(),
    s.has_ghosts ? 1 : 0);
  exclusive_scan_u64(s, o, o, n, nullptr);
  off.resize(n + 1);
  FSPH_CUDA_OK(cudaMemcpyAsync(off.data(), o, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  // rows of remote copies are dropped: the caller sees the owned bodies in storage order
  std::vector<int32_t> type;
  auto drop_ghost_rows = [&]() {
    if(!s.has_ghosts) return;
    if(type.empty()) {
      type.resize(n);
      FSPH_CUDA_OK(cudaMemcpyAsync(type.data(), s.col_type.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
      FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    }
    std::vector<int64_t> o2;
    o2.reserve(s.n_owned + 1);
    for(int64_t i = 0; i < n; ++i)
      if(!(type[i] & TYPE_GHOST)) o2.push_back(off[i]);
    o2.push_back(off[n]);
    off.swap(o2);
  };
  if(!ids) { drop_ghost_rows(); return; }
  const int64_t tot = off[n];
  ids->resize(tot);
  DevBuf tmp;
  tmp.reserve((size_t)std::max<int64_t>(tot, 1) * 8);
  FSPH_LAUNCH(k_export, s.search.nblk, SR_NT, s.stream, s.cols(), s.search, o, tmp.as<int64_t>());
  FSPH_CUDA_OK(cudaMemcpyAsync(ids->data(), tmp.p, (size_t)tot * 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  tmp.release();
  for(int64_t i = 0; i < n; ++i) std::sort(ids->begin() + off[i], ids->begin() + off[i + 1]);
  drop_ghost_rows();  // ghost rows are empty, so the id array is unchanged
}

}  // namespace fsph
