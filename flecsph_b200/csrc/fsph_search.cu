// fsph_search.cu -- smoothing-length neighbour search (compiled with -fmad=false: bit-exact predicates).
//
// Replaces tree_topology::traversal_sph (include/tree_topology/tree_topology.h:212-442) with the predicates
// of tree_geometry (include/tree_topology/tree_geometry.h: within_distance2 :60,161,265; intersects_box_box
// :118,198,306; intersects_sphere_box :107,219,329).  One CTA per work cell G (<= 128 particles, the
// reference's own partition, so that its non-conservative culls for unequal h are reproduced exactly):
//   1. breadth-first walk from the root; a node survives iff its box overlaps G's box and some particle
//      sphere of G touches it; the leaf children of surviving nodes that pass the |x_e - X_G| prefilter
//      become G's candidates (written to HBM in walk order: runs of consecutive Morton indices);
//   2. candidates are staged in shared memory in tiles of 512 (coordinates + h^2) and every particle of G
//      tests them with the exact fp64 pair predicate (dx^2+dy^2)+dz^2 <= max(h_a,h_b)^2; the outcome is
//      one bit per (particle, candidate) written as 32-bit words [word][particle] -- the neighbour
//      structure all sweeps of the step reuse (positions and h do not change between them).
// Pass 1 runs twice (count, then fill) so that the candidate and mask arrays are sized exactly and laid out
// deterministically by a prefix sum; it is a few per cent of the pair-test cost.
#include <stdexcept>
#include <algorithm>
#include <cstdlib>
#include "fsph_internal.h"

namespace fsph {

// 0: every work cell walks the tree from the root.
// 1: work cells scan a flat node list collected once per super cell (k_super_walk).  Measured on B200 (64 M Sedov): the
//    walk part of the search is 58 + 11 ms against 53 ms for mode 0 (four times as many node tests outweigh the saved
//    levels).
// 2: work cells walk from a frontier seeded by their super cell's list (the top levels of the walk are done once per
//    super cell of <= 4096 bodies): search 86.3 ms against 92.1 ms for mode 0 (lists 2.4 ms).  Default.
// All three give identical neighbour sets (the GPU test-suite passes with each).
#ifndef FSPH_SEARCH_TWO_LEVEL
#define FSPH_SEARCH_TWO_LEVEL 2
#endif
constexpr int SR_NT = 64;       // threads per CTA of the export kernel
// 1: decide the pair predicate in fp32 (cell-centred coordinates, rigorous error margin, exact fp64 evaluation when the
// fp32 result is inside the margin).  Same neighbour sets (tests), but SLOWER on B200 (search 97.8 vs 86.5 ms at 64 M):
// the pair tests are issue-bound, not fp64-bound, and the filter adds instructions and shared memory.  Default 0.
#ifndef FSPH_SR_FP32_FILTER
#define FSPH_SR_FP32_FILTER 0
#endif
constexpr int SR_QCAP_GLOBAL = 8190;    // frontier capacity per level of the global-memory fall-back (x 8 children fits 16 bits)
#ifndef FSPH_SR_CHUNK
#define FSPH_SR_CHUNK 2
#endif
#ifndef FSPH_SR_SPREAD
#define FSPH_SR_SPREAD 64
#endif
constexpr uint32_t SR_SPREAD = FSPH_SR_SPREAD;
constexpr int SW_QCAP = 4096;   // node frontier capacity per level of the super-cell walk
#ifndef FSPH_SR_QCAP
#define FSPH_SR_QCAP 2048
#endif
constexpr int SR_QCAP = FSPH_SR_QCAP;   // node frontier capacity per level (wide kernels: quintic at eta 1.2 reaches ~1000)
constexpr int SR_TILE = 256;    // candidates per shared-memory tile
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

struct WorkCell {
  int start, P, gnode;
};

__device__ __forceinline__ WorkCell load_work_cell(const Tree & t, const Search & sr, int g) {
  WorkCell w;
  int v = t.cell[sr.wc_cell[g]];
  if(v >= 0) { w.start = v; w.P = 1; w.gnode = -1; }
  else { w.gnode = ~v; w.start = t.node_start[w.gnode]; w.P = t.sub[w.gnode]; }
  return w;
}

// MODE 0: count candidates only (exact sizes for the two-pass path)
// MODE 1: fill with offsets from a prefix sum over the counts of MODE 0 (deterministic layout)
// MODE 2: single pass: persistent CTAs, candidates go to a per-CTA scratch strip during the walk, then the
//         CTA reserves its exact share of the candidate / mask arrays with one atomicAdd each (capacity is
//         estimated from the previous step; on overflow the host falls back to MODE 0 + 1)
// resident-thread target per SM for the register allocator (threads = FSPH_SR_THREADS_PER_SM / CTA size CTAs)
#ifndef FSPH_SR_THREADS_PER_SM
#define FSPH_SR_THREADS_PER_SM 448
#endif
// GQ: the node frontier lives in a per-CTA strip of global memory (SR_QCAP_GLOBAL entries per level) instead of shared
// memory: the fall-back for cells whose frontier outgrows SR_QCAP (very wide kernels / strongly varying h).
template<int D, int NW, int MODE, bool GQ = false>
__global__ void __launch_bounds__(32 * NW, FSPH_SR_THREADS_PER_SM / (32 * NW)) k_search(Tree t, Cols c, Search sr, int g_begin, int g_end, int32_t * __restrict__ scratch,
  int scratch_cap, unsigned long long * __restrict__ cursors, unsigned long long cand_cap, unsigned long long mask_cap,
  uint32_t * __restrict__ status, int32_t * __restrict__ gqueue = nullptr) {
  constexpr int NT = 32 * NW;
  constexpr bool FILL = MODE != 0;
  __shared__ double gx[D][32 * NW];
  __shared__ double gh[32 * NW];
  constexpr int QC = GQ ? SR_QCAP_GLOBAL : SR_QCAP;
  __shared__ int32_t queue_sm[GQ ? 1 : 2][GQ ? 1 : SR_QCAP];
  __shared__ uint16_t qres_sm[GQ ? 1 : SR_QCAP];   // per frontier node: node-children mask | candidate-children mask << 8
  // accessors that keep the shared-memory address space visible to the compiler in the common (non-GQ) case
  int32_t * const gqb = GQ ? gqueue + (size_t)blockIdx.x * 3 * QC : nullptr;
  uint16_t * const gqr = GQ ? reinterpret_cast<uint16_t *>(gqueue + (size_t)blockIdx.x * 3 * QC + 2 * QC) : nullptr;
#define SRQ(c, j) (*(GQ ? &gqb[(size_t)(c) * QC + (j)] : &queue_sm[GQ ? 0 : (c)][GQ ? 0 : (j)]))
#define SRR(j) (*(GQ ? &gqr[(j)] : &qres_sm[GQ ? 0 : (j)]))
  __shared__ uint32_t scan_sm[NW];
  __shared__ double red_sm[NW][2 * D + 1];
  __shared__ unsigned long long base_sm[2];
  __shared__ unsigned long long repkey_sm[NW][1 << D];
  __shared__ int rep_sm[1 << D];
  __shared__ double cx[FILL ? D : 1][FILL ? SR_TILE : 1];
  __shared__ double ch2[FILL ? SR_TILE : 1];
#if FSPH_SR_FP32_FILTER
  __shared__ float cxf[FILL ? D : 1][FILL ? SR_TILE : 1];
  __shared__ float ch2f[FILL ? SR_TILE : 1];
  __shared__ uint32_t rmax_sm[2];
#endif

  const int tid = threadIdx.x;
  // Work cells are dealt to the CTAs in chunks of FSPH_SR_CHUNK consecutive cells, round-robin: the CTAs that run at
  // the same time then work on neighbouring regions and share tree nodes / candidates in L2 (a contiguous run per CTA
  // spreads the resident CTAs over the whole volume: 64 GB of DRAM reads per search at 64 M bodies), while the cells
  // of a chunk still follow each other on one SM (L1).
  const int ncell = g_end - g_begin;
  const int nchunk = (ncell + FSPH_SR_CHUNK - 1) / FSPH_SR_CHUNK;
  for(int ck = blockIdx.x; ck < nchunk; ck += gridDim.x)
  for(int g = g_begin + ck * FSPH_SR_CHUNK, g_e = min(g_end, g + FSPH_SR_CHUNK); g < g_e; ++g) {
    const WorkCell wc = load_work_cell(t, sr, g);
    const int P = wc.P, gnode = wc.gnode;
    __syncthreads();
    if(tid < P) {
#pragma unroll
      for(int d = 0; d < D; ++d) gx[d][tid] = c.f[F_POS + d][wc.start + tid];
      gh[tid] = c.f[F_H][wc.start + tid];
    }
    double gmn[D], gmx[D], gcom[D], glap = 0.0, grad = 0.0;
    if(gnode >= 0) {
#pragma unroll
      for(int d = 0; d < D; ++d) { gmn[d] = t.bmin[d][gnode]; gmx[d] = t.bmax[d][gnode]; gcom[d] = t.com[d][gnode]; }
      glap = t.lap[gnode];
      grad = t.radius[gnode];
    }
    int32_t * cout = MODE == 1 ? sr.cand + sr.cand_off[g] : (MODE == 2 ? scratch + (size_t)blockIdx.x * scratch_cap : nullptr);
#if FSPH_SEARCH_TWO_LEVEL == 0
    if(tid == 0) SRQ(0, 0) = 0;  // the root node
#endif
    // Tight box of G's positions and max h: an extra, conservative candidate cull (it only drops leaves that
    // cannot satisfy the pair predicate for any particle of G, so the neighbour sets are unchanged).
    double tmn[D], tmx[D], ghmax = 0.0;
    {
#pragma unroll
      for(int d = 0; d < D; ++d) { tmn[d] = 1.79769313486231570e308; tmx[d] = -1.79769313486231570e308; }
      if(tid < P) {
#pragma unroll
        for(int d = 0; d < D; ++d) { tmn[d] = c.f[F_POS + d][wc.start + tid]; tmx[d] = tmn[d]; }
        ghmax = c.f[F_H][wc.start + tid];
      }
#pragma unroll
      for(int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for(int d = 0; d < D; ++d) {
          tmn[d] = fmin(tmn[d], __shfl_xor_sync(0xffffffffu, tmn[d], o));
          tmx[d] = fmax(tmx[d], __shfl_xor_sync(0xffffffffu, tmx[d], o));
        }
        ghmax = fmax(ghmax, __shfl_xor_sync(0xffffffffu, ghmax, o));
      }
      if((tid & 31) == 0) {
#pragma unroll
        for(int d = 0; d < D; ++d) { red_sm[tid >> 5][d] = tmn[d]; red_sm[tid >> 5][D + d] = tmx[d]; }
        red_sm[tid >> 5][2 * D] = ghmax;
      }
      __syncthreads();
#pragma unroll
      for(int ww = 0; ww < NW; ++ww) {
#pragma unroll
        for(int d = 0; d < D; ++d) {
          tmn[d] = fmin(tmn[d], red_sm[ww][d]);
          tmx[d] = fmax(tmx[d], red_sm[ww][D + d]);
        }
        ghmax = fmax(ghmax, red_sm[ww][2 * D]);
      }
    }
    // Representative particles: for each corner of G's tight box the particle of G nearest to it.  The
    // "does any particle sphere of G touch this node box" loop tries the representative on the node's side
    // first (pure ordering heuristic: the outcome is the exact any-of-all-particles result either way).
    {
      constexpr int NC = 1 << D;
#pragma unroll
      for(int q = 0; q < NC; ++q) {
        unsigned long long key = ~0ull;
        if(tid < P) {
          double d2 = 0.0;
#pragma unroll
          for(int d = 0; d < D; ++d) {
            const double df = gx[d][tid] - (((q >> d) & 1) ? tmx[d] : tmn[d]);
            d2 += df * df;
          }
          key = ((unsigned long long)__double_as_longlong(d2) & ~0x7full) | (unsigned long long)tid;
        }
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
          key = other < key ? other : key;
        }
        if((tid & 31) == 0) repkey_sm[tid >> 5][q] = key;
      }
      __syncthreads();
      if(tid < NC) {
        unsigned long long key = repkey_sm[0][tid];
#pragma unroll
        for(int ww = 1; ww < NW; ++ww) key = repkey_sm[ww][tid] < key ? repkey_sm[ww][tid] : key;
        rep_sm[tid] = (int)(key & 0x7full);
      }
      __syncthreads();
    }
    // hmax^2 with a safety margin for the quick node reject (distance between a node box and G's tight box)
    const double ghmax2 = ghmax * ghmax * (1.0 + 1e-9);
#if FSPH_SEARCH_TWO_LEVEL == 1
    uint32_t ncands = 0;
    bool overflow = false;
    // two-level walk: the nodes that can hold candidates of G were collected once for G's super cell (k_super_walk);
    // a node's ancestors need no test of their own (boxes are nested and both predicates are monotone in the box),
    // so scanning that flat list with the exact tests gives exactly the leaves the reference's top-down walk reaches
    const int sidx = sr.wc_super[g];
    const uint64_t sl_lo = sr.sl_off[sidx];
    const int sl_n = (int)(sr.sl_off[sidx + 1] - sl_lo);
    for(int lbase = 0; lbase < sl_n; lbase += SR_QCAP) {
      const int ncur = min(SR_QCAP, sl_n - lbase);
      // ---- phase A: test every frontier node (results to shared memory), count children / candidates -------
      uint32_t my_cands = 0;
      for(int j = tid; j < ncur; j += NT) {
        const int nd = sr.sl_nodes[sl_lo + lbase + j];
        const NodeRec rec = t.noderec[nd];
        uint32_t cand_mask = 0;
        SRQ(0, j) = rec.child_begin;
        bool ok = true;
        if(gnode >= 0) {
          // intersects_box_box(c, G)
          bool sep = false;
#pragma unroll
          for(int d = 0; d < D; ++d) sep = sep || (rec.bmax[d] < gmn[d]) || (gmx[d] < rec.bmin[d]);
          ok = !sep;
        }
        if(ok) {
          // quick reject (conservative, exact outcome): the node box is farther than hmax from every position of G
          double dd = 0.0;
#pragma unroll
          for(int d = 0; d < D; ++d) {
            const double tt = fmax(0.0, fmax(rec.bmin[d] - tmx[d], tmn[d] - rec.bmax[d]));
            dd = dd + tt * tt;
          }
          ok = dd <= ghmax2;
        }
        if(ok) {
          int q = 0;
#pragma unroll
          for(int d = 0; d < D; ++d) q |= (int)((rec.bmin[d] + rec.bmax[d]) > (tmn[d] + tmx[d])) << d;
          const int kr = rep_sm[q];
          double p[D];
#pragma unroll
          for(int d = 0; d < D; ++d) p[d] = gx[d][kr];
          bool acc = sphere_box<D>(rec.bmin, rec.bmax, p, gh[kr]);
          for(int k = 0; k < P && !acc; ++k) {
#pragma unroll
            for(int d = 0; d < D; ++d) p[d] = gx[d][k];
            acc = sphere_box<D>(rec.bmin, rec.bmax, p, gh[k]);
          }
          ok = acc;
        }
        if(ok) {
          const int cb = rec.child_begin;
          const uint32_t leafmask = rec.leafmask;
          // leaf children in two batches of four: all loads of a batch are issued before the first use
#pragma unroll
          for(int half = 0; half < ((1 << D) + 3) / 4; ++half) {
            const uint32_t lm4 = (leafmask >> (4 * half)) & 0xfu;
            if(lm4 == 0) continue;
            LeafRec lr[4];
#pragma unroll
            for(int jj = 0; jj < 4; ++jj)
              if((lm4 >> jj) & 1u) lr[jj] = t.leafrec[cb + 4 * half + jj];
#pragma unroll
            for(int jj = 0; jj < 4; ++jj) {
              if(!((lm4 >> jj) & 1u)) continue;
              bool pass = true;
              if(gnode >= 0) {
                // |x_e - X_G|^2 <= (max(h_e, lap_G) + radius_G)^2, tree_topology.h:365-370
                const double ext = fmax(lr[jj].h, glap) + grad;
                double s2 = 0.0;
#pragma unroll
                for(int d = 0; d < D; ++d) {
                  const double df = lr[jj].x[d] - gcom[d];
                  s2 = d == 0 ? df * df : s2 + df * df;
                }
                pass = s2 <= ext * ext;
              }
              if(pass) {
                const double hm = fmax(lr[jj].h, ghmax);
                double dd = 0.0;
#pragma unroll
                for(int d = 0; d < D; ++d) {
                  const double tt = fmax(0.0, fmax(tmn[d] - lr[jj].x[d], lr[jj].x[d] - tmx[d]));
                  dd = dd + tt * tt;
                }
                pass = dd <= hm * hm * (1.0 + 1e-9);
              }
              if(pass) cand_mask |= 1u << (4 * half + jj);
            }
          }
        }
        SRR(j) = (uint16_t)cand_mask;
        my_cands += __popc(cand_mask);
      }
      // ---- one block scan per level: output order = (thread, frontier slot, child), fixed => deterministic --
      uint32_t total;
      const uint32_t ex = block_scan_u32<NT>(my_cands, total, scan_sm);
      uint32_t cpos = ncands + ex;
      // ---- phase B: write the next frontier and the candidates -----------------------------------------
      for(int j = tid; j < ncur; j += NT) {
        const uint32_t cand_mask = SRR(j);
        if(FILL && cand_mask) {
          const int child_begin = SRQ(0, j);
          int cv[1 << D];
#pragma unroll
          for(int b = 0; b < (1 << D); ++b)
            if((cand_mask >> b) & 1u) cv[b] = t.cell[child_begin + b];
#pragma unroll
          for(int b = 0; b < (1 << D); ++b) {
            if(!((cand_mask >> b) & 1u)) continue;
            if(MODE == 1 || cpos < (uint32_t)scratch_cap) cout[cpos] = cv[b];
            ++cpos;
          }
        }
      }
      ncands += total;
      __syncthreads();
    }
#else
    int ncur = 1, cur = 0;
    uint32_t ncands = 0;
    bool overflow = false;
#if FSPH_SEARCH_TWO_LEVEL == 2
    // The walk starts from the frontier collected once for G's super cell (k_super_walk, SEED): the nodes of about
    // super-cell size around it, plus bigger nodes that have leaf children (flagged: their node children are in the
    // list themselves).  A node's ancestors need no test of their own (boxes are nested, both predicates are monotone
    // in the box), so the leaves reached are exactly those of the walk from the root -- without its top levels.
    {
      const int sidx = sr.wc_super[g];
      const uint64_t sl_lo = sr.sl_off[sidx];
      ncur = (int)(sr.sl_off[sidx + 1] - sl_lo);
      if(ncur > QC) {
        if(tid == 0) atomicOr(status, (uint32_t)ERR_QUEUE_OVERFLOW);
        ncur = QC;
      }
      for(int j = tid; j < ncur; j += NT) SRQ(0, j) = sr.sl_nodes[sl_lo + j];
      __syncthreads();
    }
#endif
    while(ncur > 0) {
      // ---- phase A: test every frontier node (results to shared memory), count children / candidates -------
      uint32_t my_nodes = 0, my_cands = 0;
      for(int j = tid; j < ncur; j += NT) {
        const int nd_raw = SRQ(cur, j);
        const int nd = nd_raw & 0x3fffffff;
        const NodeRec rec = t.noderec[nd];
        uint32_t node_mask = 0, cand_mask = 0;
        bool ok = true;
        if(gnode >= 0) {
          // intersects_box_box(c, G)
          bool sep = false;
#pragma unroll
          for(int d = 0; d < D; ++d) sep = sep || (rec.bmax[d] < gmn[d]) || (gmx[d] < rec.bmin[d]);
          ok = !sep;
        }
        if(ok) {
          // quick reject (conservative, exact outcome): the node box is farther than hmax from every position of G
          double dd = 0.0;
#pragma unroll
          for(int d = 0; d < D; ++d) {
            const double tt = fmax(0.0, fmax(rec.bmin[d] - tmx[d], tmn[d] - rec.bmax[d]));
            dd = dd + tt * tt;
          }
          ok = dd <= ghmax2;
        }
        if(ok) {
          int q = 0;
#pragma unroll
          for(int d = 0; d < D; ++d) q |= (int)((rec.bmin[d] + rec.bmax[d]) > (tmn[d] + tmx[d])) << d;
          const int kr = rep_sm[q];
          double p[D];
#pragma unroll
          for(int d = 0; d < D; ++d) p[d] = gx[d][kr];
          bool acc = sphere_box<D>(rec.bmin, rec.bmax, p, gh[kr]);
          for(int k = 0; k < P && !acc; ++k) {
#pragma unroll
            for(int d = 0; d < D; ++d) p[d] = gx[d][k];
            acc = sphere_box<D>(rec.bmin, rec.bmax, p, gh[k]);
          }
          ok = acc;
        }
        if(ok) {
          const int cb = rec.child_begin;
          const uint32_t leafmask = rec.leafmask;
          node_mask = (nd_raw & 0x40000000) ? 0u : (((1u << rec.nchild) - 1u) & ~leafmask);
          // leaf children in two batches of four: all loads of a batch are issued before the first use
#pragma unroll
          for(int half = 0; half < ((1 << D) + 3) / 4; ++half) {
            const uint32_t lm4 = (leafmask >> (4 * half)) & 0xfu;
            if(lm4 == 0) continue;
            LeafRec lr[4];
#pragma unroll
            for(int jj = 0; jj < 4; ++jj)
              if((lm4 >> jj) & 1u) lr[jj] = t.leafrec[cb + 4 * half + jj];
#pragma unroll
            for(int jj = 0; jj < 4; ++jj) {
              if(!((lm4 >> jj) & 1u)) continue;
              bool pass = true;
              if(gnode >= 0) {
                // |x_e - X_G|^2 <= (max(h_e, lap_G) + radius_G)^2, tree_topology.h:365-370
                const double ext = fmax(lr[jj].h, glap) + grad;
                double s2 = 0.0;
#pragma unroll
                for(int d = 0; d < D; ++d) {
                  const double df = lr[jj].x[d] - gcom[d];
                  s2 = d == 0 ? df * df : s2 + df * df;
                }
                pass = s2 <= ext * ext;
              }
              if(pass) {
                const double hm = fmax(lr[jj].h, ghmax);
                double dd = 0.0;
#pragma unroll
                for(int d = 0; d < D; ++d) {
                  const double tt = fmax(0.0, fmax(tmn[d] - lr[jj].x[d], lr[jj].x[d] - tmx[d]));
                  dd = dd + tt * tt;
                }
                pass = dd <= hm * hm * (1.0 + 1e-9);
              }
              if(pass) cand_mask |= 1u << (4 * half + jj);
            }
          }
        }
        SRR(j) = (uint16_t)(node_mask | (cand_mask << 8));
        my_nodes += __popc(node_mask);
        my_cands += __popc(cand_mask);
      }
      // ---- one block scan per level: output order = (thread, frontier slot, child), fixed => deterministic --
      uint32_t total;
      const uint32_t ex = block_scan_u32<NT>(my_nodes | (my_cands << 16), total, scan_sm);
      uint32_t qpos = ex & 0xffffu, cpos = ncands + (ex >> 16);
      const uint32_t nnext_raw = total & 0xffffu;
      // ---- phase B: write the next frontier and the candidates -----------------------------------------
      for(int j = tid; j < ncur; j += NT) {
        const uint32_t r = SRR(j);
        if(r == 0) continue;
        const int nd = SRQ(cur, j) & 0x3fffffff;
        const NodeRec rec = t.noderec[nd];
        uint32_t node_mask = r & 0xffu, cand_mask = r >> 8;
        while(node_mask) {
          const int b = __ffs(node_mask) - 1;
          node_mask &= node_mask - 1;
          // node children of a node have consecutive node indices (level order)
          if(qpos < QC) SRQ(cur ^ 1, qpos) = rec.first_child_node + __popc(~rec.leafmask & ((1u << b) - 1u));
          ++qpos;
        }
        if(FILL && cand_mask) {
          int cv[1 << D];
#pragma unroll
          for(int b = 0; b < (1 << D); ++b)
            if((cand_mask >> b) & 1u) cv[b] = t.cell[rec.child_begin + b];
#pragma unroll
          for(int b = 0; b < (1 << D); ++b) {
            if(!((cand_mask >> b) & 1u)) continue;
            if(MODE == 1 || cpos < (uint32_t)scratch_cap) cout[cpos] = cv[b];
            ++cpos;
          }
        }
      }
      ncands += total >> 16;
      uint32_t nnext = nnext_raw;
      if(nnext > QC) {
        if(tid == 0) atomicOr(status, (uint32_t)ERR_QUEUE_OVERFLOW);
        nnext = QC;
      }
      __syncthreads();
      ncur = (int)nnext;
      cur ^= 1;
    }
#endif
    const int Ppad = (P + 31) & ~31;
    const uint64_t nwords = (uint64_t)((ncands + 31) / 32) * Ppad;
    if(MODE == 0) {
      if(tid == 0) {
        sr.ncand[g] = (int32_t)ncands;
        sr.cand_off[g] = ncands;
        sr.mask_off[g] = nwords;
      }
      continue;
    }
    uint64_t co, mo;
    if(MODE == 2) {
      if(ncands > (uint32_t)scratch_cap) overflow = true;
      if(tid == 0) {
        base_sm[0] = atomicAdd(&cursors[0], (unsigned long long)ncands);
        base_sm[1] = atomicAdd(&cursors[1], (unsigned long long)nwords);
      }
      __syncthreads();
      co = base_sm[0];
      mo = base_sm[1];
      if(co + ncands > cand_cap || mo + nwords > mask_cap) overflow = true;
      if(overflow) {
        if(tid == 0) atomicOr(status, (uint32_t)ERR_SEARCH_CAPACITY);
        continue;
      }
      if(tid == 0) {
        sr.ncand[g] = (int32_t)ncands;
        sr.cand_off[g] = co;
        sr.mask_off[g] = mo;
      }
      // The candidates leave the walk in runs of key-neighbours.  They are stored interleaved (walk position i goes
      // to row i % SR_SPREAD of a SR_SPREAD-row transpose): every tile of the sweeps then holds an even sample of the
      // whole neighbourhood, so all particles of the cell find about the same number of neighbours in it and the
      // lanes of the pair loops stay busy together (measured: -12 % on the pair sweeps).
      {
        const uint32_t q0 = ncands / SR_SPREAD, rem = ncands % SR_SPREAD;
        for(uint32_t i = tid; i < ncands; i += NT) {
          const uint32_t r = i % SR_SPREAD, k = i / SR_SPREAD;
          sr.cand[co + r * q0 + min(r, rem) + k] = cout[i];
        }
      }
      __syncthreads();
    }
    else {
      co = sr.cand_off[g];
      mo = sr.mask_off[g];
    }
    // ---- pair tests ------------------------------------------------------------------------------------
    // The predicate is decided in fp32 on coordinates relative to the cell centre whenever the fp32 result is farther
    // from the threshold than a rigorous error bound (|s32 - s| <= 100 eps R^2 for points within R of the centre per
    // axis; 2^-17 R^2 is used), and evaluated exactly in fp64 otherwise: same outcome as the all-fp64 test at a
    // fraction of the fp64-pipe work (the fp64 pipe is the scarce one: it also runs every sweep).
    const int C = (int)ncands;
    const bool active = tid < P;
    double a0[D], h0 = 0.0;
#pragma unroll
    for(int d = 0; d < D; ++d) a0[d] = active ? gx[d][tid] : 0.0;
    if(active) { h0 = gh[tid]; h0 = h0 * h0; }
#if FSPH_SR_FP32_FILTER
    double ctr[D];
    float a0f[D], extf = 0.f;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      ctr[d] = 0.5 * (tmn[d] + tmx[d]);
      a0f[d] = (float)(a0[d] - ctr[d]);
      extf = fmaxf(extf, (float)(tmx[d] - tmn[d]));   // >= twice the largest |x_k - ctr| component
    }
    const float h0f = (float)h0;
    if(tid == 0) { rmax_sm[0] = 0u; rmax_sm[1] = 0u; }
#endif
    int cnt0 = 0;
    for(int tb = 0; tb < C; tb += SR_TILE) {
      const int tc = min(SR_TILE, C - tb);
      __syncthreads();
#if FSPH_SR_FP32_FILTER
      const int par = (tb / SR_TILE) & 1;
      if(tid == 0) rmax_sm[par ^ 1] = 0u;
      float lm = 0.f;
#endif
      for(int i = tid; i < tc; i += NT) {
        const int p = sr.cand[co + tb + i];
#pragma unroll
        for(int d = 0; d < D; ++d) {
          const double x = c.f[F_POS + d][p];
          cx[d][i] = x;
#if FSPH_SR_FP32_FILTER
          const float wf = (float)(x - ctr[d]);
          cxf[d][i] = wf;
          lm = fmaxf(lm, fabsf(wf));
#endif
        }
        const double hh = c.f[F_H][p];
        ch2[i] = hh * hh;
#if FSPH_SR_FP32_FILTER
        ch2f[i] = (float)(hh * hh);
#endif
      }
#if FSPH_SR_FP32_FILTER
      atomicMax(&rmax_sm[par], __float_as_uint(lm));
#endif
      __syncthreads();
      if(!active) continue;
#if FSPH_SR_FP32_FILTER
      const float Rm = fmaxf(__uint_as_float(rmax_sm[par]), extf);
      const float marg = Rm * Rm * (1.0f / 131072.0f);
#endif
      const int nw = (tc + 31) >> 5;
      for(int w = 0; w < nw; ++w) {
        const int nb = min(32, tc - w * 32);
        // within_distance2(x_k, x_e, max(h_k, h_e)): max(h_k,h_e)^2 == max(h_k^2,h_e^2) exactly (rounding is monotonic)
        uint32_t m0 = 0;
#pragma unroll 8
        for(int b = 0; b < nb; ++b) {
          const int ci = w * 32 + b;
#if FSPH_SR_FP32_FILTER
          const float f0 = a0f[0] - cxf[0][ci];
          float sf = f0 * f0;
#pragma unroll
          for(int d = 1; d < D; ++d) {
            const float fd = a0f[d] - cxf[d][ci];
            sf = sf + fd * fd;
          }
          const float thr = fmaxf(h0f, ch2f[ci]);
          const float mt = marg + thr * 2.4e-7f;
          bool in = sf <= thr - mt;
          if(!in && !(sf > thr + mt)) {
#endif
          const double e0 = cx[0][ci];
          double s0 = (a0[0] - e0) * (a0[0] - e0);
#pragma unroll
          for(int d = 1; d < D; ++d) {
            const double ed = cx[d][ci];
            s0 = s0 + (a0[d] - ed) * (a0[d] - ed);
          }
#if FSPH_SR_FP32_FILTER
            in = s0 <= fmax(h0, ch2[ci]);
          }
          m0 |= (uint32_t)in << b;
#else
          m0 |= (uint32_t)(s0 <= fmax(h0, ch2[ci])) << b;
#endif
        }
        sr.mask[mo + (uint64_t)((tb >> 5) + w) * Ppad + tid] = m0;
        cnt0 += __popc(m0);
      }
    }
    if(active) c.nneigh[wc.start + tid] = cnt0;
  }
}
#undef SRQ
#undef SRR


// ---- level 1 of the two-level walk: one node list per super cell -------------------------------------------
// A super cell S (maximal sub-tree of <= SUPER_ENTITIES bodies) walks the tree once with a test that every node any
// work cell G inside S can accept must pass: G accepts C only if some body sphere of G touches C's box
// (intersects_sphere_box, tree_topology.h:348), hence the distance between C's box and the tight box of S's positions
// is at most the largest h in S.  The nodes that pass and have leaf children are S's list (in walk order: fixed).
// MODE 0 counts, MODE 1 fills (offsets from a prefix sum).
template<int D, int MODE>
__global__ void __launch_bounds__(64) k_super_walk(Tree t, Cols c, Search sr, uint32_t * __restrict__ status) {
  constexpr int NT = 64;
  __shared__ int32_t queue[2][SW_QCAP];
  __shared__ uint8_t qres[SW_QCAP];
  __shared__ uint32_t scan_sm[NT / 32];
  __shared__ double red_sm[NT / 32][2 * D + 1];
  const int tid = threadIdx.x, sidx = blockIdx.x;
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
  if(tid == 0) queue[0][0] = 0;  // the root node
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
  int ncur = 1, cur = 0;
  uint32_t nlist = 0;
  while(ncur > 0) {
    uint32_t my_nodes = 0, my_list = 0;
    for(int j = tid; j < ncur; j += NT) {
      const NodeRec rec = t.noderec[queue[cur][j]];
      double dd = 0.0;
#pragma unroll
      for(int d = 0; d < D; ++d) {
        const double tt = fmax(0.0, fmax(rec.bmin[d] - tmx[d], tmn[d] - rec.bmax[d]));
        dd = dd + tt * tt;
      }
      uint32_t r = 0;
      if(dd <= hmax2) {
        const uint32_t node_mask = ((1u << rec.nchild) - 1u) & ~rec.leafmask;
#if FSPH_SEARCH_TWO_LEVEL == 2
        // seed lists: nodes of at most super-cell size end the descent and are listed whole; bigger nodes are
        // descended and listed (flagged "leaf children only") when they have leaf children
        if(t.sub[queue[cur][j]] <= SUPER_ENTITIES) r = 0x80u;
        else r = (uint32_t)__popc(node_mask) | (rec.leafmask ? 0xc0u : 0u);
#else
        r = (uint32_t)__popc(node_mask) | (rec.leafmask ? 0x80u : 0u);
#endif
      }
      qres[j] = (uint8_t)r;
      my_nodes += r & 0x3fu;
      my_list += (r >> 7) & 1u;
    }
    uint32_t total;
    const uint32_t ex = block_scan_u32<NT>(my_nodes | (my_list << 16), total, scan_sm);
    uint32_t qpos = ex & 0xffffu, lpos = nlist + (ex >> 16);
    const uint32_t nnext_raw = total & 0xffffu;
    for(int j = tid; j < ncur; j += NT) {
      const uint32_t r = qres[j];
      if(r == 0) continue;
      const int nd = queue[cur][j];
      if(r & 0x3fu) {
        const NodeRec rec = t.noderec[nd];
        uint32_t node_mask = ((1u << rec.nchild) - 1u) & ~rec.leafmask;
        while(node_mask) {
          const int b = __ffs(node_mask) - 1;
          node_mask &= node_mask - 1;
          if(qpos < SW_QCAP) queue[cur ^ 1][qpos] = rec.first_child_node + __popc(~rec.leafmask & ((1u << b) - 1u));
          ++qpos;
        }
      }
      if(r & 0x80u) {
        if(MODE == 1) out[lpos] = (r & 0x40u) ? (nd | 0x40000000) : nd;
        ++lpos;
      }
    }
    nlist += total >> 16;
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
// super cell of a work cell = the one whose body range holds the work cell's first body
__global__ void __launch_bounds__(256) k_wc_super(Tree t, Search sr, const uint32_t * __restrict__ at, const uint32_t * __restrict__ ex) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if(g >= sr.nwc) return;
  const int v = t.cell[sr.wc_cell[g]];
  const int start = v >= 0 ? v : t.node_start[~v];
  sr.wc_super[g] = (int32_t)(ex[start] + (at[start] ? 1u : 0u)) - 1;
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
  sr.super_cell = s.b_super_cell.as<int32_t>();
  sr.sl_off = s.b_sl_off.as<uint64_t>();
  sr.wc_super = s.b_wc_super.as<int32_t>();
  sr.nsuper = (int)nsuper;
  FSPH_LAUNCH(k_super_compact, nbp, 256, s.stream, at, ex, n, sr.super_cell);
  FSPH_LAUNCH(k_wc_super, (unsigned)((sr.nwc + 255) / 256), 256, s.stream, s.tree, sr, at, ex);
  FSPH_LAUNCH((k_super_walk<D, 0>), nsuper, 64, s.stream, s.tree, s.cols(), sr, status);
  exclusive_scan_u64(s, sr.sl_off, sr.sl_off, nsuper, nullptr);
  uint64_t tot = 0;
  uint32_t st = 0;
  FSPH_CUDA_OK(cudaMemcpyAsync(&tot, sr.sl_off + nsuper, 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  if(st & ERR_QUEUE_OVERFLOW) throw std::runtime_error("neighbour search: node frontier of a super cell exceeds the shared-memory queue");
  s.b_sl_nodes.reserve((size_t)std::max<uint64_t>(tot, 1) * 4);
  sr.sl_nodes = s.b_sl_nodes.as<int32_t>();
  FSPH_LAUNCH((k_super_walk<D, 1>), nsuper, 64, s.stream, s.tree, s.cols(), sr, status);
  FSPH_CUDA_OK(cudaGetLastError());
}

template<int D, int MODE>
static void launch_search(System & s, int32_t * scratch, int scratch_cap, int grid_cap, unsigned long long * cursors,
  unsigned long long cand_cap, unsigned long long mask_cap, bool global_queue = false) {
  Search & sr = s.search;
  uint32_t * status = s.b_status.as<uint32_t>();
  if(global_queue) {
    // fall-back (MODE 0 / 1 only): persistent CTAs, each with its own frontier strip in global memory
    const int gq_grid = 148 * 6;
    s.b_search_gq.reserve((size_t)gq_grid * 3 * SR_QCAP_GLOBAL * 4);
    int32_t * gq = s.b_search_gq.as<int32_t>();
    for(int cls = 0; cls < 4; ++cls) {
      const int b = sr.class_off[cls], e = sr.class_off[cls + 1];
      if(e <= b) continue;
      const int grid = std::min(e - b, gq_grid);
      constexpr int M = MODE == 2 ? 1 : MODE;
      switch(cls) {
        case 0: FSPH_LAUNCH((k_search<D, 1, M, true>), grid, 32, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status, gq); break;
        case 1: FSPH_LAUNCH((k_search<D, 2, M, true>), grid, 64, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status, gq); break;
        case 2: FSPH_LAUNCH((k_search<D, 3, M, true>), grid, 96, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status, gq); break;
        default: FSPH_LAUNCH((k_search<D, 4, M, true>), grid, 128, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status, gq); break;
      }
    }
    FSPH_CUDA_OK(cudaGetLastError());
    return;
  }
  for(int cls = 0; cls < 4; ++cls) {
    const int b = sr.class_off[cls], e = sr.class_off[cls + 1];
    if(e <= b) continue;
    const int grid = MODE == 2 ? std::min(e - b, grid_cap) : e - b;
    switch(cls) {
      case 0: FSPH_LAUNCH((k_search<D, 1, MODE>), grid, 32, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status); break;
      case 1: FSPH_LAUNCH((k_search<D, 2, MODE>), grid, 64, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status); break;
      case 2: FSPH_LAUNCH((k_search<D, 3, MODE>), grid, 96, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status); break;
      default: FSPH_LAUNCH((k_search<D, 4, MODE>), grid, 128, s.stream, s.tree, s.cols(), sr, b, e, scratch, scratch_cap, cursors, cand_cap, mask_cap, status); break;
    }
  }
  FSPH_CUDA_OK(cudaGetLastError());
}

template<int D>
static void neighbor_search_t(System & s) {
  Search & sr = s.search;
  const int nwc = sr.nwc;
  if(nwc <= 0) throw std::runtime_error("no work cells");
  uint32_t * status = s.b_status.as<uint32_t>();
  s.b_cand_off.reserve((size_t)(nwc + 1) * 8);
  s.b_mask_off.reserve((size_t)(nwc + 1) * 8);
  s.b_ncand.reserve((size_t)(nwc + 1) * 4);
  sr.cand_off = s.b_cand_off.as<uint64_t>();
  sr.mask_off = s.b_mask_off.as<uint64_t>();
  sr.ncand = s.b_ncand.as<int32_t>();
  unsigned long long * cursors = reinterpret_cast<unsigned long long *>(s.b_status.as<uint32_t>() + 8);
#if FSPH_SEARCH_TWO_LEVEL
  {
    PhaseScope ph(s, "search:super_lists");
    build_super_lists<D>(s);
  }
#endif
  // single pass when the previous search on this system gives a capacity estimate
  // (with ghosts the body count drifts a little from step to step: the estimate is scaled by the ratio)
  if(getenv("FSPH_SEARCH_FORCE_GLOBAL_QUEUE")) s.search_global_queue = true;   // test hook
  const bool est_ok = !s.search_global_queue && s.est_cand > 0 && s.est_n > 0 && (s.est_n == s.n || (s.comm && std::llabs(s.est_n - s.n) < s.est_n / 4));
  if(est_ok) {
    if(s.est_n != s.n) {
      const double r = (double)s.n / (double)s.est_n;
      if(r > 1.0) { s.est_cand = (unsigned long long)(s.est_cand * r); s.est_mask = (unsigned long long)(s.est_mask * r); }
      s.est_n = s.n;
    }
    const int grid_cap = 148 * 12, scratch_cap = 16384;
    const unsigned long long cand_cap = s.est_cand + s.est_cand / 4 + 65536, mask_cap = s.est_mask + s.est_mask / 4 + 65536;
    s.b_cand.reserve((size_t)cand_cap * 4);
    s.b_mask.reserve((size_t)mask_cap * 4);
    s.b_search_scratch.reserve((size_t)grid_cap * scratch_cap * 4);
    sr.cand = s.b_cand.as<int32_t>();
    sr.mask = s.b_mask.as<uint32_t>();
    FSPH_CUDA_OK(cudaMemsetAsync(cursors, 0, 16, s.stream));
    launch_search<D, 2>(s, s.b_search_scratch.as<int32_t>(), scratch_cap, grid_cap, cursors, cand_cap, mask_cap);
    unsigned long long tot[2];
    uint32_t st = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(tot, cursors, 16, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(!(st & (ERR_SEARCH_CAPACITY | ERR_QUEUE_OVERFLOW))) {
      s.est_cand = tot[0];
      s.est_mask = tot[1];
      return;
    }
    // capacity estimate too small / frontier too large for shared memory: clear the flags, take the exact two-pass path
    if(st & ERR_QUEUE_OVERFLOW) s.search_global_queue = true;
    uint32_t cleared = st & ~(uint32_t)(ERR_SEARCH_CAPACITY | ERR_QUEUE_OVERFLOW);
    FSPH_CUDA_OK(cudaMemcpyAsync(status, &cleared, 4, cudaMemcpyHostToDevice, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  }
  // exact two-pass path (count, prefix sum, fill); with the frontier in global memory when shared memory is too small
  auto clear_flag = [&](uint32_t st, uint32_t bit) {
    const uint32_t cleared = st & ~bit;
    FSPH_CUDA_OK(cudaMemcpyAsync(status, &cleared, 4, cudaMemcpyHostToDevice, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  };
  auto two_pass = [&](bool gq) -> bool {
    launch_search<D, 0>(s, nullptr, 0, 0, nullptr, 0, 0, gq);
    exclusive_scan_u64(s, sr.cand_off, sr.cand_off, nwc, nullptr);
    exclusive_scan_u64(s, sr.mask_off, sr.mask_off, nwc, nullptr);
    uint64_t tot[2];
    uint32_t st = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&tot[0], sr.cand_off + nwc, 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&tot[1], sr.mask_off + nwc, 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(st & ERR_QUEUE_OVERFLOW) {
      clear_flag(st, ERR_QUEUE_OVERFLOW);
      return false;
    }
    s.b_cand.reserve((size_t)std::max<uint64_t>(tot[0], 1) * 4);
    s.b_mask.reserve((size_t)std::max<uint64_t>(tot[1], 1) * 4);
    sr.cand = s.b_cand.as<int32_t>();
    sr.mask = s.b_mask.as<uint32_t>();
    launch_search<D, 1>(s, nullptr, 0, 0, nullptr, 0, 0, gq);
    s.est_cand = tot[0];
    s.est_mask = tot[1];
    s.est_n = s.n;
    return true;
  };
  if(!s.search_global_queue && two_pass(false)) return;
  s.search_global_queue = true;   // stays on for this system: its cells need the big frontier
  s.est_cand = 0;                  // the single-pass kernel has no global-queue variant
  if(!two_pass(true)) throw std::runtime_error("neighbour search: node frontier exceeds even the global-memory queue");
}

void neighbor_search(System & s) {
  if(s.dim == 1) neighbor_search_t<1>(s);
  else if(s.dim == 2) neighbor_search_t<2>(s);
  else neighbor_search_t<3>(s);
  s.search_valid = true;
}

// ---- CSR export for the parity tests --------------------------------------------------------------------
__global__ void __launch_bounds__(SR_NT) k_export(Tree t, Cols c, Search sr, const uint64_t * __restrict__ off, int64_t * __restrict__ ids) {
  const int g = blockIdx.x;
  const WorkCell wc = load_work_cell(t, sr, g);
  const int Ppad = (wc.P + 31) & ~31;
  const uint64_t co = sr.cand_off[g], mo = sr.mask_off[g];
  const int C = sr.ncand[g];
  const int nw = (C + 31) >> 5;
  for(int k = threadIdx.x; k < wc.P; k += SR_NT) {
    uint64_t o = off[wc.start + k];
    for(int w = 0; w < nw; ++w) {
      uint32_t m = sr.mask[mo + (uint64_t)w * Ppad + k];
      while(m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        ids[o++] = c.id[sr.cand[co + w * 32 + b]];
      }
    }
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
  FSPH_LAUNCH(k_export, s.search.nwc, SR_NT, s.stream, s.tree, s.cols(), s.search, o, tmp.as<int64_t>());
  FSPH_CUDA_OK(cudaMemcpyAsync(ids->data(), tmp.p, (size_t)tot * 8, cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  tmp.release();
  for(int64_t i = 0; i < n; ++i) std::sort(ids->begin() + off[i], ids->begin() + off[i + 1]);
  drop_ghost_rows();  // ghost rows are empty, so the id array is unchanged
}

}  // namespace fsph
