// fsph_fmm.cu -- FMM tree gravity (compiled with -fmad=false: the acceptance test must decide exactly as the reference).
//
// Replaces body_system::gravitation_fmm (mpisph/bodies_system.h:282-288) =
//   tree_topology::traversal_fmm        include/tree_topology/tree_topology.h:448-795
//   tree_geometry::mac                  include/tree_topology/tree_geometry.h:343-350
//   fmm::taylor_c2c / taylor_p2c / fmm_p2p / fmm_c2p / gravitation_fc / gravitation_p2p   include/physics/fmm.h:37-159
//
// The reference walks pairs (sink cell, source cell) breadth-first from (root, root).  With its constants
// (fmm_sub_entities_ = 0, tree_topology.h:2122) the walk is:
//   both cells are bodies                 -> direct body-body term (skipped for the body itself, fmm.h:131)
//   same node                             -> all ordered pairs of its children (tree_topology.h:512-527)
//   r1 + r2 < theta * |X1 - X2|           -> accepted: a sink node adds the source's monopole evaluated at its own
//                                            centre of mass to its (pc, fc) pair; a sink body adds it to (g_pot, g_acc)
//   otherwise                             -> split the cell with the larger radius (the sink on ties: the source)
// and afterwards every body adds, for each ancestor that was a sink, fc and pc - (x - X)·fc (first-order Taylor, :95-110).
// The result is a set of independent additive terms, so the order in which pairs are processed is free.  Here the
// pending pairs live on a device stack that is consumed in chunks of FMM_CHUNK pairs (bounded memory: the reference's
// breadth-first queue holds every pair of a level at once); one thread handles one pair; children pairs are appended
// with one atomicAdd per warp; node and body accumulators are updated with fp64 atomics (sum order is not fixed:
// results agree with the reference to rounding, far inside the 1e-10 criterion, but are not bit-reproducible).
#include <stdexcept>
#include <algorithm>
#include "fsph_internal.h"

namespace fsph {

constexpr int64_t FMM_CHUNK = 1 << 22;       // pairs consumed per launch
constexpr int64_t FMM_OUT_FACTOR = 24;       // children pairs per consumed pair the output strip can take

struct FmmNodes {
  double * pc;
  double * fc[3];
  uint32_t * affected;
};

__device__ __forceinline__ double dist3(const double * a, const double * b) {
  // flecsi::distance, space_vector.h:46-58
  double s = 0.0;
#pragma unroll
  for(int d = 0; d < 3; ++d) s = s + (a[d] - b[d]) * (a[d] - b[d]);
  return sqrt(s);
}

// does cell cc hold a body of the index range [lo, hi)?  (the sinks of a rank: "iam_owner", tree_topology.h:518-524,602)
__device__ __forceinline__ bool cell_in_range(const Tree & t, int cc, int64_t lo, int64_t hi) {
  const int v = t.cell[cc];
  const int64_t b = v >= 0 ? v : t.node_start[~v];
  const int64_t e = v >= 0 ? b + 1 : b + t.sub[~v];
  return b < hi && e > lo;
}

// PRUNE: only sink cells with bodies in [own_lo, own_hi) are followed (several ranks share one global tree)
template<bool PRUNE>
__global__ void __launch_bounds__(256) k_fmm_pairs(Tree t, Cols c, FmmNodes fn, const uint64_t * __restrict__ in, int64_t nin,
  uint64_t * __restrict__ out, unsigned long long * __restrict__ nout, unsigned long long out_cap, double theta, double gc,
  uint32_t * __restrict__ status, int64_t own_lo, int64_t own_hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nemit = 0;       // number of pairs this thread appends
  int mode = 0;        // 1: same node, 2: split sink, 3: split source
  int cb1 = 0, ce1 = 0, cb2 = 0, ce2 = 0;
  uint32_t c1 = 0, c2 = 0;
  uint32_t keep = 0xffu;  // PRUNE: children of the sink that hold bodies of this rank
  if(i < nin) {
    const uint64_t pr = in[i];
    c1 = (uint32_t)(pr >> 32);
    c2 = (uint32_t)pr;
    const int v1 = t.cell[c1], v2 = t.cell[c2];
    if(v1 >= 0 && v2 >= 0) {
      if(c1 != c2) {
        // fmm::gravitation_p2p, fmm.h:37-46
        double x1[3], x2[3];
#pragma unroll
        for(int d = 0; d < 3; ++d) { x1[d] = c.f[F_POS + d][v1]; x2[d] = c.f[F_POS + d][v2]; }
        const double sm = c.f[F_MASS][v2];
        const double dist = dist3(x1, x2);
        atomicAdd(&c.f[F_GPOT][v1], -gc * sm / dist);
        const double k = -gc * sm / (dist * dist * dist);
#pragma unroll
        for(int d = 0; d < 3; ++d) atomicAdd(&c.f[F_GACC + d][v1], k * (x1[d] - x2[d]));
      }
    }
    else if(c1 == c2) {
      const int nd = ~v1;
      cb1 = t.child_begin[nd];
      ce1 = t.child_end[nd];
      mode = 1;
      if(PRUNE) {
        keep = 0;
        for(int a = cb1; a < ce1; ++a) keep |= (cell_in_range(t, a, own_lo, own_hi) ? 1u : 0u) << (a - cb1);
      }
      nemit = (PRUNE ? __popc(keep) : (ce1 - cb1)) * (ce1 - cb1);
    }
    else {
      double X1[3], X2[3], r1 = 0.0, r2 = 0.0, m2;
      if(v1 < 0) {
#pragma unroll
        for(int d = 0; d < 3; ++d) X1[d] = t.com[d][~v1];
        r1 = t.radius[~v1];
      }
      else {
#pragma unroll
        for(int d = 0; d < 3; ++d) X1[d] = c.f[F_POS + d][v1];
      }
      if(v2 < 0) {
#pragma unroll
        for(int d = 0; d < 3; ++d) X2[d] = t.com[d][~v2];
        r2 = t.radius[~v2];
        m2 = t.mass[~v2];
      }
      else {
#pragma unroll
        for(int d = 0; d < 3; ++d) X2[d] = c.f[F_POS + d][v2];
        m2 = c.f[F_MASS][v2];
      }
      const double dist = dist3(X1, X2);
      if(r1 + r2 < theta * dist) {  // tree_geometry::mac
        // fmm::gravitation_fc, fmm.h:51-84 (the same expression for node and body sources)
        const double d3 = dist * dist * dist;
        const double pot = -gc * m2 / dist;
        double f[3];
#pragma unroll
        for(int d = 0; d < 3; ++d) f[d] = -gc * m2 * (X1[d] - X2[d]) / d3;
        if(v1 < 0) {
          const int nd = ~v1;
          atomicAdd(&fn.pc[nd], pot);
#pragma unroll
          for(int d = 0; d < 3; ++d) atomicAdd(&fn.fc[d][nd], f[d]);
          fn.affected[nd] = 1u;
        }
        else {
          atomicAdd(&c.f[F_GPOT][v1], pot);
#pragma unroll
          for(int d = 0; d < 3; ++d) atomicAdd(&c.f[F_GACC + d][v1], f[d]);
        }
      }
      else if(r1 > r2) {
        cb1 = t.child_begin[~v1];
        ce1 = t.child_end[~v1];
        mode = 2;
        if(PRUNE) {
          keep = 0;
          for(int a = cb1; a < ce1; ++a) keep |= (cell_in_range(t, a, own_lo, own_hi) ? 1u : 0u) << (a - cb1);
        }
        nemit = PRUNE ? __popc(keep) : ce1 - cb1;
      }
      else {
        cb2 = t.child_begin[~v2];
        ce2 = t.child_end[~v2];
        mode = 3;
        nemit = ce2 - cb2;
      }
    }
  }
  // one reservation per warp
  const int lane = threadIdx.x & 31;
  int incl = nemit;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if(lane >= o) incl += y;
  }
  const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if(lane == 31 && wtotal > 0) base = atomicAdd(nout, (unsigned long long)wtotal);
  base = __shfl_sync(0xffffffffu, base, 31);
  if(nemit == 0) return;
  unsigned long long o = base + (unsigned long long)(incl - nemit);
  if(o + (unsigned long long)nemit > out_cap) {
    atomicOr(status, (uint32_t)ERR_QUEUE_OVERFLOW);
    return;
  }
  if(mode == 1) {
    // (k1,k1), then (k1,k2) and (k2,k1) for k2 > k1: every ordered pair of children
    for(int a = cb1; a < ce1; ++a) {
      if(PRUNE && !((keep >> (a - cb1)) & 1u)) continue;
      for(int b = cb1; b < ce1; ++b) out[o++] = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    }
  }
  else if(mode == 2) {
    for(int a = cb1; a < ce1; ++a) {
      if(PRUNE && !((keep >> (a - cb1)) & 1u)) continue;
      out[o++] = ((uint64_t)(uint32_t)a << 32) | c2;
    }
  }
  else {
    for(int b = cb2; b < ce2; ++b) out[o++] = ((uint64_t)c1 << 32) | (uint32_t)b;
  }
}

__global__ void __launch_bounds__(256) k_leaf_parent(Tree t, int64_t nc, int32_t * __restrict__ leaf_parent) {
  const int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(cc >= nc) return;
  const int v = t.cell[cc];
  if(v >= 0) leaf_parent[v] = t.cell_parent[cc];
}

// fmm::fmm_c2p / interaction_c2p (fmm.h:95-120) for every ancestor that received an interaction
__global__ void __launch_bounds__(256) k_fmm_c2p(Tree t, Cols c, FmmNodes fn, const int32_t * __restrict__ leaf_parent, int64_t lo, int64_t hi) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= hi) return;
  double x[3], g[3] = {0.0, 0.0, 0.0}, pot = 0.0;
#pragma unroll
  for(int d = 0; d < 3; ++d) x[d] = c.f[F_POS + d][i];
  for(int nd = leaf_parent[i]; nd >= 0; nd = t.node_parent[nd]) {
    if(!fn.affected[nd]) continue;
    double p = fn.pc[nd];
#pragma unroll
    for(int d = 0; d < 3; ++d) {
      const double f = fn.fc[d][nd];
      p += -(x[d] - t.com[d][nd]) * f;
      g[d] += f;
    }
    pot += p;
  }
  c.f[F_GPOT][i] += pot;
#pragma unroll
  for(int d = 0; d < 3; ++d) c.f[F_GACC + d][i] += g[d];
}

void gravitation_fmm(System & s) {
  if(s.comm) dist_gravitation_fmm(s);
  else gravitation_fmm_range(s, 0, s.n);
}

void gravitation_fmm_range(System & s, int64_t own_lo, int64_t own_hi) {
  if(s.dim != 3) throw std::runtime_error("fmm needs 3 dimensions (bodies_system.h:283)");
  if(!s.tree_valid) throw std::runtime_error("gravitation_fmm before update_iteration");
  const int64_t n = s.n;
  const bool prune = own_lo > 0 || own_hi < n;
  const Tree & t = s.tree;
  const int nn = std::max(t.nnodes, 1);
  // node accumulators: pc, fc[3], affected
  s.b_fmm_nodes.reserve((size_t)nn * (4 * 8 + 4));
  FSPH_CUDA_OK(cudaMemsetAsync(s.b_fmm_nodes.p, 0, (size_t)nn * (4 * 8 + 4), s.stream));
  FmmNodes fn;
  fn.pc = s.b_fmm_nodes.as<double>();
  for(int d = 0; d < 3; ++d) fn.fc[d] = fn.pc + (size_t)nn * (1 + d);
  fn.affected = reinterpret_cast<uint32_t *>(fn.pc + (size_t)nn * 4);
  // pair stack + output strip
  const int64_t out_cap = FMM_CHUNK * FMM_OUT_FACTOR;
  s.b_fmm_out.reserve((size_t)out_cap * 8);
  if(s.b_fmm_stack.cap < (size_t)out_cap * 8 * 2) s.b_fmm_stack.reserve((size_t)out_cap * 8 * 2);
  unsigned long long * cnt = reinterpret_cast<unsigned long long *>(s.b_status.as<uint32_t>() + 12);
  uint32_t * status = s.b_status.as<uint32_t>();
  const uint64_t root_pair = 0;  // (cell 0, cell 0): the root
  FSPH_CUDA_OK(cudaMemcpyAsync(s.b_fmm_stack.p, &root_pair, 8, cudaMemcpyHostToDevice, s.stream));
  int64_t top = 1;
  unsigned long long interactions = 0;
  while(top > 0) {
    const int64_t take = std::min<int64_t>(top, FMM_CHUNK);
    const uint64_t * in = s.b_fmm_stack.as<uint64_t>() + (top - take);
    FSPH_CUDA_OK(cudaMemsetAsync(cnt, 0, 8, s.stream));
    if(prune)
      FSPH_LAUNCH((k_fmm_pairs<true>), (unsigned)((take + 255) / 256), 256, s.stream, t, s.cols(), fn, in, take, s.b_fmm_out.as<uint64_t>(), cnt,
        (unsigned long long)out_cap, s.par.fmm_macangle, s.par.gravitational_constant, status, own_lo, own_hi);
    else
      FSPH_LAUNCH((k_fmm_pairs<false>), (unsigned)((take + 255) / 256), 256, s.stream, t, s.cols(), fn, in, take, s.b_fmm_out.as<uint64_t>(), cnt,
        (unsigned long long)out_cap, s.par.fmm_macangle, s.par.gravitational_constant, status, own_lo, own_hi);
    unsigned long long nout = 0;
    uint32_t st = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&nout, cnt, 8, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaMemcpyAsync(&st, status, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if((st & ERR_QUEUE_OVERFLOW) || nout > (unsigned long long)out_cap)
      throw std::runtime_error("gravitation_fmm: a chunk of pairs produced more children pairs than the output strip holds");
    interactions += (unsigned long long)take;
    top -= take;
    if(nout) {
      s.b_fmm_stack.grow_keep((size_t)(top + (int64_t)nout) * 8, (size_t)top * 8, s.stream);
      FSPH_CUDA_OK(cudaMemcpyAsync(s.b_fmm_stack.as<uint64_t>() + top, s.b_fmm_out.p, (size_t)nout * 8, cudaMemcpyDeviceToDevice, s.stream));
      top += (int64_t)nout;
    }
  }
  s.fmm_pairs = interactions;
  // bodies <- ancestors
  s.b_tmp[0].reserve((size_t)n * 4);
  int32_t * leaf_parent = s.b_tmp[0].as<int32_t>();
  FSPH_LAUNCH(k_leaf_parent, (unsigned)((t.ncells + 255) / 256), 256, s.stream, t, (int64_t)t.ncells, leaf_parent);
  FSPH_LAUNCH(k_fmm_c2p, (unsigned)((own_hi - own_lo + 255) / 256), 256, s.stream, t, s.cols(), fn, leaf_parent, own_lo, own_hi);
  FSPH_CUDA_OK(cudaGetLastError());
}

}  // namespace fsph
