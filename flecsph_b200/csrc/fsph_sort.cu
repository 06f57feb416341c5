// fsph_sort.cu -- hand-written device scan and LSD radix sort (64-bit Morton keys + 32-bit index payload).
//
// Replaces the local std::sort of whole 368-byte body records in psort::psort (mpisph/psort.h:247,310)
// and is reused as the one-pass stable counting sort that brings the octree cells into level order.
// HBM-bound: each pass reads key+index once and writes them once (12 B + 12 B per particle); passes whose
// digit is constant over all keys are skipped (the digit histograms of all 8 passes come from one read).
#include <stdexcept>
#include "fsph_internal.h"

namespace fsph {

std::atomic<unsigned long long> g_launches{0};

void DevBuf::reserve(size_t bytes) {
  if(bytes <= cap) return;
  if(p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if(e != cudaSuccess) {
    p = nullptr;
    throw std::runtime_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  }
  cap = want;
}
void DevBuf::reserve_exact(size_t bytes) {
  if(bytes <= cap) return;
  if(p) cudaFree(p);
  p = nullptr;
  cap = 0;
  cudaError_t e = cudaMalloc(&p, bytes);
  if(e != cudaSuccess) {
    p = nullptr;
    throw std::runtime_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  }
  cap = bytes;
}
void DevBuf::grow_keep(size_t bytes, size_t keep, cudaStream_t st) {
  if(bytes <= cap) return;
  void * old = p;
  p = nullptr;
  cap = 0;
  try { reserve(bytes + bytes / 4); }
  catch(...) { p = old; throw; }
  if(old) {
    if(keep) {
      cudaMemcpyAsync(p, old, keep, cudaMemcpyDeviceToDevice, st);
      cudaStreamSynchronize(st);
    }
    cudaFree(old);
  }
}
void DevBuf::release() {
  if(p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

void release_buffers(System & s) {
  DevBuf * all[] = {&s.col_gd, &s.col_id, &s.col_key, &s.col_type, &s.col_nneigh, &s.scratch_col, &s.b_perm, &s.b_keys2, &s.b_perm2,
    &s.b_node_parent, &s.b_fmm_nodes, &s.b_fmm_out, &s.b_fmm_stack, &s.b_cell, &s.b_cell_parent, &s.b_cell_desc, &s.b_cell_desc2,
    &s.b_child_begin, &s.b_child_end, &s.b_node_start, &s.b_node_depth, &s.b_sub, &s.b_noderec, &s.b_leafrec, &s.b_nmass, &s.b_nradius,
    &s.b_nlap, &s.b_wc_super, &s.b_super_cell, &s.b_sl_off, &s.b_sl_nodes, &s.b_wc_cell, &s.b_cand,
    &s.b_cnt0, &s.b_cnt1, &s.b_status, &s.b_red, &s.b_levels, &s.b_search_scratch, &s.b_blk_wc, &s.b_blk_meta, &s.b_dir, &s.b_nl, &s.b_super_queue, &s.b_wcinfo, &s.b_wcrep, &s.b_wcand, &s.b_sl_rec, &s.b_sl_len, &s.b_sl_stage,
    &s.b_pair_a, &s.b_pair_pi};
  for(DevBuf * b : all) b->release();
  for(int f = 0; f < NF64; ++f) s.col[f].release();
  for(int i = 0; i < 6; ++i) s.b_tmp[i].release();
  for(int i = 0; i < 7; ++i) s.b_gather[i].release();
  for(int d = 0; d < 3; ++d) { s.b_ncom[d].release(); s.b_nbmin[d].release(); s.b_nbmax[d].release(); }
}

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan helper (NT threads, NT multiple of 32, <= 1024)
template<typename T, int NT>
__device__ __forceinline__ T block_exclusive_scan(T v, T & total, T * sm /* NT/32 + 1 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if(lane >= o) x += y;
  }
  if(lane == 31) sm[w] = x;
  __syncthreads();
  if(w == 0) {
    T t = lane < NT / 32 ? sm[lane] : T(0);
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, t, o);
      if(lane >= o) t += y;
    }
    if(lane < NT / 32) sm[lane] = t;  // inclusive over warps
  }
  __syncthreads();
  T base = w > 0 ? sm[w - 1] : T(0);
  total = sm[NT / 32 - 1];
  __syncthreads();
  return base + x - v;
}

constexpr int SCAN_NT = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_NT * SCAN_IPT;

template<typename T>
__global__ void __launch_bounds__(SCAN_NT) k_scan_reduce(const T * __restrict__ in, T * __restrict__ bsum, int64_t n) {
  __shared__ T sm[SCAN_NT / 32 + 1];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
  T acc = 0;
#pragma unroll
  for(int i = 0; i < SCAN_IPT; ++i)
    if(base + i < n) acc += in[base + i];
  T total;
  block_exclusive_scan<T, SCAN_NT>(acc, total, sm);
  if(threadIdx.x == 0) bsum[blockIdx.x] = total;
}

template<typename T>
__global__ void __launch_bounds__(SCAN_NT) k_scan_final(const T * in, T * out, const T * __restrict__ boff, int64_t n) {
  __shared__ T sm[SCAN_NT / 32 + 1];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
  T v[SCAN_IPT];
  T acc = 0;
#pragma unroll
  for(int i = 0; i < SCAN_IPT; ++i) {
    v[i] = base + i < n ? in[base + i] : T(0);
    acc += v[i];
  }
  T total;
  T ex = block_exclusive_scan<T, SCAN_NT>(acc, total, sm) + (boff ? boff[blockIdx.x] : T(0));
#pragma unroll
  for(int i = 0; i < SCAN_IPT; ++i) {
    if(base + i < n) {
      out[base + i] = ex;
      ex += v[i];
      if(base + i == n - 1) out[n] = ex;
    }
  }
}

template<typename T>
static void scan_impl(System & s, const T * in, T * out, int64_t n, int depth) {
  if(n <= 0) {
    FSPH_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(T), s.stream));
    return;
  }
  int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if(nb == 1) {
    FSPH_LAUNCH((k_scan_final<T>), 1, SCAN_NT, s.stream, in, out, nullptr, n);
    return;
  }
  if(depth > 2) throw std::runtime_error("scan recursion too deep");
  DevBuf & b = s.b_tmp[3 + depth];
  b.reserve((nb + 1) * sizeof(T));
  T * bs = b.as<T>();
  FSPH_LAUNCH((k_scan_reduce<T>), (unsigned)nb, SCAN_NT, s.stream, in, bs, n);
  scan_impl<T>(s, bs, bs, nb, depth + 1);
  FSPH_LAUNCH((k_scan_final<T>), (unsigned)nb, SCAN_NT, s.stream, in, out, bs, n);
}

// out has n+1 entries: out[n] = total.  In-place (in == out) is allowed.
void exclusive_scan_u32(System & s, const uint32_t * in, uint32_t * out, int64_t n, uint32_t *) { scan_impl<uint32_t>(s, in, out, n, 0); }
void exclusive_scan_u64(System & s, const uint64_t * in, uint64_t * out, int64_t n, uint64_t *) { scan_impl<uint64_t>(s, in, out, n, 0); }

// ---------------------------------------------------------------------------------------------
constexpr int RS_NT = 256, RS_IPT = 16, RS_TILE = RS_NT * RS_IPT, RS_WARPS = RS_NT / 32;

__global__ void __launch_bounds__(256) k_hist_all(const uint64_t * __restrict__ k, int64_t n, uint32_t * __restrict__ hist /*[8][256]*/) {
  __shared__ uint32_t sh[8 * 256];
  for(int i = threadIdx.x; i < 8 * 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key = k[i];
#pragma unroll
    for(int p = 0; p < 8; ++p) atomicAdd(&sh[p * 256 + ((key >> (8 * p)) & 255)], 1u);
  }
  __syncthreads();
  for(int i = threadIdx.x; i < 8 * 256; i += blockDim.x)
    if(sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void __launch_bounds__(RS_NT) k_radix_count(const uint64_t * __restrict__ k, int64_t n, int shift,
  uint32_t * __restrict__ tile_hist, int ntiles) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for(int r = 0; r < RS_IPT; ++r) {
    int64_t i = base + r * RS_NT + threadIdx.x;
    if(i < n) atomicAdd(&sh[(k[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  tile_hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

// VMODE: 0 keys only, 1 values from vin, 2 values = source index
// Ranking: per warp, 16 rounds of 32 consecutive items; __match_any_sync groups the lanes of a round by digit, the
// group's leader advances the warp's digit counter, so an item's rank follows the input order (stable).  The tile is
// then put in digit order in shared memory and leaves in runs: items of one digit are consecutive in the tile and in
// global memory, so the stores of a warp cover a few whole sectors instead of one sector per item (29 sectors per
// store request before, on the pass whose digit changes fastest along the nearly sorted input).
constexpr int RS_SMEM = RS_TILE * 8 + RS_TILE * 4 + RS_WARPS * 256 * 4 + 256 * 4 + 256 * 4 + 64;
#ifndef FSPH_RS_OCC
#define FSPH_RS_OCC 2
#endif
template<int VMODE>
__global__ void __launch_bounds__(RS_NT, FSPH_RS_OCC) k_radix_scatter(const uint64_t * __restrict__ kin, const uint32_t * __restrict__ vin,
  uint64_t * __restrict__ kout, uint32_t * __restrict__ vout, int64_t n, int shift, const uint32_t * __restrict__ tile_off,
  int ntiles) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  uint64_t * const skey = reinterpret_cast<uint64_t *>(rs_raw);                                  // [RS_TILE] tile in digit order
  uint32_t * const sval = reinterpret_cast<uint32_t *>(rs_raw + RS_TILE * 8);                    // [RS_TILE]
  uint32_t (* const wcnt)[256] = reinterpret_cast<uint32_t (*)[256]>(rs_raw + RS_TILE * 12);     // [RS_WARPS][256]
  uint32_t * const lstart = reinterpret_cast<uint32_t *>(rs_raw + RS_TILE * 12 + RS_WARPS * 1024);   // [256] first tile slot of a digit
  uint32_t * const gdelta = lstart + 256;                                                        // [256] global position - tile slot
  uint32_t * const scan_sm = gdelta + 256;                                                       // [RS_WARPS]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for(int i = threadIdx.x; i < RS_WARPS * 256; i += RS_NT) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
  const int64_t wbase = tbase + (int64_t)w * (32 * RS_IPT);
  uint64_t key[RS_IPT];
  uint32_t rank[RS_IPT];
#pragma unroll
  for(int r = 0; r < RS_IPT; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool valid = i < n;
    key[r] = valid ? kin[i] : 0;
    uint32_t d = valid ? (uint32_t)((key[r] >> shift) & 255) : (256u + lane);
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if(valid && lane == leader) {
      old = wcnt[w][d];
      wcnt[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {
    // thread d: the warps' counts of digit d -> exclusive prefix over the warps; the digits' tile totals -> exclusive scan
    const int d = threadIdx.x;  // RS_NT == 256 digits
    uint32_t run = 0;
#pragma unroll
    for(int ww = 0; ww < RS_WARPS; ++ww) {
      const uint32_t c = wcnt[ww][d];
      wcnt[ww][d] = run;
      run += c;
    }
    uint32_t x = run;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if(lane >= o) x += y;
    }
    if(lane == 31) scan_sm[w] = x;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for(int ww = 0; ww < RS_WARPS; ++ww)
      if(ww < w) base += scan_sm[ww];
    const uint32_t first = base + x - run;
    lstart[d] = first;
    gdelta[d] = tile_off[(size_t)d * ntiles + blockIdx.x] - first;
  }
  __syncthreads();
#pragma unroll
  for(int r = 0; r < RS_IPT; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if(i < n) {
      const uint32_t d = (uint32_t)((key[r] >> shift) & 255);
      const uint32_t slot = lstart[d] + wcnt[w][d] + rank[r];
      skey[slot] = key[r];
      if(VMODE == 1) sval[slot] = vin[i];
      if(VMODE == 2) sval[slot] = (uint32_t)i;
    }
  }
  __syncthreads();
  const int tile_n = (int)(n - tbase < (int64_t)RS_TILE ? n - tbase : (int64_t)RS_TILE);
  for(int j = threadIdx.x; j < tile_n; j += RS_NT) {
    const uint64_t k = skey[j];
    const uint32_t dst = gdelta[(uint32_t)((k >> shift) & 255)] + (uint32_t)j;
    kout[dst] = k;
    if(VMODE != 0) vout[dst] = sval[j];
  }
}

static void radix_pass_mode(System & s, const uint64_t * kin, const uint32_t * vin, uint64_t * kout, uint32_t * vout, int64_t n,
  int shift, int vmode) {
  if(n <= 0) return;
  if(n >= (int64_t(1) << 32)) throw std::runtime_error("radix sort: more than 2^32 items");
  int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  s.b_tmp[2].reserve(((size_t)256 * ntiles + 1) * sizeof(uint32_t));
  uint32_t * th = s.b_tmp[2].as<uint32_t>();
  FSPH_LAUNCH(k_radix_count, ntiles, RS_NT, s.stream, kin, n, shift, th, ntiles);
  exclusive_scan_u32(s, th, th, (int64_t)256 * ntiles, nullptr);
  if(vmode == 0) FSPH_LAUNCH_SMEM((k_radix_scatter<0>), ntiles, RS_NT, RS_SMEM, s.stream, kin, vin, kout, vout, n, shift, th, ntiles);
  else if(vmode == 1) FSPH_LAUNCH_SMEM((k_radix_scatter<1>), ntiles, RS_NT, RS_SMEM, s.stream, kin, vin, kout, vout, n, shift, th, ntiles);
  else FSPH_LAUNCH_SMEM((k_radix_scatter<2>), ntiles, RS_NT, RS_SMEM, s.stream, kin, vin, kout, vout, n, shift, th, ntiles);
  FSPH_CUDA_OK(cudaGetLastError());
}

void radix_pass(System & s, const uint64_t * kin, const uint32_t * vin, uint64_t * kout, uint32_t * vout, int64_t n, int shift) {
  radix_pass_mode(s, kin, vin, kout, vout, n, shift, vin ? 1 : 0);
}

// Runs of equal high key halves after the passes on the four high digits: the stable passes left each run in its order
// of the previous step; it is finished here by an insertion sort on the full key (runs are almost always single bodies:
// 2^32 high halves for at most 2^31 bodies).  A run longer than 32 sets *flag (the caller then sorts all eight digits).
__global__ void __launch_bounds__(256) k_fix_runs(uint64_t * __restrict__ k, uint32_t * __restrict__ v, int64_t n, uint32_t * __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint64_t hi = k[i] >> 32;
  if(hi == 0xffffffffull) {
    // the run of the largest high half holds the remote copies of the previous iteration (key ~0, dropped by the caller:
    // their order is irrelevant); a real key up there (a body in the last 2^-10 of the range on every axis) needs the full sort
    if(k[i] != ~0ull) *flag = 1u;
    return;
  }
  if(i > 0 && (k[i - 1] >> 32) == hi) return;        // not the first of its run
  if(i + 1 >= n || (k[i + 1] >> 32) != hi) return;   // run of one
  int L = 2;
  while(i + L < n && (k[i + L] >> 32) == hi) {
    if(++L > 32) { *flag = 1u; return; }
  }
  for(int a = 1; a < L; ++a) {
    const uint64_t ka = k[i + a];
    const uint32_t va = v[i + a];
    int b = a - 1;
    while(b >= 0 && k[i + b] > ka) {
      k[i + b + 1] = k[i + b];
      v[i + b + 1] = v[i + b];
      --b;
    }
    k[i + b + 1] = ka;
    v[i + b + 1] = va;
  }
}

// Sort (key, index) pairs: on entry k0 holds the keys (v0 contents ignored: the payload starts as the
// identity); on return k0/v0 point at the sorted keys / permutation and k1/v1 at the scratch pair.
// The bodies arrive in the key order of the previous step, so the low key half only matters inside runs of equal high
// halves: four stable passes on the high digits, then k_fix_runs; eight passes when a run is too long for that.
void radix_sort_pairs(System & s, uint64_t *& k0, uint32_t *& v0, uint64_t *& k1, uint32_t *& v1, int64_t n, const uint64_t * keys_src) {
  s.b_tmp[1].reserve(8 * 256 * sizeof(uint32_t));
  uint32_t * hist = s.b_tmp[1].as<uint32_t>();
  FSPH_CUDA_OK(cudaMemsetAsync(hist, 0, 8 * 256 * sizeof(uint32_t), s.stream));
  int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  FSPH_LAUNCH(k_hist_all, nb, 256, s.stream, k0, n, hist);
  std::vector<uint32_t> h(8 * 256);
  FSPH_CUDA_OK(cudaMemcpyAsync(h.data(), hist, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
  FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
  auto trivial = [&](int p) {
    for(int d = 0; d < 256; ++d)
      if(h[p * 256 + d] == (uint32_t)n) return true;
    return false;
  };
  for(int attempt = 0; attempt < 2; ++attempt) {
    const int p_lo = (attempt == 0 && keys_src) ? 4 : 0;
    bool first = true;
    for(int p = p_lo; p < 8; ++p) {
      if(trivial(p)) continue;
      radix_pass_mode(s, k0, v0, k1, v1, n, 8 * p, first ? 2 : 1);
      first = false;
      std::swap(k0, k1);
      std::swap(v0, v1);
    }
    if(first) {
      // all passes trivial: identity permutation
      radix_pass_mode(s, k0, v0, k1, v1, n, 0, 2);
      std::swap(k0, k1);
      std::swap(v0, v1);
    }
    if(p_lo == 0) return;
    uint32_t * flag = s.b_status.as<uint32_t>() + 2;
    FSPH_CUDA_OK(cudaMemsetAsync(flag, 0, 4, s.stream));
    FSPH_LAUNCH(k_fix_runs, (unsigned)((n + 255) / 256), 256, s.stream, k0, v0, n, flag);
    uint32_t f = 0;
    FSPH_CUDA_OK(cudaMemcpyAsync(&f, flag, 4, cudaMemcpyDeviceToHost, s.stream));
    FSPH_CUDA_OK(cudaStreamSynchronize(s.stream));
    if(!f) return;
    // a long run of equal high halves (bodies clustered below 2^-10 of the range): start again from the caller's keys
    FSPH_CUDA_OK(cudaMemcpyAsync(k0, keys_src, (size_t)n * 8, cudaMemcpyDeviceToDevice, s.stream));
  }
}

}  // namespace fsph
