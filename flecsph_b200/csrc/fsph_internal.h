// fsph_internal.h -- internal declarations of libflecsph_b200.so (sm_100a CUDA, no CPU path).
//
// Device state of one `fsph_system` (= one body_system<double,D> of the reference,
// mpisph/bodies_system.h:38-379): fp64 / int SoA columns in HBM, the level-ordered octree
// tables, the work-cell list and the per-work-cell neighbour structure (candidate list +
// bitmask).  DESIGN.md describes the layout and each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <atomic>
#include <vector>
#include <utility>
#include <stdexcept>
#include "flecsph_b200.h"
#include "fsph_eforce.h"

namespace fsph {

// ---- fp64 columns -----------------------------------------------------------------------------
enum F64 {
  F_POS = 0, F_VEL = 3, F_VELH = 6, F_ACC = 9, F_GACC = 12,
  F_MASS = 15, F_H, F_RHO, F_P, F_U, F_E, F_CS, F_DUDT, F_DEDT, F_DT, F_ALPHA, F_VSIG, F_GPOT, F_ADIA,
  F_CSOLD,  // soundspeed before the current density sweep (signal-speed order dependence)
  F_PRHO2,  // P / rho^2, derived right before each pair sweep that needs it (not permuted by the sort)
  // optional columns (allocated only for the physics that uses them, System::col_used):
  F_YE,                                  // electron fraction: white-dwarf EOS (eos/eos.h:225-315)
  F_DIVV, F_DDIVVDT,                     // Cullen-Dehnen viscosity state (default_physics.h:273-300)
  F_TRIGGER, F_XI, F_TRACESS, F_GRADV,   // ... and the diagnostics viscosity::compute_alpha stores (viscosity.h:193-363)
  NF64
};
constexpr int F_FIRST_OPTIONAL = F_YE;

constexpr int SUB_ENTITIES = 128;  // tree_topology.h:2121
#ifndef FSPH_SUPER_ENTITIES
#define FSPH_SUPER_ENTITIES 4096
#endif
constexpr int SUPER_ENTITIES = FSPH_SUPER_ENTITIES;  // bodies per super cell of the neighbour search
constexpr int MAX_LEVELS = 66;

// error bits latched on the device (surfaced at the next call boundary)
enum ErrBits {
  ERR_DUP_KEYS = 1, ERR_RHO_NONPOS = 2, ERR_U_NONPOS = 4, ERR_NAN = 8, ERR_NEGATIVE = 16, ERR_QUEUE_OVERFLOW = 32,
  ERR_DT_LOOP = 64, ERR_RANGE = 128, ERR_SEARCH_CAPACITY = 256, ERR_DD_CAPACITY = 512
};

struct DevBuf {
  void * p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes);  // grows with 12 % slack (contents NOT preserved)
  void reserve_roomy(size_t bytes) { if(bytes > cap) reserve(bytes + bytes / 2); }  // halo-sized buffers: their sizes move by several
                                                                                    // per cent from step to step, and cudaFree + cudaMalloc cost ~40 ms
  void reserve_exact(size_t bytes);  // grows to exactly `bytes` (for buffers that swap roles: capacities converge)
  void grow_keep(size_t bytes, size_t keep, cudaStream_t st);  // grows, first `keep` bytes preserved
  void release();
  template<typename T> T * as() const { return static_cast<T *>(p); }
};

// pointers handed to kernels by value
struct Cols {
  double * f[NF64];
  int64_t * id;
  uint64_t * key;
  int32_t * type;
  int32_t * nneigh;
  // Owned bodies occupy the index range [own_lo, own_hi) (everything when one rank owns all): with a halo the storage
  // order is [ghosts of lower ranks | owned | ghosts of higher ranks], so element-wise passes and reductions simply run
  // on the sub-range (a per-body flag test in those streaming kernels costs 50 % of their time).
  int64_t own_lo, own_hi;
};

constexpr int32_t TYPE_GHOST = 0x100;  // flag in the type column: copy of a body owned by another rank (tree_types.h:365 "locality")

// packed records for the neighbour walk: one 64-byte load per frontier node, one 32-byte load per leaf child
struct __align__(16) NodeRec {
  double bmin[3], bmax[3];
  int32_t child_begin;       // first child cell
  int32_t first_child_node;  // node index of the first node child (node children are consecutive nodes)
  uint32_t nchild;           // number of child cells (<= 2^D)
  uint32_t leafmask;         // bit j: child cell child_begin + j is a leaf
};
struct __align__(16) LeafRec {
  double x[3], h;
};

struct Tree {
  NodeRec * noderec;   // [nnodes]
  LeafRec * leafrec;   // [ncells] (entries of node cells unused)
  // cells in level order (depth, then key): value >= 0 -> particle index (leaf), < 0 -> ~node index
  int32_t * cell;
  int32_t * cell_parent;  // node index of the parent (-1 for the root cell)
  int32_t * node_parent;  // [nnodes] parent node (-1 for the root)
  uint64_t * cell_desc;   // depth << 32 | start << 1 | is_leaf
  // nodes in level order
  int32_t * child_begin;
  int32_t * child_end;
  int32_t * node_start;
  int32_t * node_depth;
  int32_t * sub;
  double * mass;
  double * com[3];
  double * radius;
  double * lap;
  double * bmin[3];
  double * bmax[3];
  int32_t ncells, nnodes;
};

// neighbour structure of one step (fsph_search.cu): blocks of <= BLK_BODIES consecutive bodies (<= BLK_CELLS whole work
// cells of one super cell), per block a union candidate list, per (block, tile of NL_TILE candidates, warp) a
// neighbour list stored as rows of 32 x uint16
constexpr int BLK_BODIES = 256;   // bodies per block = threads per CTA of the search and the sweeps
constexpr int BLK_CELLS = 16;     // work cells per block (bits of a candidate's cell mask)
constexpr int NL_TILE = 512;      // candidates per list tile (tile-local positions fit 16 bits)
// A block's candidates (walk order) are cut into chunks of NL_CHUNK; every chunk is sorted by micro-cell and dealt to
// its <= NL_CHUNK_TILES tiles round-robin (fsph_search.cu).  Tile ti of a block with C candidates covers the candidate
// positions [tb, tb + tc) of the block:
constexpr int NL_CHUNK_TILES = 3;
constexpr int NL_CHUNK = NL_CHUNK_TILES * NL_TILE;
__host__ __device__ inline int nl_ntiles(int C) { return (C / NL_CHUNK) * NL_CHUNK_TILES + ((C % NL_CHUNK) + NL_TILE - 1) / NL_TILE; }
__host__ __device__ inline void nl_tile_range(int C, int ti, int & tb, int & tc) {
  const int ch = ti / NL_CHUNK_TILES, tj = ti - ch * NL_CHUNK_TILES;
  const int base = ch * NL_CHUNK;
  const int n = C - base < NL_CHUNK ? C - base : NL_CHUNK;
  const int nt = (n + NL_TILE - 1) / NL_TILE;
  const int q = n / nt, r = n % nt;
  tb = base + tj * q + (tj < r ? tj : r);
  tc = q + (tj < r ? 1 : 0);
}
struct __align__(16) TileDir {
  uint32_t row[8];     // per warp: its first row of the tile in the list arena
  uint32_t cmax[8];    // per warp: the longest list of its lanes in this tile (the warp owns 1 + cmax rows)
};
// one entry of a super cell's node list as the per-block filter reads it
struct __align__(16) SlRec {
  float mn[3], mx[3];
  int32_t node, pad;
};
struct Search {
  int32_t * wc_cell;     // work cells in Morton order: index into Tree::cell
  int32_t nwc;
  // super cells = maximal sub-trees of <= SUPER_ENTITIES bodies; each owns the list of tree nodes with leaf children
  // that any of its work cells can accept (superset by a conservative box test)
  int32_t * wc_super;    // [nwc] super cell of each work cell
  int32_t * super_cell;  // [nsuper] index into Tree::cell
  uint64_t * sl_off;     // [nsuper] offset of each super cell's list in sl_nodes / sl_rec
  int32_t * sl_len;      // [nsuper] its length
  int32_t * sl_nodes;    // node indices
  struct SlRec * sl_rec; // the same nodes with their boxes in fp32 (rounded outwards): what the per-block filter reads
  int32_t nsuper;
  // blocks
  int32_t nblk;
  int32_t * blk_wc;        // [nblk+1] first work cell of each block
  uint64_t * blk_cand_off; // [nblk] first candidate of the block in cand
  uint64_t * blk_dir_off;  // [nblk] first tile of the block in dir
  int32_t * blk_ncand;     // [nblk] candidates of the block
  int32_t * blk_body;      // [nblk] first body
  int32_t * blk_nbody;     // [nblk] bodies (thread k of the block's CTA <-> body blk_body + k)
  float * blk_hmax;        // [nblk] largest h among the block's candidates (rounded up)
  int32_t * cand;          // candidate particle indices, block after block, tile after tile
  TileDir * dir;           // one entry per (block, tile)
  uint16_t * nl;           // list arena: rows of 32 uint16; per (tile, warp): row 0 = the lanes' counts, row 1 + j = the
                           // tile-local position of each lane's j-th neighbour
  uint64_t nrows;          // rows in use
};

struct PhysParams {
  int kernel, thermokinetic, evolve_u, eos, visc;
  double eta, gamma, alpha, beta, epsilon, cfl, kernel_width;
  double dt;
  int64_t iteration;
  int relaxation_steps;
  double alpha_max, visc_l, visc_delta;                      // Cullen-Dehnen viscosity
  double gamma2, rho_thr;                                    // piecewise polytrope
  double relax_beta, relax_gamma, repulsion_radius, repulsion_gamma;   // relaxation phase
  int fuse_kick_v;   // the acceleration sweep of fsph_step also applies the second velocity kick (integration.h:71-79)
  int ext_potential; // external potential present: u is recovered from e by an element-wise pass before the density sweep
};

// eos::compute_pressure + compute_soundspeed (include/physics/eos/eos.h): ideal gas :166-183 from the specific internal
// energy, polytropic :100-118 from the particle's adiabatic invariant K = P / rho^gamma (eos::init, :68-83), piecewise
// polytrope :365-394 (K = K1 of the first segment), cold white dwarf :254-275 from the electron fraction
__host__ __device__ inline void eos_p_cs(const PhysParams & pp, double rho, double u, double K, double ye, double & P, double & cs) {
  if(pp.eos == FSPH_EOS_POLYTROPIC) {
    P = K * pow(rho, pp.gamma);
    cs = sqrt(K * pp.gamma * pow(rho, pp.gamma - 1.));
  }
  else if(pp.eos == FSPH_EOS_PPT) {
    if(rho < pp.rho_thr) {
      P = K * pow(rho, pp.gamma);
      cs = sqrt(K * pp.gamma * pow(rho, pp.gamma - 1.));
    }
    else {
      const double K2 = K * pow(pp.rho_thr, pp.gamma - pp.gamma2);
      P = K2 * pow(rho, pp.gamma2);
      cs = sqrt(K2 * pp.gamma2 * pow(rho, pp.gamma2 - 1.));
    }
  }
  else if(pp.eos == FSPH_EOS_WD) {
    const double A_wd = 6.00233181e22, B_wd_nm = 9.73932099e5, C_LIGHT_CGS = 2.99792458e10;
    const double x = cbrt(rho * ye / B_wd_nm), x2 = x * x;
    P = A_wd * (x * (2. * x2 - 3.) * sqrt(x2 + 1.) + 3. * asinh(x));
    const double numer = (1. + x2) * (6. * x2 - 3.) + 3. + x2 * (2. * x2 - 3.);
    const double denom = (1. + x2) * (6. * x2 + 1.) - 1. + x2 * (2. * x2 + 1.);
    cs = sqrt(numer / (3. * denom)) * C_LIGHT_CGS;
  }
  else if(pp.eos == FSPH_EOS_IDEAL) {
    P = (pp.gamma - 1.) * rho * u;
    cs = sqrt(pp.gamma * (pp.gamma - 1.) * u);
  }
  // FSPH_EOS_NONE: pressure and sound speed are left as they are (eos.h:425-432)
}

struct Phase {
  std::string name;
  cudaEvent_t e0, e1;
};

struct Comm;
struct Dist;

struct System {
  fsph_params par;
  ExtForce ef;   // external_force_type parsed (fsph_eforce.h)
  // domain decomposition (fsph_dist.cu); comm == nullptr: one rank owns everything
  Comm * comm = nullptr;
  Dist * dist = nullptr;
  int64_t n_owned = 0;     // bodies [0, n) hold n_owned owned bodies and n - n_owned ghosts once the halo is in
  bool has_ghosts = false;
  int64_t own_begin = 0;   // index of the first owned body while has_ghosts
  DevBuf col_gd;           // int32: depth of the body's leaf in the global tree (multi-rank only)
  const uint32_t * last_perm = nullptr;  // permutation of the last sort_bodies (valid until the next sort / tree build)
  int dim = 3, device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t n = 0;
  // time globals (default_physics.h:31-40)
  double dt = 0, dt_saved = 0, totaltime = 0, totaltime_next = 0, t_out[3] = {0, 0, 0};
  int64_t iteration = 0;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool tree_valid = false, search_valid = false, tree_dirty = false;
  // inside fsph_step the sort need not permute columns that are recomputed before their next read
  bool skip_dead_columns = false;
  bool in_step = false;   // inside fsph_step: the order of the sweeps and ghost refreshes is the driver's
  bool fuse_kick_v = false;   // ... and the acceleration sweep applies leapfrog_kick_v to the bodies it finishes
  bool dead_col[NF64] = {};
  bool col_used[NF64] = {};   // optional columns this system allocates (fsph_create)

  // storage
  DevBuf col[NF64], col_id, col_key, col_type, col_nneigh, scratch_col;
  DevBuf b_perm, b_keys2, b_perm2, b_tmp[6];
  DevBuf b_gather[7];   // spare column buffers of the fused gather (fsph_tree.cu)
  DevBuf b_node_parent;
  DevBuf b_fmm_nodes, b_fmm_out, b_fmm_stack;   // FMM: node accumulators, output strip and stack of pending cell pairs
  unsigned long long fmm_pairs = 0;             // cell pairs the last gravitation_fmm processed
  DevBuf b_cell, b_cell_parent, b_cell_desc, b_cell_desc2, b_child_begin, b_child_end, b_node_start, b_node_depth, b_sub;
  DevBuf b_noderec, b_leafrec;
  DevBuf b_nmass, b_ncom[3], b_nradius, b_nlap, b_nbmin[3], b_nbmax[3];
  DevBuf b_wc_super, b_super_cell, b_sl_off, b_sl_len, b_sl_nodes, b_sl_rec, b_sl_stage;
  uint64_t est_sl = 0;   // entries of the previous step's node lists (capacity of the single-pass list build)
  DevBuf b_wc_cell, b_cand, b_cnt0, b_cnt1;
  DevBuf b_status, b_red, b_levels, b_search_scratch;
  DevBuf b_blk_wc, b_blk_meta, b_dir, b_nl, b_super_queue, b_wcinfo, b_wcrep, b_wcand;
  DevBuf b_pair_a, b_pair_pi;         // per-pair caches of the sweeps, parallel to the list arena (fsph_sweeps.cu)
  unsigned long long est_cand = 0, est_dir = 0, est_rows = 0;  // sizes of the previous neighbour structure (capacity estimate)
  int64_t est_n = 0;
  int search_scratch_cap = 8192;      // candidates per warp strip of the walk (grows on demand)
  bool search_tiny_done = false;      // test hook FSPH_SEARCH_TINY_CAP used up
  uint64_t search_epoch = 0;          // bumped by every neighbour search (validity of the pair caches)
  // pair caches (fsph_sweeps.cu): decided per neighbour structure; A is valid for the structure it was stored with,
  // Pi additionally only while rho, c_s and v_1/2 are unchanged
  uint64_t pair_cache_epoch = ~0ull, pair_a_epoch = 0;
  bool pair_cache_on = false, pair_pi_valid = false;
  void * h_pinned = nullptr;  // small pinned staging (4 KB)
  Tree tree;
  Search search;
  int level_node_begin[MAX_LEVELS + 1];
  int nlevels = 0;
  double kernel_width = 2.0;
  uint32_t host_err = 0;

  // fsph_step_streamed: results copied out on a side stream as soon as they are final
  struct StreamOut { int col, comp; void * host; bool done; cudaEvent_t ev; };
  std::vector<StreamOut> stream_out;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_event = nullptr;
  // fsph_step_pipelined: device-to-host copies of the previous step still in flight when the call returned, the stream
  // the next step's inputs arrive on, and a pool of (untimed) events, one per copy in flight
  std::vector<StreamOut> prev_out;
  cudaStream_t up_stream = nullptr;
  cudaEvent_t up_event = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  bool in_pipelined_call = false;

  // phase timing of the last fsph_step
  std::vector<Phase> phases;
  std::vector<double> phase_ms;
  std::string phase_names;
  bool timing = false;

  Cols cols() const;
  PhysParams phys() const;
  int64_t n_owned_or_n() const { return has_ghosts ? n_owned : n; }
};

// CUDA-event bracket of one phase of fsph_step (no-op unless the step is being timed)
struct PhaseScope {
  System & s;
  PhaseScope(System & sys, const char * name) : s(sys) {
    if(!s.timing) return;
    Phase p;
    p.name = name;
    cudaEventCreate(&p.e0);
    cudaEventCreate(&p.e1);
    cudaEventRecord(p.e0, s.stream);
    s.phases.push_back(p);
    idx = s.phases.size() - 1;
  }
  ~PhaseScope() {
    if(!s.timing) return;
    cudaEventRecord(s.phases[idx].e1, s.stream);
  }
  size_t idx = 0;
};

#define FSPH_CUDA_OK(call)                                                                       \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if(e_ != cudaSuccess) {                                                                      \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));              \
    }                                                                                            \
  } while(0)

// every kernel launch of the library goes through this macro (bench.py reports the count as gpu_launches)
extern std::atomic<unsigned long long> g_launches;
#define FSPH_LAUNCH(kernel, grid, block, stream, ...)            \
  do {                                                           \
    ++fsph::g_launches;                                          \
    kernel<<<grid, block, 0, stream>>>(__VA_ARGS__);             \
  } while(0)

// launch with dynamic shared memory (the opt-in above 48 KB is per device: set before every launch, it is cheap)
#define FSPH_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                            \
  do {                                                                                                      \
    FSPH_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));   \
    ++fsph::g_launches;                                                                                     \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                                                     \
  } while(0)

// ---- fsph_sort.cu ---------------------------------------------------------------------------
// One stable LSD pass on an 8-bit digit at `shift`; vals may be null (keys only).
void radix_pass(System & s, const uint64_t * kin, const uint32_t * vin, uint64_t * kout, uint32_t * vout, int64_t n, int shift);
// Full 64-bit sort of (key, index) pairs; result in kout/vout (may swap internally): returns pointers to the sorted arrays.
// keys_src: the caller's copy of the unsorted keys (enables the four-pass + run fix-up path; null: eight passes)
void radix_sort_pairs(System & s, uint64_t *& k0, uint32_t *& v0, uint64_t *& k1, uint32_t *& v1, int64_t n, const uint64_t * keys_src = nullptr);
void exclusive_scan_u32(System & s, const uint32_t * in, uint32_t * out, int64_t n, uint32_t * total_dev);  // out[n] = total (callers rely on it)
void exclusive_scan_u64(System & s, const uint64_t * in, uint64_t * out, int64_t n, uint64_t * total_dev);

// ---- fsph_tree.cu ---------------------------------------------------------------------------
void compute_range(System & s);                 // bodies_system.h:231-268
void compute_keys(System & s);                  // filling_curve.h:503-525
// psort.h:227-322 (one rank): radix sort + column gather; only the first n_keep bodies of the sorted order are kept
void sort_bodies(System & s, int64_t n_keep = -1);
void live_columns(System & s, std::vector<std::pair<DevBuf *, int>> & out);  // columns that carry state, element size
void halo_columns(System & s, std::vector<std::pair<DevBuf *, int>> & out);  // columns a remote copy arrives with inside fsph_step
void gather_bodies(System & s, const uint32_t * perm, int64_t n_new);  // every live column: out[i] = in[perm[i]], i < n_new
void build_tree(System & s);                    // tree_topology.h:878-997 + default_physics.h:59-119
// Multi-rank variants: the local tree is the global canonical tree restricted to the bodies present here.
struct TreeEdges {
  bool has_before = false, has_after = false;   // last key of the previous rank / first key of the next rank
  uint64_t key_before = 0, key_after = 0;
  const int32_t * gd_in = nullptr;              // per body: leaf depth in the global tree (given for the merged tree)
  int32_t * gd_out = nullptr;                   // per body: leaf depth computed here (owned tree)
  bool finish = true;                           // pack the walk records and collect the work cells
};
void build_tree_edges(System & s, const TreeEdges & e);
void finish_tree(System & s);                   // walk records + work cells (after node-table overrides)
void find_work_cells(System & s);               // tree_topology.h:221-235
// cells that are maximal sub-trees of <= `limit` bodies (or lone leaves under bigger nodes), marked at their first body:
// at[first body] = cell index + 1.  skip_ghosts: cells whose first body is a remote copy are left out.
void mark_maximal_cells(System & s, int limit, bool skip_ghosts, uint32_t * at);

// ---- fsph_search.cu -------------------------------------------------------------------------
void neighbor_search(System & s);               // tree_topology.h:212-442
void export_neighbors(System & s, std::vector<int64_t> & off, std::vector<int64_t> * ids);

// ---- fsph_sweeps.cu -------------------------------------------------------------------------
void sweep(System & s, int op);                 // default_physics.h:193-566

// ---- fsph_boundary.cu -----------------------------------------------------------------------
void periodic_walls(System & s);                // boundary.h:45-248 as used by bodies_system.h:160-169

// ---- fsph_fmm.cu ----------------------------------------------------------------------------
void gravitation_fmm(System & s);               // bodies_system.h:282 -> tree_topology.h:448-795 + fmm.h
// the same walk restricted to sink cells that hold bodies of the index range [own_lo, own_hi) (a rank's share)
void gravitation_fmm_range(System & s, int64_t own_lo, int64_t own_hi);
void release_buffers(System & s);               // every device buffer of a system

// ---- fsph_apply.cu --------------------------------------------------------------------------
void apply(System & s, int op);                  // integration.h, default_physics.h element-wise
void apply_fused(System & s, const int * ops, int nops);  // several of them in one pass (same per-particle order)
void reduce(System & s, int op, double * out, int nout);
double max_h(System & s);                        // over the bodies of this rank
void global_combine(System & s, int kind, double v[3]);  // combine a reduction across ranks (rank order: deterministic)
void check_status(System & s);                  // throws on latched device errors

double * reduce_scratch(System & s, size_t n);

// ---- fsph_comm.cu / fsph_dist.cu --------------------------------------------------------------
// Collectives on device buffers, stream-ordered.  Two backends: NCCL (one process per GPU, libnccl.so.2 resolved at
// run time) and an in-process group of host threads (several systems in one process; used by the 1-GPU tests).
struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  virtual void abort() {}  // this rank failed: wake peers blocked in a collective (MPI_Abort of mpisph/utils.h:34-38)
  virtual void all_gather(const void * send, void * recv, size_t bytes, cudaStream_t st) = 0;
  // block for rank q: send + soff[q] (scount[q] bytes); the block from rank q lands at recv + roff[q] (rcount[q] bytes)
  virtual void all_to_all_v(const void * send, const size_t * soff, const size_t * scount, void * recv, const size_t * roff,
    const size_t * rcount, cudaStream_t st) = 0;
};
Comm * make_nccl_comm(const char * unique_id, int rank, int nranks);
void nccl_unique_id(char * out128);
struct LocalGroup;
LocalGroup * local_group_create(int nranks);
void local_group_destroy(LocalGroup * g);
Comm * make_local_comm(LocalGroup * g, int rank);

void dd_choose_splitters(int nranks, int nsamples, const uint64_t * keys, const int64_t * counts, uint64_t * out);
uint64_t dd_align_splitter(int D, std::vector<uint64_t> & below, std::vector<uint64_t> & above, uint64_t s);
void dist_all_gather_host(System & s, const void * in, void * out, size_t bytes);  // small host values, staged through the device
void dist_update_iteration(System & s);          // bodies_system.h:152-229 with size > 1 (psort + tree exchange)
void dist_refresh_ghosts(System & s);            // bodies_system.h:273 reset_ghosts with size > 1
void dist_drop_ghosts(System & s);
void dist_gravitation_fmm(System & s);           // gravitation_fmm with size > 1
void dist_release(System & s);

}  // namespace fsph
