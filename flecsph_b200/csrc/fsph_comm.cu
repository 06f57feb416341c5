// fsph_comm.cu -- the two collectives the domain decomposition needs (all-gather, all-to-all-v) on device buffers.
//
// Replaces the MPI calls of the reference's multi-rank path ("every collective / message call
// site"): MPI_Allreduce / MPI_Allgather (bodies_system.h:253-262, psort.h:258-301, tree_topology.h:1976-2003),
// MPI_Alltoall(v) (psort.h:303-308, mpi_utils.h:60-113) and the request / reply traffic of traversal_sph
// (tree_topology.h:1276-1542).  No MPI anywhere:
//   * NcclComm   one process per GPU; libnccl.so.2 is resolved with dlopen at run time (the copy torch loaded, when the
//                process imported torch first), so the library links against nothing but the CUDA runtime;
//   * LocalComm  several fsph_systems of ONE process, each driven by its own host thread, exchange through plain
//                device-to-device copies (same GPU or peers).  It exists so that the whole decomposition logic is
//                exercised by the 1-GPU test suite; the data path is identical above this interface.
#include <dlfcn.h>
#include <cstring>
#include <string>
#include <vector>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <chrono>
#include "fsph_internal.h"

namespace fsph {

// ---- NCCL through dlopen ---------------------------------------------------------------------------------
namespace {

typedef struct ncclComm * ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
enum { ncclSuccess_ = 0, ncclChar_ = 0 };

struct NcclApi {
  void * h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char * (*GetErrorString)(int) = nullptr;
};

NcclApi & nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char * names[] = {"libnccl.so.2", "libnccl.so"};
    for(const char * nm : names) {
      api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if(api.h) break;
    }
    if(!api.h) return;
#define FSPH_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name))
    FSPH_SYM(GetUniqueId, "ncclGetUniqueId");
    FSPH_SYM(CommInitRank, "ncclCommInitRank");
    FSPH_SYM(CommDestroy, "ncclCommDestroy");
    FSPH_SYM(CommAbort, "ncclCommAbort");
    FSPH_SYM(GroupStart, "ncclGroupStart");
    FSPH_SYM(GroupEnd, "ncclGroupEnd");
    FSPH_SYM(Send, "ncclSend");
    FSPH_SYM(Recv, "ncclRecv");
    FSPH_SYM(AllGather, "ncclAllGather");
    FSPH_SYM(GetErrorString, "ncclGetErrorString");
#undef FSPH_SYM
  });
  if(!api.h || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllGather || !api.GroupStart || !api.GroupEnd)
    throw std::runtime_error("libnccl.so.2 not found (multi-GPU runs need NCCL; there is no MPI / host fallback)");
  return api;
}

void nccl_ok(int rc, const char * what) {
  if(rc != ncclSuccess_) {
    NcclApi & a = nccl();
    throw std::runtime_error(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
  }
}

struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  ~NcclComm() override {
    if(comm) nccl().CommDestroy(comm);
  }
  // This rank failed (device guard, capacity, allocation): tear the communicator down so that the peers' pending
  // collectives return an error instead of blocking for ever -- the reference's MPI_Abort (mpisph/utils.h:34-38).
  void abort() override {
    if(comm && nccl().CommAbort) {
      nccl().CommAbort(comm);
      comm = nullptr;
    }
  }
  void all_gather(const void * send, void * recv, size_t bytes, cudaStream_t st) override {
    nccl_ok(nccl().AllGather(send, recv, bytes, ncclChar_, comm, st), "ncclAllGather");
  }
  void all_to_all_v(const void * send, const size_t * soff, const size_t * scount, void * recv, const size_t * roff,
    const size_t * rcount, cudaStream_t st) override {
    NcclApi & a = nccl();
    const char * sp = static_cast<const char *>(send);
    char * rp = static_cast<char *>(recv);
    if(scount[rank] != rcount[rank]) throw std::runtime_error("all_to_all_v: self block size mismatch");
    if(scount[rank]) FSPH_CUDA_OK(cudaMemcpyAsync(rp + roff[rank], sp + soff[rank], scount[rank], cudaMemcpyDeviceToDevice, st));
    nccl_ok(a.GroupStart(), "ncclGroupStart");
    for(int q = 0; q < nranks; ++q) {
      if(q == rank) continue;
      if(scount[q]) nccl_ok(a.Send(sp + soff[q], scount[q], ncclChar_, q, comm, st), "ncclSend");
      if(rcount[q]) nccl_ok(a.Recv(rp + roff[q], rcount[q], ncclChar_, q, comm, st), "ncclRecv");
    }
    nccl_ok(a.GroupEnd(), "ncclGroupEnd");
  }
};

}  // namespace

void nccl_unique_id(char * out128) {
  ncclUniqueId_ id;
  nccl_ok(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, id.internal, 128);
}

Comm * make_nccl_comm(const char * unique_id, int rank, int nranks) {
  ncclUniqueId_ id;
  memcpy(id.internal, unique_id, 128);
  NcclComm * c = new NcclComm;
  c->rank = rank;
  c->nranks = nranks;
  int rc = nccl().CommInitRank(&c->comm, nranks, id, rank);
  if(rc != ncclSuccess_) {
    c->comm = nullptr;
    delete c;
    nccl_ok(rc, "ncclCommInitRank");
  }
  return c;
}

// ---- in-process group ------------------------------------------------------------------------------------
struct LocalGroup {
  int n = 0;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long long generation = 0;
  bool broken = false;
  struct Slot {
    const void * send = nullptr;
    const size_t * soff = nullptr;
    const size_t * scount = nullptr;
  };
  std::vector<Slot> slot;

  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    if(broken) throw std::runtime_error("local group: a peer rank failed");
    const unsigned long long gen = generation;
    if(++waiting == n) {
      waiting = 0;
      ++generation;
      cv.notify_all();
      return;
    }
    if(!cv.wait_for(lk, std::chrono::seconds(120), [&]() { return generation != gen || broken; })) {
      broken = true;
      cv.notify_all();
      throw std::runtime_error("local group: barrier timed out (a peer rank stopped issuing collectives)");
    }
    if(broken) throw std::runtime_error("local group: a peer rank failed");
  }
};

LocalGroup * local_group_create(int nranks) {
  LocalGroup * g = new LocalGroup;
  g->n = nranks;
  g->slot.resize(nranks);
  return g;
}
void local_group_destroy(LocalGroup * g) { delete g; }

namespace {
struct LocalComm : Comm {
  LocalGroup * g = nullptr;
  void abort() override {
    std::unique_lock<std::mutex> lk(g->m);
    g->broken = true;
    g->cv.notify_all();
  }
  void all_gather(const void * send, void * recv, size_t bytes, cudaStream_t st) override {
    FSPH_CUDA_OK(cudaStreamSynchronize(st));  // my contribution is complete
    g->slot[rank].send = send;
    g->barrier();
    char * rp = static_cast<char *>(recv);
    for(int q = 0; q < nranks; ++q)
      FSPH_CUDA_OK(cudaMemcpyAsync(rp + (size_t)q * bytes, g->slot[q].send, bytes, cudaMemcpyDefault, st));
    FSPH_CUDA_OK(cudaStreamSynchronize(st));
    g->barrier();  // nobody reuses a send buffer before every peer has read it
  }
  void all_to_all_v(const void * send, const size_t * soff, const size_t * scount, void * recv, const size_t * roff,
    const size_t * rcount, cudaStream_t st) override {
    FSPH_CUDA_OK(cudaStreamSynchronize(st));
    g->slot[rank].send = send;
    g->slot[rank].soff = soff;
    g->slot[rank].scount = scount;
    g->barrier();
    char * rp = static_cast<char *>(recv);
    for(int q = 0; q < nranks; ++q) {
      const LocalGroup::Slot & sl = g->slot[q];
      if(sl.scount[rank] != rcount[q]) {
        std::unique_lock<std::mutex> lk(g->m);
        g->broken = true;
        g->cv.notify_all();
        throw std::runtime_error("all_to_all_v: send / receive counts disagree");
      }
      if(rcount[q])
        FSPH_CUDA_OK(cudaMemcpyAsync(rp + roff[q], static_cast<const char *>(sl.send) + sl.soff[rank], rcount[q], cudaMemcpyDefault, st));
    }
    FSPH_CUDA_OK(cudaStreamSynchronize(st));
    g->barrier();
  }
};
}  // namespace

Comm * make_local_comm(LocalGroup * g, int rank) {
  if(!g || rank < 0 || rank >= g->n) throw std::runtime_error("bad local group / rank");
  LocalComm * c = new LocalComm;
  c->g = g;
  c->rank = rank;
  c->nranks = g->n;
  return c;
}

}  // namespace fsph
