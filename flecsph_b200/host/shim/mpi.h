/* Single-process MPI stand-in for hosts WITHOUT an MPI installation (this image has none): lets the unmodified
 * reference drivers and headers (/root/reference) compile -- against the GPU facade (flecsph_b200/host/bodies_system.h)
 * for the product drivers, and against the reference's own body_system for the test oracle (oracle/_ref).  Rank 0 of 1:
 * collectives copy send->recv, point-to-point calls abort (never reached with one rank).
 * With a real MPI, drop this directory from the include path. */
#ifndef FSPH_SHIM_MPI_H
#define FSPH_SHIM_MPI_H
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; };

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 8
#define MPI_INT64_T 9
#define MPI_UNSIGNED_LONG 10
#define MPI_LONG_LONG 11
#define MPI_C_BOOL 12
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_THREAD_MULTIPLE 3
#define MPI_INFO_NULL 0
#define MPI_REQUEST_NULL 0

static inline size_t fsph_mpi_size_(MPI_Datatype t) {
  switch(t) {
    case 1: return 1; case 4: return 4; case 5: return 4; case 8: return 8;
    case 9: return 8; case 10: return 8; case 11: return 8; case 12: return 1;
    default: return t >= 1000 ? (size_t)(t - 1000) : 1;
  }
}
static inline int fsph_mpi_die_(const char * w) {
  std::fprintf(stderr, "mpi shim: %s called on a single rank\n", w);
  std::abort();
  return 1;
}
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Init_thread(int *, char ***, int req, int * prov) { if(prov) *prov = req; return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Abort(MPI_Comm, int c) { std::fprintf(stderr, "MPI_Abort(%d)\n", c); std::abort(); return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int * r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int * s) { *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void * s, void * r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy(r, s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Reduce(const void * s, void * r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy(r, s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Allgather(const void * s, int n, MPI_Datatype t, void * r, int, MPI_Datatype, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy(r, s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Gather(const void * s, int n, MPI_Datatype t, void * r, int, MPI_Datatype, int, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy(r, s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Alltoall(const void * s, int n, MPI_Datatype t, void * r, int, MPI_Datatype, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy(r, s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Allgatherv(const void * s, int n, MPI_Datatype t, void * r, const int *, const int * displs, MPI_Datatype, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy((char *)r + (displs ? displs[0] : 0) * fsph_mpi_size_(t), s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Gatherv(const void * s, int n, MPI_Datatype t, void * r, const int *, const int * displs, MPI_Datatype, int, MPI_Comm) {
  if(s != MPI_IN_PLACE) std::memcpy((char *)r + (displs ? displs[0] : 0) * fsph_mpi_size_(t), s, n * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Alltoallv(const void * s, const int * sc, const int * sd, MPI_Datatype t, void * r, const int *, const int * rd, MPI_Datatype, MPI_Comm) {
  std::memcpy((char *)r + rd[0] * fsph_mpi_size_(t), (const char *)s + sd[0] * fsph_mpi_size_(t), sc[0] * fsph_mpi_size_(t));
  return 0;
}
static inline int MPI_Iprobe(int, int, MPI_Comm, int * flag, MPI_Status *) { *flag = 0; return 0; }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { return fsph_mpi_die_("MPI_Probe"); }
static inline int MPI_Test(MPI_Request *, int * flag, MPI_Status *) { *flag = 1; return 0; }
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { return fsph_mpi_die_("MPI_Get_count"); }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return fsph_mpi_die_("MPI_Send"); }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return fsph_mpi_die_("MPI_Recv"); }
static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return fsph_mpi_die_("MPI_Isend"); }
static inline int MPI_Issend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return fsph_mpi_die_("MPI_Issend"); }
static inline int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return fsph_mpi_die_("MPI_Sendrecv"); }
static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype * nt) { *nt = 1000 + (int)(n * fsph_mpi_size_(t)); return 0; }
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }
#endif
