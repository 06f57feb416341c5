// Stand-in for <omp.h>: the reference only uses omp_get_wtime() for its
// trace-level timers (mpisph/bodies_system.h:188, tree_topology.h:881); the
// image has no libgomp.
#pragma once
#include <chrono>
static inline double omp_get_wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline int omp_get_max_threads() { return 1; }
static inline int omp_get_thread_num() { return 0; }
static inline int omp_get_num_threads() { return 1; }
