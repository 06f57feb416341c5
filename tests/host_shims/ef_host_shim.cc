// TEST INFRASTRUCTURE: evaluates the __host__ __device__ force / potential terms of flecsph_b200/csrc/fsph_eforce.h on the
// host so that tests/test_eforce.py can compare the header's arithmetic with the reference's eforce.h (through
// oracle/_ref) without a GPU.  Built by the test with g++; never part of libflecsph_b200.so.
#define __host__
#define __device__
#include "fsph_eforce.h"

template<int D>
static void eval(const fsph::ExtForce & ef, int n, const double * x, double * acc, double * pot) {
  for(int i = 0; i < n; ++i) {
    fsph::ef_acceleration<D>(ef, x + (size_t)i * D, acc + (size_t)i * D);
    pot[i] = fsph::ef_potential<D>(ef, x + (size_t)i * D);
  }
}

extern "C" int ef_host_eval(const fsph_params * p, int n, const double * x, double * acc, double * pot, int * nterms) {
  try {
    const fsph::ExtForce ef = fsph::ef_parse(*p);
    *nterms = ef.n;
    if(p->dim == 1) eval<1>(ef, n, x, acc, pot);
    else if(p->dim == 2) eval<2>(ef, n, x, acc, pot);
    else eval<3>(ef, n, x, acc, pot);
    return 0;
  }
  catch(const std::exception &) { return 1; }
}
