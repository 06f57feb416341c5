/*
 * h5part_harness.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's particle I/O, mpisph/io.h, UNMODIFIED (read where it lies under /root/reference at build time),
 * compiled against flecsph_b200/host/h5mini/hdf5.h -- the stand-in for the HDF5 C API this image lacks -- and exposed
 * through three C entry points, so that tests/test_h5part.py can do what the reference's own tests do: write bodies
 * and read them back (mpisph/test/io.cc:22-67, "write_N_read") and read the golden file mpisph/test/io_test.h5part
 * (mpisph/test/bs.cc:24).  Built by oracle/Makefile into oracle/_ref/libh5part_ref_3d.so.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "io.h"

namespace flecsi { namespace execution {
void driver(int, char **) {}
void specialization_tlt_init(int, char **) {}
} }

static constexpr int D = gdimension;

extern "C" {

// io::outputDataHDF5 (mpisph/io.h:874): one output step of `n` bodies; step numbering restarts when `reset_step` is set
int
h5ref_write(const char * prefix, int64_t n, const double * pos, const double * vel, const double * mass, const double * h,
  const double * rho, const double * u, const int64_t * id, int64_t iteration, double time, double dt, int separate, int reset_step) {
  log_set_output_rank(0);
  param::_out_h5data_separate_iterations = separate != 0;
  if(reset_step) io::output_step = 0;
  physics::dt = dt;
  physics::totaltime = time;
  std::vector<body> bodies((size_t)n);
  for(int64_t i = 0; i < n; ++i) {
    point_t p, v;
    for(int d = 0; d < D; ++d) { p[d] = pos[i * D + d]; v[d] = vel[i * D + d]; }
    bodies[i].set_coordinates(p);
    bodies[i].setVelocity(v);
    bodies[i].set_mass(mass[i]);
    bodies[i].set_radius(h[i]);
    bodies[i].setDensity(rho[i]);
    bodies[i].setInternalenergy(u[i]);
    bodies[i].set_id(id[i]);
  }
  io::outputDataHDF5(bodies, prefix, iteration, time);
  return 0;
}

// io::inputDataHDF5 (mpisph/io.h:631) of <prefix>.h5part from iteration 0; arrays hold at least `cap` bodies
int64_t
h5ref_read(const char * prefix, int64_t cap, double * pos, double * vel, double * mass, double * h, double * rho, int64_t * id) {
  log_set_output_rank(0);
  std::vector<body> bodies;
  int64_t total = 0, local = 0;
  const std::string out = std::string(prefix) + "_unused_output";
  io::inputDataHDF5(bodies, prefix, out.c_str(), total, local, 0);
  if(local != (int64_t)bodies.size() || total != local) return -1;
  if(local > cap) return -2;
  for(int64_t i = 0; i < local; ++i) {
    for(int d = 0; d < D; ++d) { pos[i * D + d] = bodies[i].coordinates()[d]; vel[i * D + d] = bodies[i].getVelocity()[d]; }
    mass[i] = bodies[i].mass();
    h[i] = bodies[i].radius();
    rho[i] = bodies[i].getDensity();
    id[i] = bodies[i].id();
  }
  return local;
}

}
