// io.h -- particle file I/O entry points of the reference (mpisph/io.h:631-637 io::inputDataHDF5,
// :874-878 io::outputDataHDF5) for builds WITHOUT an HDF5 installation.
//
// The reference's body_system reads and writes H5part files through these two functions
// (mpisph/bodies_system.h:83-104).  When <hdf5.h> is available this header simply forwards to the
// reference's own mpisph/io.h (`#include_next`), so H5part I/O is unchanged.  This image has no HDF5,
// so otherwise the same two functions move the same datasets (names of doc/H5_STRUCTURE.md: x y z vx vy vz
// ax ay az h rho u P m dt id type state neighbors, step attributes time / iteration / timestep) through a
// flat self-describing container:
//
//   "FSPHBIN1"  int64 n  int32 dim  int32 nsets  int64 iteration  double time  double timestep
//   nsets x { char name[16]; int32 dtype (0 f64, 1 i64, 2 i32); int32 reserved; data[n] }
//
// input:  <prefix>.fsphbin                       (single step, the initial data)
// output: <prefix>_<step:05d>.fsphbin            (one file per output step)
// flecsph_b200/fsphbin.py reads and writes the same container from numpy.
#pragma once
#if __has_include(<hdf5.h>) && !defined(FSPH_FORCE_FLAT_IO)
#include_next "io.h"
#else
#ifndef _mpisph_io_h_
#define _mpisph_io_h_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <mpi.h>

#include "default_physics.h" // as mpisph/io.h:35 (the reference's bodies_system.h gets its physics:: names this way)
#include "params.h"

namespace io {

inline int output_step = 0;

struct flat_set {
  char name[16];
  int32_t dtype;
  int32_t reserved;
};

inline void
flat_fail(const std::string & what) {
  std::fprintf(stderr, "flecsph_b200 io: %s\n", what.c_str());
  MPI_Abort(MPI_COMM_WORLD, 1);
}

// Reads <prefix>.fsphbin; every rank takes its contiguous share like the reference (mpisph/io.h:786-793).
inline void
inputDataHDF5(std::vector<body> & bodies,
  const char * input_file_prefix,
  const char * output_file_prefix,
  int64_t & totalnbodies,
  int64_t & nbodies,
  int startIteration,
  MPI_Comm comm = MPI_COMM_WORLD) {
  int rank = 0, size = 1;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  (void)output_file_prefix;
  const std::string fn = std::string(input_file_prefix) + ".fsphbin";
  FILE * f = std::fopen(fn.c_str(), "rb");
  if(!f)
    flat_fail("cannot open " + fn);
  char magic[8];
  int64_t n = 0, iteration = 0;
  int32_t dim = 0, nsets = 0;
  double time = 0, timestep = 0;
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "FSPHBIN1", 8) == 0;
  ok = ok && std::fread(&n, 8, 1, f) == 1 && std::fread(&dim, 4, 1, f) == 1 && std::fread(&nsets, 4, 1, f) == 1;
  ok = ok && std::fread(&iteration, 8, 1, f) == 1 && std::fread(&time, 8, 1, f) == 1 && std::fread(&timestep, 8, 1, f) == 1;
  if(!ok || dim != (int32_t)gdimension)
    flat_fail("bad header or dimension in " + fn);
  int64_t share = n / size;
  const int64_t rem = n % size;
  const int64_t first = rank * share + std::min<int64_t>(rank, rem);
  if(rank < rem)
    ++share;
  totalnbodies = n;
  nbodies = share;
  bodies.clear();
  bodies.resize(share);
  if(startIteration != 0) {
    physics::dt = timestep;
    physics::totaltime = time;
  }
  std::vector<double> buf(share + 1);
  for(int32_t k = 0; k < nsets; ++k) {
    flat_set hd;
    if(std::fread(&hd, sizeof(hd), 1, f) != 1)
      flat_fail("truncated " + fn);
    const size_t esz = hd.dtype == 2 ? 4 : 8;
    const long at = std::ftell(f);
    std::fseek(f, at + (long)(first * esz), SEEK_SET);
    if(share && std::fread(buf.data(), esz, share, f) != (size_t)share)
      flat_fail("truncated " + fn);
    std::fseek(f, at + (long)(n * esz), SEEK_SET);
    const std::string nm(hd.name, strnlen(hd.name, 16));
    const double * d = buf.data();
    const int64_t * l = reinterpret_cast<const int64_t *>(buf.data());
    const int32_t * w = reinterpret_cast<const int32_t *>(buf.data());
    auto axis = [&](const char * base) -> int {
      for(int a = 0; a < (int)gdimension; ++a)
        if(nm == std::string(base) + "xyz"[a])
          return a;
      return -1;
    };
    int a;
    if((a = axis("")) >= 0) {
      for(int64_t i = 0; i < share; ++i) {
        point_t p = bodies[i].coordinates();
        p[a] = d[i];
        bodies[i].set_coordinates(p);
      }
    }
    else if((a = axis("v")) >= 0) {
      for(int64_t i = 0; i < share; ++i) {
        point_t p = bodies[i].getVelocity();
        p[a] = d[i];
        bodies[i].setVelocity(p);
      }
    }
    else if(nm == "m")
      for(int64_t i = 0; i < share; ++i) bodies[i].set_mass(d[i]);
    else if(nm == "rho")
      for(int64_t i = 0; i < share; ++i) bodies[i].setDensity(d[i]);
    else if(nm == "h")
      for(int64_t i = 0; i < share; ++i) bodies[i].set_radius(d[i]);
    else if(nm == "P")
      for(int64_t i = 0; i < share; ++i) bodies[i].setPressure(d[i]);
    else if(nm == "u")
      for(int64_t i = 0; i < share; ++i) bodies[i].setInternalenergy(d[i]);
    else if(nm == "dt")
      for(int64_t i = 0; i < share; ++i) bodies[i].setDt(d[i]);
    else if(nm == "id")
      for(int64_t i = 0; i < share; ++i) bodies[i].set_id((size_t)l[i]);
    else if(nm == "type")
      for(int64_t i = 0; i < share; ++i) bodies[i].setType((int)w[i]);
    else if(nm == "state")
      for(int64_t i = 0; i < share; ++i) bodies[i].set_state((int)w[i]);
  }
  std::fclose(f);
}

// Writes <prefix>_<step>.fsphbin with the datasets of mpisph/io.h:945-1070 (a single rank writes all bodies;
// with several ranks each rank writes its share to "<file>.r<rank>": there is no parallel file layer here).
inline void
outputDataHDF5(std::vector<body> & bodies,
  const char * fileprefix,
  int64_t iteration,
  double totaltime,
  MPI_Comm comm = MPI_COMM_WORLD) {
  int rank = 0, size = 1;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  const int step = output_step++;
  char fn[4096];
  if(size == 1)
    std::snprintf(fn, sizeof(fn), "%s_%05d.fsphbin", fileprefix, step);
  else
    std::snprintf(fn, sizeof(fn), "%s_%05d.fsphbin.r%d", fileprefix, step, rank);
  FILE * f = std::fopen(fn, "wb");
  if(!f)
    flat_fail(std::string("cannot create ") + fn);
  const int64_t n = (int64_t)bodies.size();
  const int32_t dim = gdimension;
  struct out_set {
    std::string name;
    int dtype;
  };
  std::vector<out_set> sets;
  const char * ax = "xyz";
  for(int a = 0; a < dim; ++a) sets.push_back({std::string(1, ax[a]), 0});
  for(int a = 0; a < dim; ++a) sets.push_back({std::string("v") + ax[a], 0});
  for(int a = 0; a < dim; ++a) sets.push_back({std::string("a") + ax[a], 0});
  for(const char * nm : {"h", "rho", "u", "P", "m", "dt"}) sets.push_back({nm, 0});
  sets.push_back({"id", 1});
  sets.push_back({"type", 2});
  sets.push_back({"state", 2});
  sets.push_back({"neighbors", 1});
  const int32_t nsets = (int32_t)sets.size();
  std::fwrite("FSPHBIN1", 1, 8, f);
  std::fwrite(&n, 8, 1, f);
  std::fwrite(&dim, 4, 1, f);
  std::fwrite(&nsets, 4, 1, f);
  std::fwrite(&iteration, 8, 1, f);
  std::fwrite(&totaltime, 8, 1, f);
  std::fwrite(&physics::dt, 8, 1, f);
  std::vector<double> d(n + 1);
  std::vector<int64_t> l(n + 1);
  std::vector<int32_t> w(n + 1);
  for(const auto & s : sets) {
    flat_set hd;
    std::memset(&hd, 0, sizeof(hd));
    std::strncpy(hd.name, s.name.c_str(), 15);
    hd.dtype = s.dtype;
    std::fwrite(&hd, sizeof(hd), 1, f);
    const std::string & nm = s.name;
    for(int64_t i = 0; i < n; ++i) {
      body & b = bodies[i];
      if(nm.size() == 1 && nm[0] >= 'x' && nm[0] <= 'z') d[i] = b.coordinates()[nm[0] - 'x'];
      else if(nm.size() == 2 && nm[0] == 'v') d[i] = b.getVelocity()[nm[1] - 'x'];
      else if(nm.size() == 2 && nm[0] == 'a') d[i] = b.getAcceleration()[nm[1] - 'x'] + b.getGAcceleration()[nm[1] - 'x'];
      else if(nm == "h") d[i] = b.radius();
      else if(nm == "rho") d[i] = b.getDensity();
      else if(nm == "u") d[i] = b.getInternalenergy();
      else if(nm == "P") d[i] = b.getPressure();
      else if(nm == "m") d[i] = b.mass();
      else if(nm == "dt") d[i] = b.getDt();
      else if(nm == "id") l[i] = (int64_t)b.id();
      else if(nm == "type") w[i] = (int32_t)b.getType();
      else if(nm == "state") w[i] = (int32_t)b.state();
      else if(nm == "neighbors") l[i] = (int64_t)b.getNeighbors();
    }
    if(s.dtype == 0) std::fwrite(d.data(), 8, n, f);
    else if(s.dtype == 1) std::fwrite(l.data(), 8, n, f);
    else std::fwrite(w.data(), 4, n, f);
  }
  std::fclose(f);
}

} // namespace io
#endif // _mpisph_io_h_
#endif // no HDF5
