// h5tool -- command-line companion of h5mini.h (no libhdf5):
//   h5tool dump  FILE                   structure + a checksum of every dataset (what tests/h5probe.py prints for the same file)
//   h5tool copy  IN OUT                 read IN, write its content again through the writer
//   h5tool fromflat IN.fsphbin OUT.h5part   the flat container of flatio/io.h -> an H5part initial-data file (root
//                                       attributes nparticles / timestep / dimension / use_fixed_timestep, group Step#0)
//   h5tool toflat IN.h5part STEP OUT.fsphbin   one step of an H5part file -> the flat container
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "h5mini.h"

using namespace h5mini;

static const char * tname(TypeCode t) {
  switch(t) { case T_F64: return "float64"; case T_F32: return "float32"; case T_I64: return "int64"; case T_U64: return "uint64";
              case T_I32: return "int32"; case T_U32: return "uint32"; default: return "?"; }
}
static void print_attr(const Attr & a, const char * ind) {
  std::printf("%s@%s %s[%" PRIu64 "] =", ind, a.name.c_str(), tname(a.type), a.count);
  for(uint64_t i = 0; i < a.count && i < 4; ++i) {
    const uint8_t * p = a.data.data() + i * type_size(a.type);
    if(a.type == T_F64) { double v; std::memcpy(&v, p, 8); std::printf(" %.17g", v); }
    else if(a.type == T_F32) { float v; std::memcpy(&v, p, 4); std::printf(" %.9g", v); }
    else if(type_size(a.type) == 8) { int64_t v; std::memcpy(&v, p, 8); std::printf(" %" PRId64, v); }
    else { int32_t v; std::memcpy(&v, p, 4); std::printf(" %d", v); }
  }
  std::printf("\n");
}
// FNV-1a over the raw little-endian bytes
static uint64_t checksum(const File & f, const Dataset & d) {
  std::vector<uint8_t> buf((size_t)d.n * type_size(d.type));
  if(d.n) f.read_data(d, 0, d.n, buf.data());
  uint64_t h = 1469598103934665603ull;
  for(uint8_t c : buf) { h ^= c; h *= 1099511628211ull; }
  return h;
}
static void dump_group(const File & f, const Group & g, const std::string & ind) {
  for(const auto & a : g.attrs) print_attr(a, ind.c_str());
  for(const auto & kv : g.datasets) {
    std::printf("%s%s %s (%" PRIu64 ",) fnv=%016" PRIx64 "\n", ind.c_str(), kv.first.c_str(), tname(kv.second.type), kv.second.n, checksum(f, kv.second));
    for(const auto & a : kv.second.attrs) print_attr(a, (ind + "  ").c_str());
  }
  for(const auto & kv : g.groups) {
    std::printf("%s%s/\n", ind.c_str(), kv.first.c_str());
    dump_group(f, kv.second, ind + "  ");
  }
}
static void copy_group(const File & in, const Group & g, File & out, Group & og) {
  og.attrs = g.attrs;
  for(const auto & kv : g.datasets) {
    Dataset & d = og.datasets[kv.first];
    d.type = kv.second.type; d.n = kv.second.n; d.attrs = kv.second.attrs;
    std::vector<uint8_t> buf((size_t)d.n * type_size(d.type));
    if(d.n) { in.read_data(kv.second, 0, d.n, buf.data()); out.write_data(d, 0, d.n, buf.data()); }
  }
  for(const auto & kv : g.groups) copy_group(in, kv.second, out, og.groups[kv.first]);
}
template<typename T> static Attr make_attr(const char * name, TypeCode t, T v) {
  Attr a;
  a.name = name; a.type = t; a.count = 1;
  a.data.resize(sizeof(T));
  std::memcpy(a.data.data(), &v, sizeof(T));
  return a;
}
struct FlatSet { char name[16]; int32_t dtype, reserved; };

int main(int argc, char ** argv) {
  try {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if(cmd == "dump" && argc == 3) {
      File f;
      f.open(argv[2], 'r');
      dump_group(f, f.root, "");
      return 0;
    }
    if(cmd == "copy" && argc == 4) {
      File in, out;
      in.open(argv[2], 'r');
      out.open(argv[3], 'w');
      copy_group(in, in.root, out, out.root);
      out.close();
      return 0;
    }
    if(cmd == "fromflat" && argc == 4) {
      FILE * fp = std::fopen(argv[2], "rb");
      if(!fp) throw std::runtime_error("cannot open the flat file");
      char magic[8]; int64_t n, iteration; int32_t dim, nsets; double time, timestep;
      if(std::fread(magic, 1, 8, fp) != 8 || std::memcmp(magic, "FSPHBIN1", 8) != 0) throw std::runtime_error("not a flat particle file");
      if(std::fread(&n, 8, 1, fp) != 1 || std::fread(&dim, 4, 1, fp) != 1 || std::fread(&nsets, 4, 1, fp) != 1 || std::fread(&iteration, 8, 1, fp) != 1 ||
         std::fread(&time, 8, 1, fp) != 1 || std::fread(&timestep, 8, 1, fp) != 1) throw std::runtime_error("short header");
      File out;
      out.open(argv[3], 'w');
      out.root.attrs.push_back(make_attr<int64_t>("nparticles", T_I64, n));
      out.root.attrs.push_back(make_attr<double>("timestep", T_F64, timestep));
      out.root.attrs.push_back(make_attr<int32_t>("dimension", T_I32, dim));
      out.root.attrs.push_back(make_attr<int32_t>("use_fixed_timestep", T_I32, 1));
      Group & st = out.root.groups["Step#0"];
      st.attrs.push_back(make_attr<double>("time", T_F64, time));
      st.attrs.push_back(make_attr<int64_t>("iteration", T_I64, iteration));
      st.attrs.push_back(make_attr<double>("timestep", T_F64, timestep));
      for(int k = 0; k < nsets; ++k) {
        FlatSet fs;
        if(std::fread(&fs, sizeof(fs), 1, fp) != 1) throw std::runtime_error("short set header");
        const TypeCode t = fs.dtype == 0 ? T_F64 : (fs.dtype == 1 ? T_I64 : T_I32);
        std::vector<uint8_t> buf((size_t)n * type_size(t));
        if(n && std::fread(buf.data(), buf.size(), 1, fp) != 1) throw std::runtime_error("short set data");
        Dataset & d = st.datasets[std::string(fs.name, strnlen(fs.name, 16))];
        d.type = t; d.n = (uint64_t)n;
        if(n) out.write_data(d, 0, d.n, buf.data());
      }
      std::fclose(fp);
      out.close();
      return 0;
    }
    if(cmd == "toflat" && argc == 5) {
      File in;
      in.open(argv[2], 'r');
      const std::string step = std::string("Step#") + argv[3];
      auto it = in.root.groups.find(step);
      if(it == in.root.groups.end()) throw std::runtime_error("no " + step);
      const Group & st = it->second;
      int64_t n = 0, iteration = 0; int32_t dim = 3; double time = 0, timestep = 0;
      for(const auto & a : in.root.attrs) if(a.name == "dimension") std::memcpy(&dim, a.data.data(), 4);
      for(const auto & a : st.attrs) {
        if(a.name == "time") std::memcpy(&time, a.data.data(), 8);
        if(a.name == "iteration") std::memcpy(&iteration, a.data.data(), 8);
        if(a.name == "timestep") std::memcpy(&timestep, a.data.data(), 8);
      }
      if(!st.datasets.empty()) n = (int64_t)st.datasets.begin()->second.n;
      FILE * fp = std::fopen(argv[4], "wb");
      if(!fp) throw std::runtime_error("cannot create the flat file");
      const int32_t nsets = (int32_t)st.datasets.size();
      std::fwrite("FSPHBIN1", 1, 8, fp); std::fwrite(&n, 8, 1, fp); std::fwrite(&dim, 4, 1, fp); std::fwrite(&nsets, 4, 1, fp);
      std::fwrite(&iteration, 8, 1, fp); std::fwrite(&time, 8, 1, fp); std::fwrite(&timestep, 8, 1, fp);
      for(const auto & kv : st.datasets) {
        FlatSet fs;
        std::memset(&fs, 0, sizeof(fs));
        std::strncpy(fs.name, kv.first.c_str(), 15);
        const Dataset & d = kv.second;
        if((int64_t)d.n != n) throw std::runtime_error("datasets of different lengths");
        fs.dtype = d.type == T_F64 ? 0 : (type_size(d.type) == 8 ? 1 : 2);
        if(d.type == T_F32) throw std::runtime_error("float32 datasets have no flat counterpart");
        std::vector<uint8_t> buf((size_t)d.n * type_size(d.type));
        if(d.n) in.read_data(d, 0, d.n, buf.data());
        std::fwrite(&fs, sizeof(fs), 1, fp);
        if(d.n) std::fwrite(buf.data(), buf.size(), 1, fp);
      }
      std::fclose(fp);
      return 0;
    }
    std::fprintf(stderr, "usage: h5tool dump FILE | copy IN OUT | fromflat IN.fsphbin OUT.h5part | toflat IN.h5part STEP OUT.fsphbin\n");
    return 2;
  }
  catch(const std::exception & e) {
    std::fprintf(stderr, "h5tool: %s\n", e.what());
    return 1;
  }
}
