// h5mini.h -- reader / writer for the subset of the HDF5 file format that H5part files use, without libhdf5.
//
// The reference stores particles as H5part files (mpisph/io.h:631-1083; layout doc/H5_STRUCTURE.md: root attributes,
// one group "Step#n" per output step with step attributes and one 1-D dataset per particle field).  The image has no
// HDF5 installation, so this header implements the on-disk format those files have when the library writes them with
// default settings -- the layout of the reference's own mpisph/test/io_test.h5part, which tests/test_h5part.py parses with
// this code and with an independent Python reader (tests/h5probe.py):
//   superblock version 0 (8-byte offsets and lengths), version-1 object headers (continuation blocks on read),
//   groups as symbol tables: version-1 B-tree of type 0 + local heap + symbol-table nodes (any depth on read; on write
//   one symbol node per group, the superblock's leaf K sized for the largest group),
//   datasets: simple 1-D dataspace (version 1), little-endian fixed-point / IEEE floating-point datatypes, fill-value
//   message, layout version 3 contiguous (versions 1-2 contiguous and compact on read),
//   attributes as version-1 attribute messages in the object header (versions 2-3 on read).
// Per the "HDF5 File Format Specification Version 2.0" (sections II.A, III.A.1, III.B, III.D, IV.A.1.a, IV.A.2.b/d/f/i/m/r/s).
// Not supported (error on read): chunked or filtered datasets, object headers of version 2, link-message groups, other
// datatype classes, big-endian data, more than one dimension.
//
// Data blocks go to disk when they are written (append at the end of the file); the metadata lives in memory and is
// written behind the data when the file is closed, then the superblock is pointed at the new root.  Re-opening a file for
// writing loads its metadata, appends, and writes a fresh set of metadata (the old blocks become dead space).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace h5mini {

constexpr uint64_t UNDEF = ~uint64_t(0);

enum TypeCode { T_F64 = 0, T_I64, T_U64, T_I32, T_U32, T_F32, T_NONE };
inline size_t type_size(TypeCode t) { return (t == T_I32 || t == T_U32 || t == T_F32) ? 4 : 8; }
inline bool type_is_float(TypeCode t) { return t == T_F64 || t == T_F32; }
// memory type `m` can be filled from file type `f` without conversion (same class and width; signedness is ignored as
// the library's integer conversions are value-preserving for the ranges in question)
inline bool type_compatible(TypeCode m, TypeCode f) {
  return type_size(m) == type_size(f) && type_is_float(m) == type_is_float(f);
}

struct Attr {
  std::string name;
  TypeCode type = T_F64;
  uint64_t count = 1;
  std::vector<uint8_t> data;
};

struct Dataset {
  TypeCode type = T_F64;
  uint64_t n = 0;                    // elements (1-D)
  uint64_t addr = UNDEF;             // contiguous data in the file; UNDEF: nothing written yet (reads as zeros)
  std::vector<uint8_t> compact;      // compact layout (read only)
  bool is_compact = false;
  std::vector<Attr> attrs;
};

struct Group {
  std::vector<Attr> attrs;
  std::map<std::string, Dataset> datasets;
  std::map<std::string, Group> groups;
};

inline Attr * find_attr(std::vector<Attr> & v, const std::string & name) {
  for(auto & a : v)
    if(a.name == name) return &a;
  return nullptr;
}

class File {
public:
  Group root;
  ~File() { if(fd_ >= 0) ::close(fd_); }

  // mode: 'r' read only, 'a' read-write on an existing file, 'w' create / truncate
  void open(const std::string & path, char mode) {
    path_ = path;
    writable_ = mode != 'r';
    fd_ = ::open(path.c_str(), mode == 'r' ? O_RDONLY : (mode == 'a' ? O_RDWR : (O_RDWR | O_CREAT | O_TRUNC)), 0644);
    if(fd_ < 0) throw std::runtime_error("h5mini: cannot open " + path);
    if(mode == 'w') {
      eof_ = 96;   // superblock + root symbol-table entry: written on close
      dirty_ = true;
      std::vector<uint8_t> z(96, 0);
      put(0, z.data(), z.size());
    }
    else {
      struct stat st;
      if(fstat(fd_, &st) != 0) throw std::runtime_error("h5mini: cannot stat " + path);
      size_ = (uint64_t)st.st_size;
      load();
    }
  }
  bool is_open() const { return fd_ >= 0; }
  bool writable() const { return writable_; }
  void touch() { dirty_ = true; }

  // data of a dataset: elements [off, off + count)
  void write_data(Dataset & d, uint64_t off, uint64_t count, const void * src) {
    if(!writable_) throw std::runtime_error("h5mini: file is read-only");
    if(off + count > d.n) throw std::runtime_error("h5mini: write beyond the dataset's extent");
    const size_t es = type_size(d.type);
    if(d.addr == UNDEF) {
      // early allocation of the whole extent, as the (parallel) library does: zero-filled
      d.addr = alloc(d.n * es);
      if(count < d.n) {
        std::vector<uint8_t> z((size_t)std::min<uint64_t>(d.n * es, 1u << 20), 0);
        for(uint64_t p = 0; p < d.n * es; p += z.size()) put(d.addr + p, z.data(), (size_t)std::min<uint64_t>(z.size(), d.n * es - p));
      }
    }
    put(d.addr + off * es, src, (size_t)(count * es));
    dirty_ = true;
  }
  void read_data(const Dataset & d, uint64_t off, uint64_t count, void * dst) const {
    if(off + count > d.n) throw std::runtime_error("h5mini: read beyond the dataset's extent");
    const size_t es = type_size(d.type);
    if(d.is_compact) std::memcpy(dst, d.compact.data() + off * es, (size_t)(count * es));
    else if(d.addr == UNDEF) std::memset(dst, 0, (size_t)(count * es));
    else get(d.addr + off * es, dst, (size_t)(count * es));
  }

  void close() {
    if(fd_ < 0) return;
    if(writable_ && dirty_) save();
    ::close(fd_);
    fd_ = -1;
  }

private:
  std::string path_;
  int fd_ = -1;
  bool writable_ = false, dirty_ = false;
  uint64_t eof_ = 0, size_ = 0;

  // ---- raw I/O ------------------------------------------------------------------------------------------------
  void put(uint64_t at, const void * p, size_t n) {
    const char * c = static_cast<const char *>(p);
    while(n > 0) {
      const ssize_t w = ::pwrite(fd_, c, n, (off_t)at);
      if(w <= 0) throw std::runtime_error("h5mini: write failed on " + path_);
      c += w; at += (uint64_t)w; n -= (size_t)w;
    }
  }
  void get(uint64_t at, void * p, size_t n) const {
    char * c = static_cast<char *>(p);
    while(n > 0) {
      const ssize_t r = ::pread(fd_, c, n, (off_t)at);
      if(r <= 0) throw std::runtime_error("h5mini: read failed on " + path_ + " (truncated file?)");
      c += r; at += (uint64_t)r; n -= (size_t)r;
    }
  }
  std::vector<uint8_t> bytes(uint64_t at, size_t n) const {
    if(at == UNDEF || at + n > size_) throw std::runtime_error("h5mini: address outside the file");
    std::vector<uint8_t> v(n);
    get(at, v.data(), n);
    return v;
  }
  uint64_t alloc(uint64_t n) {
    const uint64_t at = (eof_ + 7) & ~uint64_t(7);
    eof_ = at + n;
    return at;
  }
  template<typename T> static T le(const uint8_t * p) { T v; std::memcpy(&v, p, sizeof(T)); return v; }   // little-endian hosts
  template<typename T> static void app(std::vector<uint8_t> & b, T v) { const uint8_t * p = reinterpret_cast<const uint8_t *>(&v); b.insert(b.end(), p, p + sizeof(T)); }
  static void pad8(std::vector<uint8_t> & b) { while(b.size() % 8) b.push_back(0); }

  // ---- reading ---------------------------------------------------------------------------------------------------
  struct SymEntry { uint64_t name_off, ohdr; uint32_t cache; uint64_t btree, heap; };
  static SymEntry sym_entry(const uint8_t * p) {
    SymEntry e;
    e.name_off = le<uint64_t>(p); e.ohdr = le<uint64_t>(p + 8); e.cache = le<uint32_t>(p + 16);
    e.btree = le<uint64_t>(p + 24); e.heap = le<uint64_t>(p + 32);
    return e;
  }
  struct Msg { uint16_t type; uint8_t flags; std::vector<uint8_t> body; };

  void load() {
    const std::vector<uint8_t> sb = bytes(0, 96 + 4);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if(std::memcmp(sb.data(), sig, 8) != 0) throw std::runtime_error("h5mini: " + path_ + " is not an HDF5 file");
    const int ver = sb[8];
    if(ver > 1) throw std::runtime_error("h5mini: superblock version " + std::to_string(ver) + " is not supported (classic layout only)");
    if(sb[13] != 8 || sb[14] != 8) throw std::runtime_error("h5mini: only 8-byte offsets and lengths");
    size_t p = ver == 0 ? 24 : 28;
    if(le<uint64_t>(&sb[p]) != 0) throw std::runtime_error("h5mini: non-zero base address");
    eof_ = le<uint64_t>(&sb[p + 16]);
    if(eof_ > size_) throw std::runtime_error("h5mini: truncated file");
    const SymEntry re = sym_entry(&sb[p + 32]);
    root = Group();
    load_group(re.ohdr, root, 0);
  }

  std::vector<Msg> messages(uint64_t addr) const {
    const std::vector<uint8_t> h = bytes(addr, 16);
    if(h[0] != 1) throw std::runtime_error("h5mini: object header version " + std::to_string(h[0]) + " is not supported");
    const unsigned nmsg = le<uint16_t>(&h[2]);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, le<uint32_t>(&h[8])}};
    std::vector<Msg> out;
    for(size_t bi = 0; bi < blocks.size() && out.size() < nmsg; ++bi) {
      const std::vector<uint8_t> blk = bytes(blocks[bi].first, (size_t)blocks[bi].second);
      size_t p = 0;
      while(p + 8 <= blk.size() && out.size() < nmsg) {
        Msg m;
        m.type = le<uint16_t>(&blk[p]);
        const size_t sz = le<uint16_t>(&blk[p + 2]);
        m.flags = blk[p + 4];
        if(p + 8 + sz > blk.size()) throw std::runtime_error("h5mini: message runs past its header block");
        m.body.assign(blk.begin() + p + 8, blk.begin() + p + 8 + sz);
        p += 8 + sz;
        if(m.type == 0x10) blocks.push_back({le<uint64_t>(&m.body[0]), le<uint64_t>(&m.body[8])});
        out.push_back(std::move(m));
      }
    }
    return out;
  }

  static TypeCode parse_type(const uint8_t * b, size_t n, size_t * used) {
    if(n < 8) throw std::runtime_error("h5mini: short datatype message");
    const int cls = b[0] & 15;
    const uint32_t size = le<uint32_t>(b + 4);
    if(b[1] & 1) throw std::runtime_error("h5mini: big-endian data is not supported");
    if(cls == 0) {
      if(used) *used = 12;
      const bool sgn = (b[1] & 8) != 0;
      if(size == 8) return sgn ? T_I64 : T_U64;
      if(size == 4) return sgn ? T_I32 : T_U32;
    }
    else if(cls == 1) {
      if(used) *used = 20;
      if(size == 8) return T_F64;
      if(size == 4) return T_F32;
    }
    throw std::runtime_error("h5mini: datatype class " + std::to_string(cls) + " / size " + std::to_string(size) + " is not supported");
  }
  // number of elements of a simple dataspace of rank <= 1 (scalar: 1)
  static uint64_t parse_space(const uint8_t * b, size_t n) {
    if(n < 4) throw std::runtime_error("h5mini: short dataspace message");
    const int ver = b[0], rank = b[1];
    if(ver == 2 && b[3] == 2) return 0;   // null dataspace
    if(rank == 0) return 1;
    if(rank != 1) throw std::runtime_error("h5mini: only 1-D dataspaces");
    return le<uint64_t>(b + (ver == 1 ? 8 : 4));
  }
  static Attr parse_attr(const std::vector<uint8_t> & b) {
    const int ver = b[0];
    if(ver < 1 || ver > 3) throw std::runtime_error("h5mini: attribute message version");
    const size_t nsz = le<uint16_t>(&b[2]), tsz = le<uint16_t>(&b[4]), ssz = le<uint16_t>(&b[6]);
    size_t p = ver == 3 ? 9 : 8;
    auto padv = [&](size_t v) { return ver == 1 ? ((v + 7) & ~size_t(7)) : v; };
    Attr a;
    a.name = std::string(reinterpret_cast<const char *>(&b[p]));
    p += padv(nsz);
    a.type = parse_type(&b[p], tsz, nullptr);
    p += padv(tsz);
    a.count = parse_space(&b[p], ssz);
    p += padv(ssz);
    const size_t bytes_ = (size_t)a.count * type_size(a.type);
    if(p + bytes_ > b.size()) throw std::runtime_error("h5mini: short attribute data");
    a.data.assign(b.begin() + p, b.begin() + p + bytes_);
    return a;
  }

  void btree_entries(uint64_t addr, uint64_t heap_data, std::vector<std::pair<std::string, uint64_t>> & out, int depth) const {
    if(depth > 32) throw std::runtime_error("h5mini: B-tree too deep");
    const std::vector<uint8_t> h = bytes(addr, 24);
    if(std::memcmp(h.data(), "TREE", 4) != 0 || h[4] != 0) throw std::runtime_error("h5mini: bad group B-tree node");
    const int level = h[5];
    const unsigned used = le<uint16_t>(&h[6]);
    const std::vector<uint8_t> kc = bytes(addr + 24, (size_t)used * 16 + 8);
    for(unsigned i = 0; i < used; ++i) {
      const uint64_t child = le<uint64_t>(&kc[(size_t)i * 16 + 8]);
      if(level > 0) { btree_entries(child, heap_data, out, depth + 1); continue; }
      const std::vector<uint8_t> sh = bytes(child, 8);
      if(std::memcmp(sh.data(), "SNOD", 4) != 0) throw std::runtime_error("h5mini: bad symbol-table node");
      const unsigned ns = le<uint16_t>(&sh[6]);
      const std::vector<uint8_t> ents = bytes(child + 8, (size_t)ns * 40);
      for(unsigned k = 0; k < ns; ++k) {
        const SymEntry e = sym_entry(&ents[(size_t)k * 40]);
        std::string name;
        for(uint64_t q = heap_data + e.name_off;; ++q) {
          const uint8_t c = bytes(q, 1)[0];
          if(!c) break;
          name.push_back((char)c);
          if(name.size() > 4096) throw std::runtime_error("h5mini: unterminated link name");
        }
        out.push_back({name, e.ohdr});
      }
    }
  }

  void load_group(uint64_t ohdr, Group & g, int depth) {
    if(depth > 16) throw std::runtime_error("h5mini: groups nested too deeply");
    const std::vector<Msg> msgs = messages(ohdr);
    const Msg * st = nullptr;
    for(const auto & m : msgs) {
      if(m.type == 0x11) st = &m;
      if(m.type == 0x0C) g.attrs.push_back(parse_attr(m.body));
      if(m.type == 0x02 || m.type == 0x06) throw std::runtime_error("h5mini: link-message groups are not supported (classic symbol tables only)");
    }
    if(!st) throw std::runtime_error("h5mini: object is not a symbol-table group");
    const uint64_t btree = le<uint64_t>(&st->body[0]), heap = le<uint64_t>(&st->body[8]);
    const std::vector<uint8_t> hh = bytes(heap, 32);
    if(std::memcmp(hh.data(), "HEAP", 4) != 0) throw std::runtime_error("h5mini: bad local heap");
    const uint64_t heap_data = le<uint64_t>(&hh[24]);
    std::vector<std::pair<std::string, uint64_t>> ents;
    btree_entries(btree, heap_data, ents, 0);
    for(const auto & e : ents) {
      const std::vector<Msg> cm = messages(e.second);
      bool is_group = false;
      for(const auto & m : cm) is_group = is_group || m.type == 0x11;
      if(is_group) load_group(e.second, g.groups[e.first], depth + 1);
      else g.datasets[e.first] = load_dataset(cm, e.first);
    }
  }

  Dataset load_dataset(const std::vector<Msg> & msgs, const std::string & name) const {
    Dataset d;
    bool have_t = false, have_s = false, have_l = false;
    for(const auto & m : msgs) {
      const std::vector<uint8_t> & b = m.body;
      if(m.type == 0x01) { d.n = parse_space(b.data(), b.size()); have_s = true; }
      else if(m.type == 0x03) { d.type = parse_type(b.data(), b.size(), nullptr); have_t = true; }
      else if(m.type == 0x0B) throw std::runtime_error("h5mini: dataset " + name + " has a filter pipeline");
      else if(m.type == 0x0C) d.attrs.push_back(parse_attr(b));
      else if(m.type == 0x08) {
        have_l = true;
        const int ver = b[0];
        if(ver == 3) {
          const int cls = b[1];
          if(cls == 1) d.addr = le<uint64_t>(&b[2]);
          else if(cls == 0) { const size_t sz = le<uint16_t>(&b[2]); d.compact.assign(b.begin() + 4, b.begin() + 4 + sz); d.is_compact = true; }
          else throw std::runtime_error("h5mini: dataset " + name + " is chunked (not supported)");
        }
        else if(ver == 1 || ver == 2) {
          const int rank = b[1], cls = b[2];
          if(cls == 1) d.addr = le<uint64_t>(&b[8]);
          else if(cls == 0) { const size_t p = 8 + 4 * (size_t)rank; const size_t sz = le<uint32_t>(&b[p]); d.compact.assign(b.begin() + p + 4, b.begin() + p + 4 + sz); d.is_compact = true; }
          else throw std::runtime_error("h5mini: dataset " + name + " is chunked (not supported)");
        }
        else throw std::runtime_error("h5mini: data layout version " + std::to_string(ver) + " is not supported");
      }
    }
    if(!have_t || !have_s || !have_l) throw std::runtime_error("h5mini: object " + name + " is neither a group nor a dataset");
    if(!d.is_compact && d.addr != UNDEF && d.addr + d.n * type_size(d.type) > size_) throw std::runtime_error("h5mini: dataset " + name + " runs past the end of the file");
    return d;
  }

  // ---- writing ---------------------------------------------------------------------------------------------------
  static void type_msg(std::vector<uint8_t> & b, TypeCode t) {
    // IV.A.2.d: class + version 1, class bit field, size, properties
    const uint32_t sz = (uint32_t)type_size(t);
    if(type_is_float(t)) {
      // little-endian IEEE: mantissa normalisation "implied msb" (bits 4-5 = 2), sign bit position in byte 1
      b.push_back(0x11); b.push_back(0x20); b.push_back(sz == 8 ? 63 : 31); b.push_back(0);
      app<uint32_t>(b, sz);
      app<uint16_t>(b, 0); app<uint16_t>(b, (uint16_t)(sz * 8));
      if(sz == 8) { b.push_back(52); b.push_back(11); b.push_back(0); b.push_back(52); app<uint32_t>(b, 1023); }
      else { b.push_back(23); b.push_back(8); b.push_back(0); b.push_back(23); app<uint32_t>(b, 127); }
    }
    else {
      const bool sgn = t == T_I64 || t == T_I32;
      b.push_back(0x10); b.push_back(sgn ? 0x08 : 0x00); b.push_back(0); b.push_back(0);
      app<uint32_t>(b, sz);
      app<uint16_t>(b, 0); app<uint16_t>(b, (uint16_t)(sz * 8));
    }
  }
  static void space_msg(std::vector<uint8_t> & b, uint64_t n) {
    // IV.A.2.b version 1, rank 1, maximum dimensions present (as the library writes it)
    b.push_back(1); b.push_back(1); b.push_back(1); b.push_back(0);
    app<uint32_t>(b, 0);
    app<uint64_t>(b, n);
    app<uint64_t>(b, n);
  }
  static void add_msg(std::vector<uint8_t> & hdr, unsigned & nmsg, uint16_t type, uint8_t flags, std::vector<uint8_t> body) {
    pad8(body);
    app<uint16_t>(hdr, type);
    app<uint16_t>(hdr, (uint16_t)body.size());
    hdr.push_back(flags); hdr.push_back(0); hdr.push_back(0); hdr.push_back(0);
    hdr.insert(hdr.end(), body.begin(), body.end());
    ++nmsg;
  }
  static void attr_msgs(std::vector<uint8_t> & hdr, unsigned & nmsg, const std::vector<Attr> & attrs) {
    for(const Attr & a : attrs) {
      // IV.A.2.m version 1: name, datatype and dataspace each padded to a multiple of 8
      std::vector<uint8_t> t, s, b;
      type_msg(t, a.type);
      space_msg(s, a.count);
      if(a.name.size() + 1 > 0xFFFF) throw std::runtime_error("h5mini: attribute name too long");
      b.push_back(1); b.push_back(0);
      app<uint16_t>(b, (uint16_t)(a.name.size() + 1));
      app<uint16_t>(b, (uint16_t)t.size());
      app<uint16_t>(b, (uint16_t)s.size());
      b.insert(b.end(), a.name.begin(), a.name.end());
      b.push_back(0);
      pad8(b);
      b.insert(b.end(), t.begin(), t.end());
      pad8(b);
      b.insert(b.end(), s.begin(), s.end());
      pad8(b);
      b.insert(b.end(), a.data.begin(), a.data.end());
      if(b.size() > 0xFFF8) throw std::runtime_error("h5mini: attribute " + a.name + " is too large for an object-header message");
      add_msg(hdr, nmsg, 0x0C, 0, b);
    }
  }
  uint64_t write_ohdr(const std::vector<uint8_t> & msgs, unsigned nmsg) {
    // IV.A.1.a version-1 prefix: version, reserved, number of messages, reference count, header size, padding to 8
    std::vector<uint8_t> h;
    h.push_back(1); h.push_back(0);
    app<uint16_t>(h, (uint16_t)nmsg);
    app<uint32_t>(h, 1);
    app<uint32_t>(h, (uint32_t)msgs.size());
    app<uint32_t>(h, 0);
    h.insert(h.end(), msgs.begin(), msgs.end());
    const uint64_t at = alloc(h.size());
    put(at, h.data(), h.size());
    return at;
  }
  uint64_t write_dataset(const Dataset & d, uint32_t mtime) {
    std::vector<uint8_t> m, b;
    unsigned nmsg = 0;
    space_msg(b, d.n);
    add_msg(m, nmsg, 0x01, 0, b);
    b.clear(); type_msg(b, d.type);
    add_msg(m, nmsg, 0x03, 1, b);
    // IV.A.2.f fill value version 2: space allocation time early, fill-value write time "if set", fill value undefined
    b = {2, 1, 2, 1, 0, 0, 0, 0};
    add_msg(m, nmsg, 0x05, 1, b);
    // IV.A.2.i layout version 3, contiguous: address, size
    b = {3, 1};
    uint64_t addr = d.addr;
    if(d.is_compact) {
      // data that came in compact form goes out contiguous
      addr = alloc(d.compact.size());
      put(addr, d.compact.data(), d.compact.size());
    }
    app<uint64_t>(b, addr);
    app<uint64_t>(b, d.n * type_size(d.type));
    add_msg(m, nmsg, 0x08, 1, b);
    b = {1, 0, 0, 0};
    app<uint32_t>(b, mtime);
    add_msg(m, nmsg, 0x12, 0, b);
    attr_msgs(m, nmsg, d.attrs);
    return write_ohdr(m, nmsg);
  }

  struct GroupAddr { uint64_t ohdr, btree, heap; };
  static size_t max_links(const Group & g) {
    size_t m = g.datasets.size() + g.groups.size();
    for(const auto & kv : g.groups) m = std::max(m, max_links(kv.second));
    return m;
  }
  GroupAddr write_group(const Group & g, unsigned leaf_k, uint32_t mtime) {
    // children first: their object-header addresses go into this group's symbol-table node
    struct Link { std::string name; uint64_t ohdr; bool group; uint64_t btree, heap; };
    std::vector<Link> links;
    for(const auto & kv : g.datasets) links.push_back({kv.first, write_dataset(kv.second, mtime), false, 0, 0});
    for(const auto & kv : g.groups) {
      const GroupAddr a = write_group(kv.second, leaf_k, mtime);
      links.push_back({kv.first, a.ohdr, true, a.btree, a.heap});
    }
    std::sort(links.begin(), links.end(), [](const Link & a, const Link & b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    // III.D local heap: the empty name at offset 0, the link names (each padded to 8), one free block at the end
    std::vector<uint8_t> hd(8, 0);
    std::vector<uint64_t> noff;
    for(const Link & l : links) {
      noff.push_back(hd.size());
      hd.insert(hd.end(), l.name.begin(), l.name.end());
      hd.push_back(0);
      pad8(hd);
    }
    const uint64_t free_off = hd.size();
    app<uint64_t>(hd, 1);    // end of the free list
    app<uint64_t>(hd, 32);   // size of this free block
    hd.resize(hd.size() + 16, 0);
    std::vector<uint8_t> hh;
    hh.insert(hh.end(), {'H', 'E', 'A', 'P', 0, 0, 0, 0});
    app<uint64_t>(hh, hd.size());
    app<uint64_t>(hh, free_off);
    const uint64_t heap = alloc(32 + hd.size());
    app<uint64_t>(hh, heap + 32);
    put(heap, hh.data(), hh.size());
    put(heap + 32, hd.data(), hd.size());
    // III.B symbol-table node (capacity 2 * leaf K entries), III.A.1 B-tree node of type 0 (capacity 2 * 16 children)
    uint64_t snod = UNDEF;
    if(!links.empty()) {
      std::vector<uint8_t> sn;
      sn.insert(sn.end(), {'S', 'N', 'O', 'D', 1, 0});
      app<uint16_t>(sn, (uint16_t)links.size());
      for(size_t i = 0; i < links.size(); ++i) {
        app<uint64_t>(sn, noff[i]);
        app<uint64_t>(sn, links[i].ohdr);
        app<uint32_t>(sn, links[i].group ? 1u : 0u);
        app<uint32_t>(sn, 0);
        app<uint64_t>(sn, links[i].group ? links[i].btree : 0);
        app<uint64_t>(sn, links[i].group ? links[i].heap : 0);
      }
      sn.resize(8 + (size_t)2 * leaf_k * 40, 0);
      snod = alloc(sn.size());
      put(snod, sn.data(), sn.size());
    }
    std::vector<uint8_t> bt;
    bt.insert(bt.end(), {'T', 'R', 'E', 'E', 0, 0});
    app<uint16_t>(bt, links.empty() ? 0 : 1);
    app<uint64_t>(bt, UNDEF);
    app<uint64_t>(bt, UNDEF);
    app<uint64_t>(bt, 0);   // key 0: the empty name
    if(!links.empty()) {
      app<uint64_t>(bt, snod);
      app<uint64_t>(bt, noff.back());   // key 1: the largest name of the node
    }
    bt.resize(24 + (2 * 16 + 1) * 8 + 2 * 16 * 8, 0);
    const uint64_t btree = alloc(bt.size());
    put(btree, bt.data(), bt.size());
    std::vector<uint8_t> m, b;
    unsigned nmsg = 0;
    app<uint64_t>(b, btree);
    app<uint64_t>(b, heap);
    add_msg(m, nmsg, 0x11, 0, b);
    attr_msgs(m, nmsg, g.attrs);
    return {write_ohdr(m, nmsg), btree, heap};
  }

  void save() {
    const uint32_t mtime = (uint32_t)std::time(nullptr);
    const size_t links = max_links(root);
    if(links > 2 * 32767) throw std::runtime_error("h5mini: group with more than 65534 links");
    const unsigned leaf_k = (unsigned)std::max<size_t>(4, (links + 1) / 2);
    const GroupAddr r = write_group(root, leaf_k, mtime);
    const uint64_t end = (eof_ + 7) & ~uint64_t(7);
    if(end > eof_) { const uint8_t z[8] = {0}; put(eof_, z, (size_t)(end - eof_)); }
    // II.A superblock version 0 + the root group's symbol-table entry
    std::vector<uint8_t> sb;
    sb.insert(sb.end(), {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n', 0, 0, 0, 0, 0, 8, 8, 0});
    app<uint16_t>(sb, (uint16_t)leaf_k);
    app<uint16_t>(sb, 16);
    app<uint32_t>(sb, 0);
    app<uint64_t>(sb, 0);
    app<uint64_t>(sb, UNDEF);
    app<uint64_t>(sb, end);
    app<uint64_t>(sb, UNDEF);
    app<uint64_t>(sb, 0);
    app<uint64_t>(sb, r.ohdr);
    app<uint32_t>(sb, 1);
    app<uint32_t>(sb, 0);
    app<uint64_t>(sb, r.btree);
    app<uint64_t>(sb, r.heap);
    put(0, sb.data(), sb.size());
    eof_ = end;
    size_ = end;
    dirty_ = false;
  }
};

}  // namespace h5mini
