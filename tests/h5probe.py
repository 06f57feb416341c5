"""TEST INFRASTRUCTURE: an independent reader, in Python, of the HDF5 subset H5part files use (the "classic" layout every
libhdf5 since 1.0 reads and H5hut writes by default): superblock version 0/1, version-1 object headers with
continuation blocks, symbol-table groups (version-1 B-tree of type 0 + local heap + SNOD nodes), contiguous / compact /
chunked-without-filter dataset layouts, fixed-point and IEEE floating-point datatypes, fixed-length strings, version-1/2/3
attributes.  Follows the published "HDF5 File Format Specification Version 2.0" (sections II.A superblock, III.A B-trees,
III.B symbol nodes, III.D local heaps, IV.A object headers and messages).

It serves two purposes: it pins the C++ reader/writer of flecsph_b200/host/h5mini against the reference's own golden
file (mpisph/test/io_test.h5part, written by the real library, copied as tests/golden/io_test.h5part), and it re-reads
what the C++ writer produces (two independent implementations of the format have to agree)."""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class Dataset:
    def __init__(self, name):
        self.name, self.shape, self.dtype, self.layout, self.data, self.attrs = name, None, None, None, None, {}


class Group:
    def __init__(self, name):
        self.name, self.groups, self.datasets, self.attrs = name, {}, {}, {}


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.stats = {"ohdr_versions": set(), "messages": set(), "layouts": set(), "attr_versions": set()}
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        self.sb_version = b[8]
        if self.sb_version not in (0, 1):
            raise H5Error("superblock version %d not supported" % self.sb_version)
        self.so, self.sl = b[13], b[14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        p = 24 if self.sb_version == 0 else 28
        self.base, self.freespace, self.eof, self.driver = struct.unpack_from("<QQQQ", b, p)
        p += 32
        self.root_entry = self._symbol_entry(p)
        self.root = self._load_object(self.root_entry["ohdr"], "/")

    # -- primitives --------------------------------------------------------------------------------------------
    def _symbol_entry(self, p):
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", self.b, p)
        e = {"name_off": name_off, "ohdr": ohdr, "cache": cache}
        if cache == 1:
            e["btree"], e["heap"] = struct.unpack_from("<QQ", self.b, p + 24)
        return e

    def _heap_data(self, addr):
        b = self.b
        if b[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature at %d" % addr)
        size, free, data = struct.unpack_from("<QQQ", b, addr + 8)
        return data, size

    def _heap_string(self, heap_addr, off):
        data, _ = self._heap_data(heap_addr)
        e = self.b.index(b"\0", data + off)
        return self.b[data + off:e].decode()

    def _btree_entries(self, addr, heap):
        """Symbol-table entries below a version-1 B-tree node of type 0, in key order."""
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree signature at %d" % addr)
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise H5Error("group B-tree expected")
        p = addr + 8 + 16   # left / right siblings
        out = []
        for i in range(used):
            child = struct.unpack_from("<Q", b, p + 8)[0]   # key i, child i
            p += 16
            if level > 0:
                out += self._btree_entries(child, heap)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol node signature at %d" % child)
                n = struct.unpack_from("<H", b, child + 6)[0]
                for k in range(n):
                    e = self._symbol_entry(child + 8 + 40 * k)
                    e["name"] = self._heap_string(heap, e["name_off"])
                    out.append(e)
        return out

    # -- object headers -----------------------------------------------------------------------------------------
    def _messages(self, addr):
        b = self.b
        version = b[addr]
        self.stats["ohdr_versions"].add(version)
        if version != 1:
            raise H5Error("object header version %d not supported" % version)
        nmsg, refcount, hsize = struct.unpack_from("<HII", b, addr + 2)
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                self.stats["messages"].add(mtype)
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, ln))
                msgs.append((mtype, flags, body))
        return msgs

    @staticmethod
    def _datatype(body):
        cls, ver = body[0] & 15, body[0] >> 4
        bits0, bits1, bits2 = body[1], body[2], body[3]
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(order + ("i" if bits0 & 8 else "u") + str(size)), 8 + 4
        if cls == 1:
            return np.dtype(order + "f" + str(size)), 8 + 12
        if cls == 3:
            return np.dtype("S%d" % size), 8
        raise H5Error("datatype class %d not supported" % cls)

    @staticmethod
    def _dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        p = 8 if ver == 1 else 4
        if ver == 2 and body[3] == 2:
            return None, p   # null dataspace
        dims = struct.unpack_from("<%dQ" % rank, body, p)
        p += 8 * rank
        if flags & 1:
            p += 8 * rank
        return tuple(dims), p

    def _attribute(self, body):
        ver = body[0]
        self.stats["attr_versions"].add(ver)
        nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
        p = 8 if ver < 3 else 9
        pad = (lambda v: (v + 7) & ~7) if ver == 1 else (lambda v: v)
        name = body[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt, _ = self._datatype(body[p:p + tsz])
        p += pad(tsz)
        shape, _ = self._dataspace(body[p:p + ssz])
        p += pad(ssz)
        cnt = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(body, dt, cnt, p)
        if dt.kind == "S":
            return name, val[0].split(b"\0")[0].decode()
        return name, (val.reshape(shape) if shape else val[0])

    def _load_object(self, addr, name):
        msgs = self._messages(addr)
        types = [m[0] for m in msgs]
        attrs = dict(self._attribute(body) for t, _, body in msgs if t == 0x0C)
        if 0x11 in types:   # group
            g = Group(name)
            g.attrs = attrs
            body = [m[2] for m in msgs if m[0] == 0x11][0]
            btree, heap = struct.unpack_from("<QQ", body, 0)
            for e in self._btree_entries(btree, heap):
                child = self._load_object(e["ohdr"], e["name"])
                (g.groups if isinstance(child, Group) else g.datasets)[e["name"]] = child
            return g
        d = Dataset(name)
        d.attrs = attrs
        for t, _, body in msgs:
            if t == 0x01:
                d.shape, _ = self._dataspace(body)
            elif t == 0x03:
                d.dtype, _ = self._datatype(body)
            elif t == 0x08:
                d.layout = body
        if d.layout is None or d.dtype is None or d.shape is None:
            raise H5Error("object %s is neither a group nor a dataset this reader understands" % name)
        d.data = self._read_layout(d)
        return d

    def _read_layout(self, d):
        body, b = d.layout, self.b
        ver = body[0]
        cnt = int(np.prod(d.shape)) if d.shape else 1
        nbytes = cnt * d.dtype.itemsize
        if ver == 3:
            cls = body[1]
            self.stats["layouts"].add(("v3", cls))
            if cls == 0:
                size = struct.unpack_from("<H", body, 2)[0]
                return np.frombuffer(body, d.dtype, cnt, 4).reshape(d.shape)
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", body, 2)
                if addr == UNDEF:
                    return np.zeros(d.shape, d.dtype)
                return np.frombuffer(b, d.dtype, cnt, addr).reshape(d.shape)
            if cls == 2:
                rank = body[2]
                btree = struct.unpack_from("<Q", body, 3)[0]
                cdims = struct.unpack_from("<%dI" % rank, body, 11)
                return self._read_chunked(d, btree, cdims[:-1])
        elif ver in (1, 2):
            rank, cls = body[1], body[2]
            self.stats["layouts"].add(("v%d" % ver, cls))
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", body, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % rank, body, p)
            p += 4 * rank
            if cls == 1:
                return np.frombuffer(b, d.dtype, cnt, addr).reshape(d.shape)
            if cls == 2:
                return self._read_chunked(d, addr, dims[:-1])
            if cls == 0:
                return np.frombuffer(body, d.dtype, cnt, p + 4).reshape(d.shape)
        raise H5Error("layout version %d not supported" % ver)

    def _read_chunked(self, d, btree, cdims):
        out = np.zeros(d.shape, d.dtype)
        rank = len(d.shape)

        def walk(addr):
            b = self.b
            if b[addr:addr + 4] != b"TREE":
                raise H5Error("bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            if ntype != 1:
                raise H5Error("chunk B-tree expected")
            p = addr + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<%dQ" % (rank + 1), b, p + 8)
                child = struct.unpack_from("<Q", b, p + ksz)[0]
                p += ksz + 8
                if level > 0:
                    walk(child)
                else:
                    if fmask != 0 and csize != int(np.prod(cdims)) * d.dtype.itemsize:
                        raise H5Error("filtered chunks are not supported")
                    chunk = np.frombuffer(b, d.dtype, int(np.prod(cdims)), child).reshape(cdims)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:rank], cdims, d.shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if btree != UNDEF:
            walk(btree)
        return out


def describe(path):
    f = File(path)
    lines = ["superblock v%d, leaf K %d, internal K %d, eof %d" % (f.sb_version, f.leaf_k, f.internal_k, f.eof)]

    def rec(g, ind):
        for k, v in g.attrs.items():
            lines.append("%s@%s = %r" % (ind, k, v))
        for k, d in g.datasets.items():
            lines.append("%s%s %s %s %s" % (ind, k, d.dtype, d.shape, " ".join("@%s=%r" % kv for kv in d.attrs.items())))
        for k, sub in g.groups.items():
            lines.append("%s%s/" % (ind, k))
            rec(sub, ind + "  ")
    rec(f.root, "")
    lines.append("stats: %r" % f.stats)
    return "\n".join(lines)


if __name__ == "__main__":
    import sys
    print(describe(sys.argv[1]))
