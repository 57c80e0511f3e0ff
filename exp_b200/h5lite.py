"""h5lite -- a minimal pure-Python HDF5 reader/writer for the reference's basis-table caches
(SURVEY.md section 8f rank 3).

EXP caches the finished basis tables in HDF5 through HighFive: SLGridSph::WriteH5Cache /
ReadH5Cache (reference exputil/SLGridMP2.cc:490-696: root attributes geometry, forceID, Version,
model, lmax, nmax, numr, cmap, rmin, rmax, rmapping, diverge, dfac; groups Harmonic/<l>/{ev, ef})
and EmpCylSL::WriteH5Cache / ReadH5Cache (exputil/EmpCylSL.cc:7378-7680: attributes mmax, numx,
numy, nmax, cmapr, cmapz, rmin, rmax, ascl, hscl, cmass ...; groups Cosine/<m>/<n>/{potC, rforceC,
zforceC, densC} and Sine/<m>/<n>/{potS, ...}).  No HDF5 library is available in this image, so this
module reads the subset of the file format those files use, as libhdf5 writes it by default:
superblock version 0/1 (optionally behind a user block), version-1 object headers, "old style" groups
(symbol-table message -> B-tree v1 + local heap + SNOD nodes), dataspace / datatype / layout /
attribute / continuation / filter messages, contiguous, compact and chunked (B-tree v1, optional
deflate + shuffle) datasets of fixed-point and IEEE floating-point types, fixed-length and
variable-length (global heap) strings.  Version-2 object headers with compact link messages are read
as well; dense (fractal-heap) groups are not.

The writer produces the same old-style layout (superblock 0, contiguous little-endian datasets,
variable-length UTF-8 string attributes as HighFive writes std::string) and exists so that tables
made here can be handed to the reference and for the round-trip tests.  The reader is also checked
against a file written by a real libhdf5 (tests/test_h5lite.py).

Format reference: "HDF5 File Format Specification Version 2.0/3.0" (The HDF Group).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
COMPLEX = np.dtype([("r", "<f8"), ("i", "<f8")])       # HighFive's on-disk std::complex<double>
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


# =============================================================================================
# Reader
# =============================================================================================
class Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters, attrs):
        self._f, self.name, self.shape, self.dtype, self._layout, self._filters = f, name, shape, dtype, layout, filters
        self.attrs = attrs

    def read(self):
        return self._f._read_data(self)

    def __repr__(self):
        return f"<h5lite.Dataset {self.name} {self.shape} {self.dtype}>"


class Group:
    def __init__(self, f, name, links, attrs):
        self._f, self.name, self._links, self.attrs = f, name, links, attrs

    def keys(self):
        return list(self._links)

    def __contains__(self, k):
        return k in self._links

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node

    def __repr__(self):
        return f"<h5lite.Group {self.name} {self.keys()}>"


class File(Group):
    def __init__(self, path_or_bytes):
        self.buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
        base = 0
        while True:                       # the superblock may sit behind a user block: 0, 512, 1024, ...
            if self.buf[base:base + 8] == SIG:
                break
            base = 512 if base == 0 else base * 2
            if base >= len(self.buf):
                raise H5Error("not an HDF5 file")
        self.sb = base
        ver = self.buf[base + 8]
        if ver in (0, 1):
            self.O, self.L = self.buf[base + 13], self.buf[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base = self._uint(p, self.O)
            p += 4 * self.O                     # base, free-space info, end of file, driver info
            # root group symbol table entry
            root_hdr = self._uint(p + self.O, self.O)
        elif ver in (2, 3):
            self.O, self.L = self.buf[base + 9], self.buf[base + 10]
            p = base + 12
            self.base = self._uint(p, self.O)
            root_hdr = self._uint(p + 3 * self.O, self.O)
        else:
            raise H5Error(f"superblock version {ver} not supported")
        self._cache = {}
        root = self._object(root_hdr, "/")
        super().__init__(self, "/", root._links, root.attrs)

    # ---- primitives -------------------------------------------------------------------------
    def _uint(self, p, n):
        return int.from_bytes(self.buf[p:p + n], "little")

    def _addr(self, a):
        return None if a == (1 << (8 * self.O)) - 1 else self.base + a

    # ---- object headers ---------------------------------------------------------------------
    def _messages(self, addr):
        """(type, flags, payload offset, size) of every header message of the object at addr"""
        p = self.base + addr
        out = []
        if self.buf[p:p + 4] == b"OHDR":                     # version 2
            flags = self.buf[p + 5]
            q = p + 6
            if flags & 0x20:
                q += 16
            if flags & 0x10:
                q += 4
            csz = 1 << (flags & 3)
            chunk0 = self._uint(q, csz)
            q += csz
            blocks = [(q, q + chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                a, b = blocks.pop(0)
                while a + 4 <= b - 4 + 4 and a + 4 <= b:
                    mtype, msize, mflags = self.buf[a], self._uint(a + 1, 2), self.buf[a + 3]
                    a += 4 + (2 if track else 0)
                    if a + msize > b:
                        break
                    if mtype == 0x10:
                        ca, cl = self.base + self._uint(a, self.O), self._uint(a + self.O, self.L)
                        blocks.append((ca + 4, ca + cl - 4))      # "OCHK" ... checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, a, msize))
                    a += msize
            return out
        if self.buf[p] != 1:
            raise H5Error(f"object header version {self.buf[p]} at {addr} not supported")
        nmsg = self._uint(p + 2, 2)
        size = self._uint(p + 8, 4)
        blocks = [(p + 16, p + 16 + size)]
        while blocks and len(out) < nmsg + 64:
            a, b = blocks.pop(0)
            while a + 8 <= b:
                mtype, msize, mflags = self._uint(a, 2), self._uint(a + 2, 2), self.buf[a + 4]
                a += 8
                if mtype == 0x10:
                    ca, cl = self.base + self._uint(a, self.O), self._uint(a + self.O, self.L)
                    blocks.append((ca, ca + cl))
                elif mtype != 0:
                    out.append((mtype, mflags, a, msize))
                a += msize
        return out

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs, links = {}, None
        shape = dtype = layout = None
        filters = []
        for mtype, mflags, p, size in msgs:
            if mtype == 0x01:
                shape = self._dataspace(p)
            elif mtype == 0x03:
                dtype, _ = self._datatype(p)
            elif mtype == 0x08:
                layout = self._layout(p)
            elif mtype == 0x0B:
                filters = self._filters(p)
            elif mtype == 0x0C:
                k, v = self._attribute(p)
                attrs[k] = v
            elif mtype == 0x11:
                btree, heap = self._uint(p, self.O), self._uint(p + self.O, self.O)
                links = self._symbol_table(btree, heap)
            elif mtype == 0x06:                                      # compact link (new-style group)
                links = links if links is not None else {}
                k, a = self._link(p)
                if a is not None:
                    links[k] = a
            elif mtype == 0x02 and links is None:
                links = {}
        if links is not None and layout is None:
            obj = Group(self, name, links, attrs)
        elif layout is not None:
            obj = Dataset(self, name, shape, dtype, layout, filters, attrs)
        else:
            obj = Group(self, name, {}, attrs)
        self._cache[addr] = obj
        return obj

    # ---- groups -----------------------------------------------------------------------------
    def _heap_string(self, heap_data, off):
        end = self.buf.index(b"\0", heap_data + off)
        return self.buf[heap_data + off:end].decode("utf-8")

    def _symbol_table(self, btree, heap):
        hp = self.base + heap
        if self.buf[hp:hp + 4] != b"HEAP":
            raise H5Error("bad local heap")
        heap_data = self.base + self._uint(hp + 8 + 2 * self.L, self.O)
        links = {}

        def walk(node):
            p = self.base + node
            if self.buf[p:p + 4] == b"SNOD":
                n = self._uint(p + 6, 2)
                q = p + 8
                for _ in range(n):
                    noff, oaddr = self._uint(q, self.O), self._uint(q + self.O, self.O)
                    links[self._heap_string(heap_data, noff)] = oaddr
                    q += 2 * self.O + 24
                return
            if self.buf[p:p + 4] != b"TREE":
                raise H5Error("bad group B-tree node")
            used = self._uint(p + 6, 2)
            q = p + 8 + 2 * self.O
            for k in range(used):
                q += self.L                                 # key k
                walk(self._uint(q, self.O))
                q += self.O
        walk(btree)
        return links

    def _link(self, p):
        flags = self.buf[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = self.buf[q]; q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        lsz = 1 << (flags & 3)
        n = self._uint(q, lsz); q += lsz
        name = self.buf[q:q + n].decode("utf-8"); q += n
        return name, (self._uint(q, self.O) if ltype == 0 else None)

    # ---- messages ---------------------------------------------------------------------------
    def _dataspace(self, p):
        ver, rank, flags = self.buf[p], self.buf[p + 1], self.buf[p + 2]
        q = p + (8 if ver == 1 else 4)
        if ver == 2 and self.buf[p + 3] == 2:
            return None                                      # null dataspace
        return tuple(self._uint(q + k * self.L, self.L) for k in range(rank))

    def _datatype(self, p):
        """-> (descriptor, bytes consumed); descriptor = numpy dtype, ('S', n) or ('vlen-str',)"""
        cls, ver = self.buf[p] & 0x0F, self.buf[p] >> 4
        b0, b1 = self.buf[p + 1], self.buf[p + 2]
        size = self._uint(p + 4, 4)
        order = ">" if b0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}"), 8 + 4
        if cls == 1:
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:
            return ("S", size), 8
        if cls == 9:
            base, used = self._datatype(p + 8)
            if (b0 & 0x0F) == 1:
                return ("vlen-str",), 8 + used
            return ("vlen", base), 8 + used
        if cls == 6:                                                   # compound -> numpy structured dtype
            nmemb = b0 | (b1 << 8)
            q = p + 8
            names, formats, offsets = [], [], []
            for _ in range(nmemb):
                end = self.buf.index(b"\0", q)
                names.append(self.buf[q:end].decode("utf-8"))
                nlen = end - q + 1
                q += (nlen + 7) // 8 * 8 if ver < 3 else nlen
                if ver == 3:
                    ob = 1 if size < 256 else 2 if size < 65536 else 4
                    offsets.append(self._uint(q, ob)); q += ob
                else:
                    offsets.append(self._uint(q, 4)); q += 4
                    if ver == 1:
                        if self.buf[q]:
                            raise H5Error("compound members with array dimensions are not supported")
                        q += 28                                        # dimensionality, permutation, sizes
                mt, used = self._datatype(q)
                q += used
                if isinstance(mt, np.dtype):
                    formats.append(mt)
                elif mt[0] == "S":
                    formats.append(np.dtype(f"S{mt[1]}"))
                else:
                    raise H5Error(f"compound member of type {mt} not supported")
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), q - p
        raise H5Error(f"datatype class {cls} not supported")

    def _layout(self, p):
        ver = self.buf[p]
        if ver == 3:
            cls = self.buf[p + 1]
            if cls == 0:
                n = self._uint(p + 2, 2)
                return ("compact", p + 4, n)
            if cls == 1:
                return ("contiguous", self._uint(p + 2, self.O), self._uint(p + 2 + self.O, self.L))
            if cls == 2:
                nd = self.buf[p + 2]
                bt = self._uint(p + 3, self.O)
                dims = tuple(self._uint(p + 3 + self.O + 4 * k, 4) for k in range(nd))
                return ("chunked", bt, dims)
        elif ver in (1, 2):
            nd, cls = self.buf[p + 1], self.buf[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._uint(q, self.O); q += self.O
            dims = tuple(self._uint(q + 4 * k, 4) for k in range(nd)); q += 4 * nd
            if cls == 0:
                n = self._uint(q, 4)
                return ("compact", q + 4, n)
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims + (0,) if ver == 1 and False else dims)
        raise H5Error(f"data layout version {ver} not supported")

    def _filters(self, p):
        ver, n = self.buf[p], self.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._uint(q, 2)
            if ver == 1 or fid >= 256:
                nlen = self._uint(q + 2, 2); q += 4
            else:
                nlen = 0; q += 2
            ncd = self._uint(q + 2, 2); q += 4
            if ver == 1:
                nlen = (nlen + 7) // 8 * 8
            q += nlen
            cd = [self._uint(q + 4 * k, 4) for k in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _decode(self, dtype, shape, raw_at, nbytes=None):
        """value of an attribute / compact dataset whose bytes start at raw_at"""
        count = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, np.dtype):
            a = np.frombuffer(self.buf, dtype=dtype, count=count, offset=raw_at).astype(dtype.newbyteorder("="))
            if shape:
                return a.reshape(shape)
            return a[0] if dtype.names else a[0].item()
        if dtype[0] == "S":
            n = dtype[1]
            vals = [self.buf[raw_at + k * n:raw_at + (k + 1) * n].split(b"\0")[0].decode("utf-8") for k in range(count)]
            return vals[0] if not shape else np.array(vals).reshape(shape)
        if dtype[0] == "vlen-str":
            vals = []
            for k in range(count):
                q = raw_at + k * (4 + self.O + 4)
                n, col, idx = self._uint(q, 4), self._uint(q + 4, self.O), self._uint(q + 4 + self.O, 4)
                vals.append(self._global_heap(col, idx)[:n].decode("utf-8"))
            return vals[0] if not shape else np.array(vals).reshape(shape)
        raise H5Error(f"cannot decode {dtype}")

    def _global_heap(self, col, idx):
        p = self.base + col
        if self.buf[p:p + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = self._uint(p + 8, self.L)
        q, end = p + 8 + self.L, p + size
        while q + 8 + self.L <= end:
            oid, osz = self._uint(q, 2), self._uint(q + 8, self.L)
            if oid == 0:
                break
            if oid == idx:
                return self.buf[q + 8 + self.L:q + 8 + self.L + osz]
            q += 8 + self.L + (osz + 7) // 8 * 8
        raise H5Error("global heap object not found")

    def _attribute(self, p):
        ver = self.buf[p]
        nsz, tsz, ssz = self._uint(p + 2, 2), self._uint(p + 4, 2), self._uint(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = self.buf[q:q + nsz].split(b"\0")[0].decode("utf-8"); q += pad(nsz)
        dtype, _ = self._datatype(q); q += pad(tsz)
        shape = self._dataspace(q); q += pad(ssz)
        return name, self._decode(dtype, shape or (), q)

    # ---- data -------------------------------------------------------------------------------
    def _read_data(self, d):
        if not isinstance(d.dtype, np.dtype):
            if d._layout[0] == "compact":
                return self._decode(d.dtype, d.shape, d._layout[1])
            if d._layout[0] == "contiguous":
                return self._decode(d.dtype, d.shape, self.base + d._layout[1])
            raise H5Error("chunked string datasets are not supported")
        count = int(np.prod(d.shape)) if d.shape else 1
        kind = d._layout[0]
        if kind == "compact":
            raw_at = d._layout[1]
        elif kind == "contiguous":
            a = self._addr(d._layout[1])
            if a is None:
                return np.zeros(d.shape, dtype=d.dtype.newbyteorder("="))
            raw_at = a
        else:
            return self._read_chunked(d)
        out = np.frombuffer(self.buf, dtype=d.dtype, count=count, offset=raw_at).astype(d.dtype.newbyteorder("="))
        return out.reshape(d.shape) if d.shape else out[0]

    def _read_chunked(self, d):
        _, bt, cdims = d._layout
        cshape = cdims[:-1]
        out = np.zeros(d.shape, dtype=d.dtype.newbyteorder("="))
        nd = len(d.shape)

        def walk(node):
            p = self.base + node
            if self.buf[p:p + 4] != b"TREE":
                raise H5Error("bad chunk B-tree node")
            level, used = self.buf[p + 5], self._uint(p + 6, 2)
            q = p + 8 + 2 * self.O
            ksz = 8 + 8 * (nd + 1)
            for _ in range(used):
                csize, fmask = self._uint(q, 4), self._uint(q + 4, 4)
                off = tuple(self._uint(q + 8 + 8 * k, 8) for k in range(nd))
                child = self._uint(q + ksz, self.O)
                q += ksz + self.O
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(self.buf[self.base + child:self.base + child + csize])
                for fi, (fid, cd) in reversed(list(enumerate(d._filters))):
                    if fmask & (1 << fi):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else d.dtype.itemsize
                        a = np.frombuffer(raw, dtype=np.uint8)
                        raw = a.reshape(es, -1).T.tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise H5Error(f"filter {fid} not supported")
                chunk = np.frombuffer(raw, dtype=d.dtype, count=int(np.prod(cshape))).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, cshape, d.shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if self._addr(bt) is not None:
            walk(bt)
        return out


# =============================================================================================
# Writer: old-style layout as libhdf5 writes it by default
# =============================================================================================
class Writer:
    """Build a tree with group()/dataset()/attrs, then write(path).  Nodes are plain dicts:
        g = w.root; h = w.group(g, "Harmonic"); w.dataset(h, "ev", array); g["attrs"]["lmax"] = 6"""
    O = L = 8
    LEAF_K, INT_K = 16, 16            # SNOD holds up to 2*LEAF_K symbols

    def __init__(self):
        self.root = self._new_group()

    @staticmethod
    def _new_group():
        return {"kind": "group", "children": {}, "attrs": {}}

    def group(self, parent, name):
        g = self._new_group()
        parent["children"][name] = g
        return g

    def dataset(self, parent, name, array):
        a = np.asarray(array)
        a = np.ascontiguousarray(a).reshape(a.shape)            # ascontiguousarray makes 0-d arrays 1-d
        if a.dtype.kind == "c":                                 # HighFive's std::complex<double>: {r, i}
            a = a.astype(np.complex128).view(COMPLEX).reshape(a.shape)
        if a.dtype.kind not in "fiuV":
            raise H5Error("datasets must be numeric or compound")
        node = {"kind": "dataset", "data": a, "attrs": {}}
        parent["children"][name] = node
        return node

    # ---- serialisation ----------------------------------------------------------------------
    def write(self, path):
        self.out = bytearray(b"\0" * 96)             # superblock v0 (56 + 40-byte root entry)
        self.gheap = []                              # variable-length strings of attributes
        self.gheap_fixups = []                       # (position of the 16-byte vlen reference, object index)
        root_addr, root_bt, root_heap = self._emit(self.root)
        gaddr = self._emit_gheap()
        for pos, idx in self.gheap_fixups:
            struct.pack_into("<Q", self.out, pos + 4, gaddr)
        eof = len(self.out)
        sb = bytearray()
        sb += SIG + bytes([0, 0, 0, 0, 0, self.O, self.L, 0])
        sb += struct.pack("<HHI", self.LEAF_K, self.INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", root_bt, root_heap)
        self.out[0:len(sb)] = sb
        with open(path, "wb") as f:
            f.write(bytes(self.out))

    def _align(self):
        while len(self.out) % 8:
            self.out += b"\0"

    def _alloc(self, data):
        self._align()
        a = len(self.out)
        self.out += data
        return a

    @staticmethod
    def _dtype_msg(dt):
        dt = np.dtype(dt)
        if dt.kind == "V" and dt.names:                      # compound, version 1 (libhdf5's default)
            body = b""
            for nm in dt.names:
                mt, off = dt.fields[nm][0], dt.fields[nm][1]
                b = nm.encode() + b"\0"
                body += b + b"\0" * (-len(b) % 8) + struct.pack("<I", off) + b"\0" * 28 + Writer._dtype_msg(mt)
            n = len(dt.names)
            return struct.pack("<BBBBI", 0x16, n & 0xFF, n >> 8, 0, dt.itemsize) + body
        if dt.kind == "S":                                   # fixed-length, null-terminated, ASCII
            return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)
        if dt.kind == "f":
            if dt.itemsize == 8:
                props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + props
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + props
        b0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, b0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)

    def _space_msg(self, shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)

    def _attr_msg(self, name, value):
        pad = lambda b: b + b"\0" * (-len(b) % 8)
        nm = name.encode() + b"\0"
        if isinstance(value, str):
            # variable-length UTF-8 string (what HighFive writes for std::string): class 9, type 1
            # (string), null-terminated, UTF-8; base type = 1-byte fixed-length string
            base = struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)
            dt = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base
            sp = self._space_msg(())
            raw = value.encode("utf-8")
            self.gheap.append(raw)
            idx = len(self.gheap)
            data = struct.pack("<IQI", len(raw), 0, idx)
            fix = ("vlen", idx)
        else:
            a = np.asarray(value)
            if a.dtype.kind == "i":
                a = a.astype(np.int32)
            elif a.dtype.kind == "f":
                a = a.astype(np.float64)
            dt = self._dtype_msg(a.dtype)
            sp = self._space_msg(a.shape)
            data = a.tobytes()
            fix = None
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + pad(nm) + pad(dt) + pad(sp)
        return body, data, fix

    def _header(self, msgs):
        """version-1 object header from [(type, payload, fixups)] -> address; fixups = [(offset in
        payload, ('vlen', idx))]"""
        body = bytearray()
        fix = []
        for mtype, payload, fx in msgs:
            payload = payload + b"\0" * (-len(payload) % 8)
            body += struct.pack("<HHB3x", mtype, len(payload), 0)
            for off, what in fx:
                fix.append((len(body) + off, what))
            body += payload
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        addr = self._alloc(hdr + bytes(body))
        for off, what in fix:
            self.gheap_fixups.append((addr + 16 + off, what[1]))
        return addr

    def _attr_msgs(self, attrs):
        out = []
        for k, v in attrs.items():
            body, data, fx = self._attr_msg(k, v)
            out.append((0x0C, body + data, [(len(body), fx)] if fx else []))
        return out

    def _emit(self, node):
        if node["kind"] == "dataset":
            a = node["data"]
            daddr = self._alloc(a.astype(a.dtype.newbyteorder("<")).tobytes()) if a.size else UNDEF
            layout = struct.pack("<BBQQ", 3, 1, daddr, a.nbytes)
            msgs = [(0x01, self._space_msg(a.shape), []), (0x03, self._dtype_msg(a.dtype), []),
                    (0x05, struct.pack("<BBBB", 2, 2, 2, 0), []),       # fill value (v2): late alloc, undefined
                    (0x08, layout, [])] + self._attr_msgs(node["attrs"])
            return self._header(msgs), None, None
        # group: children first, then heap + SNOD + B-tree + header
        entries = []
        for name in sorted(node["children"], key=lambda s: s.encode()):
            caddr, cbt, cheap = self._emit(node["children"][name])
            entries.append((name, caddr, cbt, cheap))
        heap = bytearray(b"\0" * 8)                      # offset 0: the empty string
        offs = []
        for name, *_ in entries:
            offs.append(len(heap))
            b = name.encode() + b"\0"
            heap += b + b"\0" * (-len(b) % 8)
        free = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\0" * 16   # one free block at the end
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free, heap_data))
        # symbol nodes of up to 2*LEAF_K entries under a B-tree (version 1, node type 0) whose nodes
        # hold up to 2*INT_K children; key i of a node = heap offset of the largest name in child i-1
        # (key 0 = the empty string)
        nodes = []                                       # (address, heap offset of the largest name below)
        for k0 in range(0, len(entries), 2 * self.LEAF_K):
            part = list(zip(entries, offs))[k0:k0 + 2 * self.LEAF_K]
            snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for (name, caddr, cbt, cheap), off in part:
                if cbt is not None:
                    snod += struct.pack("<QQII", off, caddr, 1, 0) + struct.pack("<QQ", cbt, cheap)
                else:
                    snod += struct.pack("<QQII", off, caddr, 0, 0) + b"\0" * 16
            snod += b"\0" * (8 + 2 * self.LEAF_K * 40 - len(snod))
            nodes.append((self._alloc(bytes(snod)), part[-1][1]))
        level = 0
        node_size = 24 + (2 * self.INT_K + 1) * 8 + 2 * self.INT_K * 8
        while True:
            parents = []
            for k0 in range(0, max(len(nodes), 1), 2 * self.INT_K):
                part = nodes[k0:k0 + 2 * self.INT_K]
                bt = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(part), UNDEF, UNDEF))
                bt += struct.pack("<Q", 0)
                for addr, last in part:
                    bt += struct.pack("<QQ", addr, last)
                bt += b"\0" * (node_size - len(bt))
                parents.append([self._alloc(bytes(bt)), part[-1][1] if part else 0])
            for a, b in zip(parents, parents[1:]):       # sibling links of this level
                struct.pack_into("<Q", self.out, a[0] + 16, b[0])
                struct.pack_into("<Q", self.out, b[0] + 8, a[0])
            nodes = [tuple(x) for x in parents]
            level += 1
            if len(nodes) == 1:
                break
        bt_addr = nodes[0][0]
        msgs = [(0x11, struct.pack("<QQ", bt_addr, heap_addr), [])] + self._attr_msgs(node["attrs"])
        return self._header(msgs), bt_addr, heap_addr

    def _emit_gheap(self):
        if not self.gheap:
            return 0
        body = bytearray()
        for k, raw in enumerate(self.gheap, 1):
            body += struct.pack("<HHIQ", k, 1, 0, len(raw)) + raw + b"\0" * (-len(raw) % 8)
        size = max(4096, 16 + len(body) + 16)
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free)          # object 0: the free space
        body += b"\0" * (size - 16 - len(body))
        return self._alloc(b"GCOL" + struct.pack("<B3xQ", 1, size) + bytes(body))


# =============================================================================================
# The reference's cache files
# =============================================================================================
def read_slgrid_cache(path):
    """SLGridSph cache (exputil/SLGridMP2.cc:490-617) -> dict(attrs..., ev[(lmax+1)][nmax],
    ef[(lmax+1)][nmax][numr]).  `rmapping` is called `scale` in older files (:573-577)."""
    f = File(path)
    a = dict(f.attrs)
    if a.get("geometry") != "sphere" or a.get("forceID") != "SLGridSph":
        raise H5Error(f"not an SLGridSph cache: geometry={a.get('geometry')} forceID={a.get('forceID')}")
    lmax, nmax, numr = int(a["lmax"]), int(a["nmax"]), int(a["numr"])
    ev = np.zeros((lmax + 1, nmax))
    ef = np.zeros((lmax + 1, nmax, numr))
    for l in range(lmax + 1):
        g = f[f"Harmonic/{l}"]
        ev[l] = np.asarray(g["ev"].read()).reshape(nmax)
        e = np.asarray(g["ef"].read())
        ef[l] = e if e.shape == (nmax, numr) else e.reshape(numr, nmax).T
    a["rmapping"] = a.get("rmapping", a.get("scale"))
    return dict(a, ev=ev, ef=ef)


def write_slgrid_cache(path, tab, model="SLGridSph.model", version="1.0", diverge=0, dfac=1.0):
    """the file SLGridSph::WriteH5Cache writes (exputil/SLGridMP2.cc:620-696) for a table image"""
    w = Writer()
    r = w.root["attrs"]
    r["geometry"], r["forceID"], r["Version"], r["model"] = "sphere", "SLGridSph", version, model
    r["lmax"], r["nmax"], r["numr"], r["cmap"] = int(tab.lmax), int(tab.nmax), int(tab.numr), int(tab.cmap)
    r["rmin"], r["rmax"], r["rmapping"] = float(tab.rmin), float(tab.rmax), float(tab.rmap)
    r["diverge"], r["dfac"] = int(diverge), float(dfac)
    h = w.group(w.root, "Harmonic")
    for l in range(tab.lmax + 1):
        g = w.group(h, str(l))
        w.dataset(g, "ev", np.asarray(tab.ev[l], dtype=np.float64))
        w.dataset(g, "ef", np.asarray(tab.ef[l], dtype=np.float64))
    w.write(path)


def read_cylinder_cache(path):
    """EmpCylSL cache (exputil/EmpCylSL.cc:7378-7680) -> dict(attrs..., potC/potS/rforceC/rforceS/
    zforceC/zforceS [(mmax+1)][nmax][numx+1][numy+1] (S arrays zero at m = 0))"""
    f = File(path)
    a = dict(f.attrs)
    if a.get("geometry") != "cylinder" or a.get("forceID") != "Cylinder":
        raise H5Error("not an EmpCylSL cache")
    M, N, NX, NY = int(a["mmax"]), int(a["nmax"]), int(a["numx"]), int(a["numy"])
    out = {k: np.zeros((M + 1, N, NX + 1, NY + 1)) for k in ("potC", "potS", "rforceC", "rforceS", "zforceC", "zforceS")}
    for m in range(M + 1):
        for n in range(N):
            g = f[f"Cosine/{m}/{n}"]
            for k in ("potC", "rforceC", "zforceC"):
                out[k][m, n] = np.asarray(g[k].read()).reshape(NX + 1, NY + 1)
            if m:
                g = f[f"Sine/{m}/{n}"]
                for k in ("potS", "rforceS", "zforceS"):
                    out[k][m, n] = np.asarray(g[k].read()).reshape(NX + 1, NY + 1)
    return dict(a, **out)


def write_cylinder_cache(path, tab, model="Exponential", version="1.0", cmass=0.0, dens=None):
    """the layout EmpCylSL::WriteH5Cache writes (exputil/EmpCylSL.cc:7378-7470); densC/densS are
    written as zeros unless given (the force path does not read them)"""
    w = Writer()
    r = w.root["attrs"]
    r["geometry"], r["forceID"], r["Version"], r["model"] = "cylinder", "Cylinder", version, model
    for k, v in (("mmax", tab.mmax), ("numx", tab.numx), ("numy", tab.numy), ("nmax", tab.norder), ("lmaxfid", 0),
                 ("nmaxfid", 0), ("neven", 0), ("nodd", 0), ("cmapr", tab.cmapr), ("cmapz", tab.cmapz)):
        r[k] = int(v)
    for k, v in (("rmin", tab.rmin), ("rmax", tab.rmax), ("ascl", tab.ascale), ("hscl", tab.hscale), ("cmass", cmass)):
        r[k] = float(v)
    zero = np.zeros((tab.numx + 1, tab.numy + 1))
    cos, sin = w.group(w.root, "Cosine"), w.group(w.root, "Sine")
    for m in range(tab.mmax + 1):
        gm = w.group(cos, str(m))
        gs = w.group(sin, str(m)) if m else None
        for n in range(tab.norder):
            g = w.group(gm, str(n))
            for k in ("potC", "rforceC", "zforceC"):
                w.dataset(g, k, getattr(tab, k)[m, n])
            w.dataset(g, "densC", zero if dens is None else dens[0][m, n])
            if m:
                g = w.group(gs, str(n))
                for k in ("potS", "rforceS", "zforceS"):
                    w.dataset(g, k, getattr(tab, k)[m, n])
                w.dataset(g, "densS", zero if dens is None else dens[1][m, n])
    w.write(path)


# =============================================================================================
# The reference's coefficient files (CoefClasses::SphCoefs / CylCoefs through HighFive)
# =============================================================================================
UNIT = np.dtype([("name", "S16"), ("unit", "S16"), ("value", "<f4")])    # expui/Coefficients.H:36-55
DEFAULT_UNITS = (("length", "none", 1.0), ("mass", "none", 1.0), ("time", "none", 1.0))


def _units_array(units):
    u = np.zeros(len(units), dtype=UNIT)
    for k, (n, t, v) in enumerate(units):
        u[k] = (n.encode()[:15], t.encode()[:15], v)
    return u


def _tree_of(node, name="/"):
    """reader object -> Writer node (the whole file in memory; coefficient files are small)"""
    if isinstance(node, Dataset):
        return {"kind": "dataset", "data": np.asarray(node.read()), "attrs": dict(node.attrs)}
    out = {"kind": "group", "children": {}, "attrs": dict(node.attrs)}
    for k in node.keys():
        out["children"][k] = _tree_of(node[k], k)
    return out


def _add_snapshot(w, snaps, count, time, center, rotation, coefs):
    """one stanza of {Sph,Cyl}Coefs::WriteH5Times (expui/Coefficients.cc:907-940, 1375-1402): group
    "%08d" with attributes Time, Center[3], Rotation[3][3] and the complex dataset "coefficients"
    (rows, nmax), row-major on disk, each element the compound {r, i}"""
    s = w.group(snaps, f"{count:08d}")
    s["attrs"]["Time"] = float(time)
    s["attrs"]["Center"] = np.asarray(center, dtype=np.float64).reshape(3)
    s["attrs"]["Rotation"] = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
    w.dataset(s, "coefficients", np.asarray(coefs, dtype=np.complex128))
    return count + 1


def dump_coefs_h5(path, geometry, name, forceID, time, coefs, *, lmax=None, mmax=None, scale=1.0, config="",
                  center=(0.0, 0.0, 0.0), rotation=None, units=DEFAULT_UNITS):
    """SphericalBasis::dump_coefs_h5 / Cylinder::dump_coefs_h5 (reference src/SphericalBasis.cc:
    1909-1975, src/Cylinder.cc:1625-1689): create the coefficient file on the first call
    (Coefs::WriteH5Coefs, expui/Coefficients.cc:3100-3163: root attributes CoefficientOutputVersion,
    geometry, name, config + the basis parameters of WriteH5Params -- lmax, nmax, scale, forceID for
    a sphere (:841-852), mmax, nmax, forceID for a cylinder (:1323-1332) --, the compound dataset
    "Units", the scalar dataset "count", the group "snapshots"), and append one snapshot to it
    afterwards (Coefs::ExtendH5Coefs, :3165-3205, after the same parameter check as CheckH5Params).
    coefs: complex (rows, nmax), cosine in the real and sine in the imaginary part."""
    import os
    coefs = np.asarray(coefs, dtype=np.complex128)
    rows, nmax = coefs.shape
    rot = np.eye(3) if rotation is None else rotation
    if geometry == "sphere":
        if lmax is None or rows != (lmax + 1) * (lmax + 2) // 2:
            raise H5Error("sphere coefficients must be ((lmax+1)(lmax+2)/2, nmax)")
        params = {"lmax": int(lmax), "nmax": int(nmax), "scale": float(scale), "forceID": forceID}
    elif geometry == "cylinder":
        if mmax is None or rows != mmax + 1:
            raise H5Error("cylinder coefficients must be (mmax+1, nmax)")
        params = {"mmax": int(mmax), "nmax": int(nmax), "forceID": forceID}
    else:
        raise H5Error(f"geometry {geometry!r} is outside the force loop's scope")
    w = Writer()
    if os.path.exists(path):
        w.root = _tree_of(File(path))
        a = w.root["attrs"]
        for k, v in params.items():                  # {Sph,Cyl}Coefs::CheckH5Params
            old = a.get(k)
            bad = abs(old - v) > 1.0e-8 * max(abs(old), abs(v)) if k == "scale" else old != v
            if old is None or bad:
                raise H5Error(f"dump_coefs_h5: {k} mismatch {v!r} != {old!r}, not extending {path}")
        count = int(w.root["children"]["count"]["data"])
        count = _add_snapshot(w, w.root["children"]["snapshots"], count, time, center, rot, coefs)
    else:
        a = w.root["attrs"]
        a["CoefficientOutputVersion"], a["geometry"], a["name"], a["config"] = "1.0", geometry, name, config
        a.update(params)
        w.dataset(w.root, "Units", _units_array(units))
        snaps = w.group(w.root, "snapshots")
        count = _add_snapshot(w, snaps, 0, time, center, rot, coefs)
    w.dataset(w.root, "count", np.uint32(count))
    tmp = path + ".tmp"
    w.write(tmp)
    os.replace(tmp, path)
    return count


def read_coefs_h5(path, stride=1, tmin=-np.inf, tmax=np.inf):
    """the H5 constructors of CoefClasses::SphCoefs / CylCoefs (expui/Coefficients.cc:225-332,
    1080-1180): -> dict(geometry, name, forceID, config, units, lmax|mmax, nmax, [scale,] times[T],
    centers[T][3], rotations[T][3][3], coefs[T][rows][nmax] complex).  Files without the
    CoefficientOutputVersion attribute hold the legacy (column-major) storage order and are
    re-ordered the way the reference does (:300-310)."""
    f = File(path)
    a = dict(f.attrs)
    geometry = a.get("geometry")
    if geometry not in ("sphere", "cylinder"):
        raise H5Error(f"coefficient file of geometry {geometry!r} not supported")
    legacy = "CoefficientOutputVersion" not in a
    out = {k: a.get(k) for k in ("geometry", "name", "forceID", "config")}
    nmax = int(a["nmax"])
    if geometry == "sphere":
        out["lmax"], out["scale"] = int(a["lmax"]), float(a["scale"])
        rows = (out["lmax"] + 1) * (out["lmax"] + 2) // 2
    else:
        out["mmax"] = int(a["mmax"])
        rows = out["mmax"] + 1
    out["nmax"] = nmax
    out["units"] = [(u["name"].decode(), u["unit"].decode(), float(u["value"])) for u in f["Units"].read()] \
        if "Units" in f else []
    count = int(f["count"].read())
    snaps = f["snapshots"]
    times, ctrs, rots, cfs = [], [], [], []
    for n in range(0, count, stride):
        s = snaps[f"{n:08d}"]
        t = float(s.attrs["Time"])
        if t < tmin or t > tmax:
            continue
        d = np.asarray(s["coefficients"].read())
        c = d["r"] + 1j * d["i"] if d.dtype.names else d
        if c.shape != (rows, nmax):
            continue                                  # the reference skips stanzas of the wrong shape
        if legacy:
            c = c.reshape(-1).reshape((rows, nmax), order="F")
        times.append(t)
        ctrs.append(np.asarray(s.attrs.get("Center", np.zeros(3)), dtype=float).reshape(-1)[:3])
        rots.append(np.asarray(s.attrs.get("Rotation", np.eye(3)), dtype=float).reshape(3, 3))
        cfs.append(c)
    out["times"] = np.array(times)
    out["centers"] = np.array(ctrs).reshape(-1, 3)
    out["rotations"] = np.array(rots).reshape(-1, 3, 3)
    out["coefs"] = np.array(cfs).reshape(-1, rows, nmax)
    return out
