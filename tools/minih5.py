"""minih5 -- a tiny read-only parser for *classic* HDF5 files (numpy only).

Test infrastructure, not product code.  h5py / libhdf5 are not available in
this image, but every fixture of the reference's regression suite
(``/root/reference/test/prog/fortnet/**.hdf5``) uses only the oldest on-disk
structures: superblock v0, v1 object headers, symbol-table groups (v1 B-tree +
local heap + SNOD nodes), contiguous (or compact) little-endian i4/i8/f8
datasets, fixed-length strings and v1 attribute messages.  This reader
handles exactly that subset and raises on anything else.

The shapes returned are the HDF5 (C-order) shapes, i.e. the Fortran shapes of
the reference reversed: ``coordinates (N,3)``, ``weights (d_out,d_in)`` ...

Usage::

    f = H5File(path)
    f["netstat/bpnn/O-subnetwork/layer1/weights"]      # -> np.ndarray
    f.attrs("netstat/mapping/function1")               # -> dict
    f.keys("fnetdata/dataset")                         # -> list of names
"""
from __future__ import annotations

import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class _Obj:
    """Parsed object header: either a group (has btree/heap) or a dataset."""

    __slots__ = ("addr", "btree", "heap", "shape", "dtype", "data_addr", "data_size",
                 "compact", "attrs", "strsize")

    def __init__(self, addr):
        self.addr = addr
        self.btree = None
        self.heap = None
        self.shape = None
        self.dtype = None
        self.data_addr = None
        self.data_size = None
        self.compact = None
        self.attrs = {}
        self.strsize = None


def _parse_datatype(buf, off):
    """Returns (numpy dtype or ('S', n) or None for unsupported, total size in bytes)."""
    cv = buf[off]
    cls = cv & 0x0F
    bits0 = buf[off + 1]
    size = struct.unpack_from("<I", buf, off + 4)[0]
    if cls == 0:  # fixed-point
        if bits0 & 1:
            raise H5Error("big-endian integers not supported")
        signed = bool(bits0 & 0x08)
        return np.dtype(("<i%d" if signed else "<u%d") % size), size
    if cls == 1:  # floating point
        if bits0 & 1:
            raise H5Error("big-endian floats not supported")
        return np.dtype("<f%d" % size), size
    if cls == 3:  # fixed-length string
        return np.dtype("S%d" % size), size
    return None, size  # e.g. variable-length (class 9): ignored by callers


def _parse_dataspace(buf, off):
    ver = buf[off]
    rank = buf[off + 1]
    if ver == 1:
        dims_off = off + 8
    elif ver == 2:
        dims_off = off + 4
        if buf[off + 3] == 2:  # null dataspace
            return None
    else:
        raise H5Error("dataspace version %d not supported" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, dims_off)) if rank else ()


def _pad8(n):
    return (n + 7) & ~7


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise H5Error("not an HDF5 file: %s" % path)
        if b[8] != 0:
            raise H5Error("superblock version %d not supported" % b[8])
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        self.base = struct.unpack_from("<Q", b, 24)[0]
        root_hdr = struct.unpack_from("<Q", b, 56 + 8)[0]
        self._cache = {}
        self.root = self._object(root_hdr)

    # -- object headers ---------------------------------------------------
    def _object(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        b = self.buf
        a = addr + self.base
        if b[a] != 1:
            raise H5Error("object header version %d not supported" % b[a])
        nmsg = struct.unpack_from("<H", b, a + 2)[0]
        hsize = struct.unpack_from("<I", b, a + 8)[0]
        obj = _Obj(addr)
        blocks = [(a + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            off, length = blocks.pop(0)
            end = off + length
            while off + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, off)
                body = off + 8
                seen += 1
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr + self.base, clen))
                elif mtype == 0x0011:  # symbol table
                    obj.btree, obj.heap = struct.unpack_from("<QQ", b, body)
                elif mtype == 0x0001:
                    obj.shape = _parse_dataspace(b, body)
                elif mtype == 0x0003:
                    obj.dtype, _ = _parse_datatype(b, body)
                elif mtype == 0x0008:
                    ver = b[body]
                    if ver != 3:
                        raise H5Error("layout version %d not supported" % ver)
                    lclass = b[body + 1]
                    if lclass == 1:
                        obj.data_addr, obj.data_size = struct.unpack_from("<QQ", b, body + 2)
                    elif lclass == 0:
                        csize = struct.unpack_from("<H", b, body + 2)[0]
                        obj.compact = bytes(b[body + 4: body + 4 + csize])
                    else:
                        raise H5Error("chunked layout not supported")
                elif mtype == 0x000C:
                    self._attribute(obj, body)
                elif mtype == 0x000B:
                    raise H5Error("filter pipeline not supported")
                off = body + msize
        self._cache[addr] = obj
        return obj

    def _attribute(self, obj, off):
        b = self.buf
        ver = b[off]
        if ver not in (1, 2, 3):
            raise H5Error("attribute version %d not supported" % ver)
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, off + 2)
        p = off + 8
        if ver == 3:
            p += 1  # name character-set encoding
        name = bytes(b[p:p + nsz]).split(b"\0")[0].decode()
        pad = _pad8 if ver == 1 else (lambda n: n)
        p += pad(nsz)
        dtype, _ = _parse_datatype(b, p)
        p += pad(tsz)
        shape = _parse_dataspace(b, p)
        p += pad(ssz)
        if dtype is None or shape is None:
            return  # variable-length strings etc.: not needed by the tests
        n = int(np.prod(shape)) if shape else 1
        raw = np.frombuffer(b, dtype=dtype, count=n, offset=p)
        if dtype.kind == "S":
            val = raw[0].split(b"\0")[0].decode().strip()
        else:
            val = raw.reshape(shape).copy() if shape else raw[0]
            if shape and n == 1:
                val = val.reshape(-1)[0]
        obj.attrs[name] = val

    # -- groups -----------------------------------------------------------
    def _heap_data(self, heap_addr):
        b = self.buf
        a = heap_addr + self.base
        if b[a:a + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        return struct.unpack_from("<Q", b, a + 24)[0] + self.base

    def _children(self, obj):
        if obj.btree is None:
            raise H5Error("object is not a group")
        heap = self._heap_data(obj.heap)
        out = {}
        self._walk_btree(obj.btree, heap, out)
        return out

    def _walk_btree(self, addr, heap, out):
        b = self.buf
        a = addr + self.base
        if b[a:a + 4] != b"TREE":
            raise H5Error("bad B-tree signature")
        level = b[a + 5]
        used = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, p + 8)[0]
            p += 16
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                self._snod(child, heap, out)

    def _snod(self, addr, heap, out):
        b = self.buf
        a = addr + self.base
        if b[a:a + 4] != b"SNOD":
            raise H5Error("bad symbol node signature")
        n = struct.unpack_from("<H", b, a + 6)[0]
        p = a + 8
        for i in range(n):
            noff, haddr = struct.unpack_from("<QQ", b, p)
            s = heap + noff
            e = b.index(b"\0", s)
            out[bytes(b[s:e]).decode()] = haddr
            p += 40

    # -- public API ---------------------------------------------------------
    def _resolve(self, path):
        obj = self.root
        for part in [p for p in path.split("/") if p]:
            ch = self._children(obj)
            if part not in ch:
                raise KeyError(path)
            obj = self._object(ch[part])
        return obj

    def exists(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def keys(self, path=""):
        return sorted(self._children(self._resolve(path)).keys())

    def is_group(self, path):
        return self._resolve(path).btree is not None

    def attrs(self, path=""):
        return dict(self._resolve(path).attrs)

    def __getitem__(self, path):
        obj = self._resolve(path)
        if obj.btree is not None:
            raise H5Error("%s is a group" % path)
        if obj.dtype is None or obj.shape is None:
            raise H5Error("%s: unsupported dataset type" % path)
        n = int(np.prod(obj.shape)) if obj.shape else 1
        if obj.compact is not None:
            raw = np.frombuffer(obj.compact, dtype=obj.dtype, count=n)
        elif obj.data_addr == UNDEF or n == 0:
            raw = np.zeros(n, dtype=obj.dtype)
        else:
            raw = np.frombuffer(self.buf, dtype=obj.dtype, count=n,
                                offset=obj.data_addr + self.base)
        return raw.reshape(obj.shape).copy()

    def walk(self, path=""):
        """Yields (path, kind) for every object below ``path``."""
        for k in self.keys(path):
            p = (path + "/" + k).lstrip("/")
            if self.is_group(p):
                yield p, "group"
                yield from self.walk(p)
            else:
                yield p, "dataset"


if __name__ == "__main__":
    import sys
    f = H5File(sys.argv[1])
    print("/", f.attrs(""))
    for p, kind in f.walk():
        if kind == "group":
            print(p + "/", f.attrs(p))
        else:
            d = f[p]
            print(p, d.dtype, d.shape, f.attrs(p), d.reshape(-1)[:4])
