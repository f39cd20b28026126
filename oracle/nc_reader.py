"""Minimal NetCDF-4/HDF5 reader for the reference's tiny test files.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  No HDF5/netCDF library is
installed in the build container, and the reference's golden vectors live in
NetCDF-4 files (`/root/reference/test/testdata/*.nc`, used by
`test/xmhw_fixtures.py:40-66`).  Those files are small and regular: superblock
v0, version-2 object headers, datasets either contiguous or a single
shuffle+deflate chunk indexed by a v1 B-tree.  This walker decodes exactly
that subset with the standard library (`zlib`, `struct`) + numpy.

It is used only by `oracle/make_golden.py`, which converts the arrays to
`tests/golden/*.npz` so that nothing at test time needs `/root/reference`.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Hdf5:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n" or b[8] != 0:
            raise ValueError("only HDF5 superblock v0 is supported")
        if b[13] != 8 or b[14] != 8:
            raise ValueError("only 8-byte offsets/lengths are supported")
        # root group symbol-table entry starts at byte 56; header address 8 bytes in
        self.root = struct.unpack_from("<Q", b, 56 + 8)[0]

    # -- version-2 object header ------------------------------------------
    def messages(self, addr):
        """Yield (type, flags, payload bytes) of every header message, following
        continuation blocks (OCHK)."""
        b = self.b
        if b[addr:addr + 4] != b"OHDR" or b[addr + 4] != 2:
            raise ValueError("only version-2 object headers are supported")
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16  # access/mod/change/birth times
        if flags & 0x10:
            p += 4   # max compact / min dense attrs
        szbytes = 1 << (flags & 3)
        chunk0 = int.from_bytes(b[p:p + szbytes], "little")
        p += szbytes
        track_order = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            start, size = blocks.pop(0)
            q, end = start, start + size
            while q + 4 <= end:
                mtype = b[q]
                msize = struct.unpack_from("<H", b, q + 1)[0]
                mflags = b[q + 3]
                q += 4
                if track_order:
                    q += 2
                payload = b[q:q + msize]
                q += msize
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", payload, 0)
                    if b[caddr:caddr + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    blocks.append((caddr + 4, clen - 8))  # minus signature, checksum
                elif mtype != 0:
                    yield mtype, mflags, payload

    def links(self, addr):
        out = {}
        for mtype, _, pl in self.messages(addr):
            if mtype != 0x06:
                continue
            ver, fl = pl[0], pl[1]
            p = 2
            ltype = 0
            if fl & 0x08:
                ltype = pl[p]
                p += 1
            if fl & 0x04:
                p += 8  # creation order
            if fl & 0x10:
                p += 1  # charset
            lsz = 1 << (fl & 3)
            nlen = int.from_bytes(pl[p:p + lsz], "little")
            p += lsz
            name = pl[p:p + nlen].decode()
            p += nlen
            if ltype == 0:
                out[name] = struct.unpack_from("<Q", pl, p)[0]
        return out

    # -- datasets ----------------------------------------------------------
    def dataset(self, addr):
        shape = dtype = None
        layout = None
        filters = []
        for mtype, _, pl in self.messages(addr):
            if mtype == 0x01:  # dataspace
                ver, rank, fl = pl[0], pl[1], pl[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, pl, p)
            elif mtype == 0x03:  # datatype
                cls = pl[0] & 0x0F
                size = struct.unpack_from("<I", pl, 4)[0]
                big = pl[1] & 1
                if cls == 1:
                    dtype = np.dtype("%sf%d" % (">" if big else "<", size))
                elif cls == 0:
                    signed = (pl[1] >> 3) & 1
                    dtype = np.dtype("%s%s%d" % (">" if big else "<", "i" if signed else "u", size))
                else:
                    dtype = None  # strings/compound/etc.: not needed
            elif mtype == 0x08:  # layout
                ver, cls = pl[0], pl[1]
                if ver != 3:
                    raise ValueError("only layout v3 supported")
                if cls == 1:
                    a, s = struct.unpack_from("<QQ", pl, 2)
                    layout = ("contiguous", a, s)
                elif cls == 2:
                    rank = pl[2]
                    btree = struct.unpack_from("<Q", pl, 3)[0]
                    dims = struct.unpack_from("<%dI" % rank, pl, 11)
                    layout = ("chunked", btree, dims)
                elif cls == 0:
                    s = struct.unpack_from("<H", pl, 2)[0]
                    layout = ("compact", pl[4:4 + s])
            elif mtype == 0x0B:  # filter pipeline
                ver, nf = pl[0], pl[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = struct.unpack_from("<H", pl, p)[0]
                    p += 2
                    nlen = 0
                    if ver == 1 or fid >= 256:
                        nlen = struct.unpack_from("<H", pl, p)[0]
                        p += 2
                    p += 2  # flags
                    ncd = struct.unpack_from("<H", pl, p)[0]
                    p += 2
                    if ver == 1:
                        nlen = (nlen + 7) // 8 * 8
                    p += nlen
                    cd = struct.unpack_from("<%dI" % ncd, pl, p)
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            a, s = layout[1], layout[2]
            if a == UNDEF:
                return None
            raw = self.b[a:a + n * dtype.itemsize]
        elif layout[0] == "compact":
            raw = layout[1]
        else:
            raw = self._single_chunk(layout[1], len(shape))
            for fid, cd in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle
                    es = cd[0]
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                elif fid == 3:  # fletcher32 checksum trailer
                    raw = raw[:-4]
                else:
                    raise ValueError("unsupported filter %d" % fid)
        return np.frombuffer(raw[:n * dtype.itemsize], dtype).reshape(shape).astype(dtype.newbyteorder("="))

    def _single_chunk(self, addr, rank):
        b = self.b
        if addr == UNDEF:
            raise ValueError("chunked dataset without storage")
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise ValueError("expected a v1 chunk B-tree")
        level, nent = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        if level != 0 or nent != 1:
            raise ValueError("only single-chunk datasets are supported")
        p = addr + 24
        csize = struct.unpack_from("<I", b, p)[0]
        p += 8 + 8 * (rank + 1)
        caddr = struct.unpack_from("<Q", b, p)[0]
        return b[caddr:caddr + csize]


def read_nc(path):
    """Return {variable name: numpy array} for every numeric dataset in the root group."""
    h = _Hdf5(path)
    out = {}
    for name, addr in h.links(h.root).items():
        try:
            arr = h.dataset(addr)
        except ValueError:
            arr = None
        if arr is not None:
            out[name] = arr
    return out


if __name__ == "__main__":
    import sys
    for k, v in read_nc(sys.argv[1]).items():
        print(k, v.dtype, v.shape, v.ravel()[:5])
