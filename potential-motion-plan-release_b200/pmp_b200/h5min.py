"""Minimal pure-Python reader for the HDF5 files the reference writes with h5py's defaults
(`h5py.File(fname, 'w')` + `create_dataset(k, data=..., compression='gzip')`: pb_diff_envs/utils/val_utils_mp.py:228-230,
scripts/gen_m2d_mp.py:208-210), read back by `OfflineEnv.get_dataset` (pb_diff_envs/environment/offline_env.py:83-103).

h5py is not in this image, and the planners need only numeric arrays out of `<dataset>-problems.hdf5` (`problems`,
`infos/wall_locations`, `maze_idx` ...), so this module implements exactly the subset of the HDF5 file format
(HDF5 File Format Specification version 2.0/3.0) such files use:
  superblock version 0 / 1, old-style groups (symbol table message, version-1 B-tree of type 0, local heap, SNOD nodes),
  version-1 object headers with continuation blocks, dataspace versions 1 / 2, little-endian fixed-point and IEEE float
  datatypes, data layout version 3 (compact, contiguous, chunked through a version-1 B-tree of type 1), filter pipeline
  versions 1 / 2 with deflate (gzip), shuffle and fletcher32.
Anything else (new-style groups, virtual datasets, variable-length types ...) raises H5FormatError naming the feature, so a
file this reader cannot handle fails loudly instead of returning wrong numbers.  pmp_b200.ingest uses h5py when it is
installed and this reader otherwise."""
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _File:
    def __init__(self, buf):
        self.b = memoryview(buf)
        if bytes(self.b[:8]) != SIG:
            raise H5FormatError("not an HDF5 file (signature missing at offset 0)")
        ver = self.b[8]
        if ver not in (0, 1):
            raise H5FormatError("superblock version %d is not supported (only the h5py default, versions 0 / 1)" % ver)
        so, sl = self.b[13], self.b[14]
        if so != 8 or sl != 8:
            raise H5FormatError("only 8-byte offsets / lengths are supported (got %d / %d)" % (so, sl))
        p = 24 if ver == 0 else 28          # v1 adds indexed-storage K + 2 reserved bytes
        self.base, _free, self.eof, _drv = struct.unpack_from("<4Q", self.b, p)
        p += 32
        # root group symbol table entry
        _name, self.root_oh, cache, _r = struct.unpack_from("<QQII", self.b, p)

    # ------------------------------------------------------------------ primitives
    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    def messages(self, addr):
        """(type, flags, bytes) of every message of a version-1 object header, following continuation blocks"""
        addr += self.base
        ver = self.b[addr]
        if ver != 1:
            if bytes(self.b[addr:addr + 4]) == b"OHDR":
                raise H5FormatError("version-2 object headers (libver='latest' files) are not supported")
            raise H5FormatError("object header version %d at %d" % (ver, addr))
        nmsg, _refc, hsize = self.u("HII", addr + 2)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u("HHB", p)
                data = self.b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff + self.base, clen))
                out.append((mtype, flags, data))
        return out

    def heap_name(self, heap_addr, off):
        heap_addr += self.base
        if bytes(self.b[heap_addr:heap_addr + 4]) != b"HEAP":
            raise H5FormatError("local heap signature missing at %d" % heap_addr)
        _size, _free, seg = self.u("QQQ", heap_addr + 8)
        p = seg + self.base + off
        e = p
        while self.b[e] != 0:
            e += 1
        return bytes(self.b[p:e]).decode("utf-8")

    def group_entries(self, btree, heap):
        """name -> object header address of an old-style group"""
        out = {}
        stack = [btree]
        while stack:
            a = stack.pop() + self.base
            sig = bytes(self.b[a:a + 4])
            if sig == b"TREE":
                ntype, level, used = self.u("BBH", a + 4)
                if ntype != 0:
                    raise H5FormatError("group B-tree node of type %d" % ntype)
                p = a + 24
                for i in range(used):
                    child = self.u("Q", p + 8 + i * 16)[0]      # key_i (8) child_i (8) ...
                    stack.append(child)
            elif sig == b"SNOD":
                nsym = self.u("H", a + 6)[0]
                p = a + 8
                for i in range(nsym):
                    noff, oh, _cache, _r = self.u("QQII", p + i * 40)
                    out[self.heap_name(heap, noff)] = oh
            else:
                raise H5FormatError("unexpected node signature %r in a group B-tree" % sig)
        return out

    # ------------------------------------------------------------------ objects
    def read_object(self, oh_addr):
        msgs = self.messages(oh_addr)
        types = {t for t, _, _ in msgs}
        if 0x0011 in types:      # symbol table message: a group
            for t, _, d in msgs:
                if t == 0x0011:
                    bt, hp = struct.unpack_from("<QQ", d, 0)
                    return {k: self.read_object(a) for k, a in self.group_entries(bt, hp).items()}
        if 0x0006 in types or 0x0002 in types or 0x000A in types:
            raise H5FormatError("new-style groups (link / link-info messages) are not supported")
        if 0x0008 not in types:
            raise H5FormatError("object at %d is neither an old-style group nor a dataset" % oh_addr)
        return self.read_dataset(msgs)

    @staticmethod
    def parse_dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                raise H5FormatError("null dataspace")
            p = 4
        else:
            raise H5FormatError("dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, d, p)) if rank else ()

    @staticmethod
    def parse_datatype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        bits0 = d[1]
        size = struct.unpack_from("<I", d, 4)[0]
        if bits0 & 1:
            raise H5FormatError("big-endian datatypes are not supported")
        if cls == 0:
            signed = bool(bits0 & 0x08)
            if size not in (1, 2, 4, 8):
                raise H5FormatError("%d-byte integers" % size)
            return np.dtype("<%s%d" % ("i" if signed else "u", size))
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5FormatError("%d-byte floats" % size)
            return np.dtype("<f%d" % size)
        names = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum",
                 9: "variable-length", 10: "array"}
        raise H5FormatError("datatype class %s (version %d) is not supported: numeric arrays only" % (names.get(cls, cls), ver))

    @staticmethod
    def parse_filters(d):
        ver, n = d[0], d[1]
        out = []
        p = 8 if ver == 1 else 2
        for _ in range(n):
            fid = struct.unpack_from("<H", d, p)[0]
            if ver == 1 or fid >= 256:
                nlen, flags, ncd = struct.unpack_from("<HHH", d, p + 2)
                p += 8
            else:
                nlen = 0
                flags, ncd = struct.unpack_from("<HH", d, p + 2)
                p += 6
            if ver == 1:
                nlen = (nlen + 7) // 8 * 8
            p += nlen
            cd = struct.unpack_from("<%dI" % ncd, d, p) if ncd else ()
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def read_dataset(self, msgs):
        shape = dtype = layout = None
        filters = []
        for t, _, d in msgs:
            if t == 0x0001:
                shape = self.parse_dataspace(d)
            elif t == 0x0003:
                dtype = self.parse_datatype(d)
            elif t == 0x0008:
                layout = d
            elif t == 0x000B:
                filters = self.parse_filters(d)
        if shape is None or dtype is None or layout is None:
            raise H5FormatError("dataset without dataspace / datatype / layout message")
        n = int(np.prod(shape)) if shape else 1
        if layout[0] != 3:
            raise H5FormatError("data layout message version %d (only 3)" % layout[0])
        cls = layout[1]
        if cls == 0:         # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            raw = bytes(layout[4:4 + size])
            return np.frombuffer(raw, dtype, n).reshape(shape).copy()
        if cls == 1:         # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return np.zeros(shape, dtype)
            a = addr + self.base
            return np.frombuffer(self.b[a:a + n * dtype.itemsize], dtype, n).reshape(shape).copy()
        if cls != 2:
            raise H5FormatError("data layout class %d" % cls)
        rank1 = layout[2]
        bt = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % rank1, layout, 11)
        rank = rank1 - 1
        if rank != len(shape) or cdims[-1] != dtype.itemsize:
            raise H5FormatError("chunk layout does not match the dataspace")
        chunk = tuple(cdims[:-1])
        out = np.zeros(shape, dtype)
        if bt == UNDEF:
            return out
        csize = int(np.prod(chunk)) * dtype.itemsize
        stack = [bt]
        while stack:
            a = stack.pop() + self.base
            if bytes(self.b[a:a + 4]) != b"TREE":
                raise H5FormatError("chunk B-tree signature missing at %d" % a)
            ntype, level, used = self.u("BBH", a + 4)
            if ntype != 1:
                raise H5FormatError("chunk B-tree node of type %d" % ntype)
            ksz = 8 + 8 * rank1
            p = a + 24
            for i in range(used):
                nbytes, mask = self.u("II", p)
                offs = self.u("%dQ" % rank1, p + 8)
                child = self.u("Q", p + ksz)[0]
                p += ksz + 8
                if level > 0:
                    stack.append(child)
                    continue
                raw = bytes(self.b[child + self.base:child + self.base + nbytes])
                for j in reversed(range(len(filters))):        # undo the pipeline, last filter first
                    if mask & (1 << j):
                        continue
                    fid, cd = filters[j]
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        m = len(raw) // es
                        raw = np.frombuffer(raw, np.uint8)[:m * es].reshape(es, m).T.tobytes() + raw[m * es:]
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise H5FormatError("filter id %d is not supported (deflate, shuffle, fletcher32 only)" % fid)
                if len(raw) < csize:
                    raise H5FormatError("chunk holds %d bytes, expected %d" % (len(raw), csize))
                blk = np.frombuffer(raw, dtype, csize // dtype.itemsize).reshape(chunk)
                sl_out, sl_in = [], []
                for o, c, s in zip(offs[:-1], chunk, shape):
                    e = min(o + c, s)
                    sl_out.append(slice(o, e))
                    sl_in.append(slice(0, e - o))
                if all(s.stop > s.start for s in sl_out):
                    out[tuple(sl_out)] = blk[tuple(sl_in)]
        return out


def _flatten(tree, prefix, out):
    for k, v in tree.items():
        if isinstance(v, dict):
            _flatten(v, prefix + k + "/", out)
        else:
            out[prefix + k] = v


def read_flat(path):
    """every dataset of the file as {'group/sub/name': numpy array} -- the flat dict `OfflineEnv.get_dataset` builds with
    its `get_keys` walk (offline_env.py:14-25)"""
    with open(path, "rb") as f:
        buf = f.read()
    hf = _File(buf)
    tree = hf.read_object(hf.root_oh)
    if not isinstance(tree, dict):
        raise H5FormatError("the root object is not a group")
    out = {}
    _flatten(tree, "", out)
    return out
