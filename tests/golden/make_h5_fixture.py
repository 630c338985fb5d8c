"""Writes tests/golden/m2d_synth-problems.hdf5: the contents of m2d_synth-problems.npz as an HDF5 file laid out the way
h5py's defaults lay out the reference's problem files (`h5py.File(fname, 'w')` + `create_dataset(k, data=...,
compression='gzip')`, pb_diff_envs/utils/val_utils_mp.py:228-230): superblock version 0, old-style groups (symbol
table + version-1 B-tree + local heap), version-1 object headers, chunked datasets indexed by a version-1 B-tree with the
deflate filter; 'infos/wall_locations' sits in the nested group the '/' in its name creates.  One dataset is written
contiguous and one compact so that every layout class of the reader (pmp_b200/h5min.py) is exercised.

h5py is not installed in this image, so the file is assembled by hand from the HDF5 File Format Specification; the reader
is independent code written from the same specification.  Both sides being wrong in the same way is the residual risk:
a real problem file should be diffed once where h5py exists (`ingest.load_problems` prefers h5py when it is importable).

    python tests/golden/make_h5_fixture.py
"""
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UNDEF = 0xFFFFFFFFFFFFFFFF


class Writer:
    def __init__(self):
        self.buf = bytearray(96)          # superblock v0 goes here last

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        a = len(self.buf)
        self.buf += data
        return a

    # ---- messages
    @staticmethod
    def msg(mtype, data, flags=0):
        data = bytes(data)
        data += b"\0" * (-len(data) % 8)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def object_header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    @staticmethod
    def dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 1) + struct.pack("<%dQ" % len(shape), *shape) * 2    # dims + max dims

    @staticmethod
    def datatype(dt):
        dt = np.dtype(dt)
        if dt.kind == "f":
            exp, man = {4: (8, 23), 8: (11, 52)}[dt.itemsize]
            bits = dt.itemsize * 8
            return (struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize) +
                    struct.pack("<HHBBBBI", 0, bits, man, exp, 0, man, (1 << (exp - 1)) - 1))
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)

    # ---- datasets
    def dataset(self, arr, layout="chunked", chunk=None, level=4):
        arr = np.ascontiguousarray(arr)
        msgs = [self.msg(0x0001, self.dataspace(arr.shape)), self.msg(0x0003, self.datatype(arr.dtype), flags=1)]
        if layout == "compact":
            raw = arr.tobytes()
            msgs.append(self.msg(0x0008, struct.pack("<BBH", 3, 0, len(raw)) + raw))
        elif layout == "contiguous":
            a = self.alloc(arr.tobytes())
            msgs.append(self.msg(0x0008, struct.pack("<BBQQ", 3, 1, a, arr.nbytes)))
        else:
            rank = arr.ndim
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
            keys = []
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = [g[i] for g, i in zip(grid, offs)]
                blk = np.zeros(chunk, arr.dtype)          # edge chunks are stored whole
                sl = tuple(slice(a_, min(a_ + c, s)) for a_, c, s in zip(o, chunk, arr.shape))
                blk[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
                z = zlib.compress(blk.tobytes(), level)
                keys.append((len(z), o, self.alloc(z, align=1)))
            K = 32
            ksz = 8 + 8 * (rank + 1)
            full = 24 + (2 * K + 1) * ksz + 2 * K * 8
            end_key = struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), *[len(g) * c for g, c in zip(grid, chunk)], 0)

            def key_bytes(n, o):
                return struct.pack("<II", n, 0) + struct.pack("<%dQ" % (rank + 1), *o, 0)

            def write_node(level, entries):      # entries: (first key bytes, child address)
                node = struct.pack("<4sBBHQQ", b"TREE", 1, level, len(entries), UNDEF, UNDEF)
                for kb, a in entries:
                    node += kb + struct.pack("<Q", a)
                node += end_key
                return self.alloc(node + b"\0" * (full - len(node)))

            # leaves hold at most 2K chunks; more chunks get one more tree level (like the library does)
            level, entries = 0, [(key_bytes(n, o), a) for n, o, a in keys]
            while True:
                groups = [entries[i:i + 2 * K] for i in range(0, len(entries), 2 * K)]
                nodes = [(g[0][0], write_node(level, g)) for g in groups]
                if len(nodes) == 1:
                    bt = nodes[0][1]
                    break
                level, entries = level + 1, nodes
            msgs.append(self.msg(0x000B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" +
                                 struct.pack("<I4x", level)))
            msgs.append(self.msg(0x0008, struct.pack("<BBBQ", 3, 2, rank + 1, bt) + struct.pack("<%dI" % (rank + 1), *chunk, arr.dtype.itemsize)))
        msgs.append(self.msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)))      # fill value v2: allocate late, write if set, undefined
        return self.object_header(msgs)

    # ---- groups
    def group(self, entries):
        """entries: name -> object header address; returns (object header, B-tree, heap) addresses"""
        names = sorted(entries)
        seg = bytearray(8)                       # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(seg)
            b = n.encode() + b"\0"
            seg += b + b"\0" * (-len(b) % 8)
        seg += b"\0" * max(0, 88 - len(seg))
        seg_a = self.alloc(bytes(seg))
        heap = self.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(seg), UNDEF, seg_a))
        # symbol nodes hold at most 2K = 8 entries (group leaf K = 4): larger groups get several under one B-tree node
        chunks = [names[i:i + 8] for i in range(0, len(names), 8)]
        K = 16
        assert len(chunks) <= 2 * K
        node = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(chunks), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for ch in chunks:
            snod = struct.pack("<4sBxH", b"SNOD", 1, len(ch))
            for n in ch:
                snod += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
            snod += b"\0" * (8 + 8 * 40 - len(snod))
            node += struct.pack("<QQ", self.alloc(snod), offs[ch[-1]])      # child, then the key = its largest name
        bt = self.alloc(node + b"\0" * (24 + (2 * K + 1) * 8 + 2 * K * 8 - len(node)))
        oh = self.object_header([self.msg(0x0011, struct.pack("<QQ", bt, heap))])
        return oh, bt, heap

    def finish(self, root):
        oh, bt, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def main():
    z = np.load(os.path.join(HERE, "m2d_synth-problems.npz"))
    w = Writer()
    top, infos = {}, {}
    for k in z.files:
        a = z[k]
        if k == "problems":
            h = w.dataset(a, chunk=(max(1, a.shape[0] // 2 + 1), max(1, a.shape[1] // 3 + 1)) + a.shape[2:])      # ragged edge chunks
        elif k == "infos/wall_locations":
            h = w.dataset(a, chunk=(1, 1) + a.shape[2:])          # 120 chunks: a two-level chunk B-tree
        elif a.size <= 64:
            h = w.dataset(a, layout="compact")
        else:
            h = w.dataset(a, layout="contiguous")
        if k.startswith("infos/"):
            infos[k[len("infos/"):]] = h
        else:
            top[k] = h
    # extra members of a real problems file the planners ignore: exercised by the reader all the same
    top["maze_idx"] = w.dataset(np.repeat(np.arange(z["problems"].shape[0], dtype=np.int64)[:, None], z["problems"].shape[1], 1),
                                chunk=(4, 4))
    infos["is_sol_kp"] = w.dataset(np.ones(z["problems"].shape[:2], np.uint8), layout="contiguous")
    for i in range(9):             # a real file has more members than one symbol node holds (8)
        top["extra_%02d" % i] = w.dataset(np.arange(3 + i, dtype=np.float64) * 0.5, layout="compact")
    top["infos"] = w.group(infos)[0]
    data = w.finish(w.group(top))
    out = os.path.join(HERE, "m2d_synth-problems.hdf5")
    open(out, "wb").write(data)
    print("wrote", out, len(data), "bytes; datasets:", sorted(top), sorted(infos))


if __name__ == "__main__":
    main()
