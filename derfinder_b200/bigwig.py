"""Host side of BigWig access for railMatrix() (R/railMatrix.R:181, 271-289): the part rtracklayer plays in R.

Only file plumbing lives here -- header, chromosome B+ tree, R-tree look-up of the leaf blocks that overlap a set
of ranges, zlib inflation.  The inflated data sections are handed to the library untouched
(`dfb_cov_from_bigwig`) and decoded on the GPU.  `read_rle` decodes a whole chromosome on the host; it is used
for the single mean track of `summaryFiles` (whose Rle goes through filterData like any coverage) and by tests.

Format: UCSC bigWig (Kent et al. 2010), little endian: 64-byte header, chromosome B+ tree (magic 0x78CA8C91),
data sections (24-byte header + bedGraph / variableStep / fixedStep items), R-tree index (magic 0x2468ACE0).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

BIGWIG_MAGIC = 0x888FFC26
CHROM_TREE_MAGIC = 0x78CA8C91
RTREE_MAGIC = 0x2468ACE0


class BigWigFile:
    def __init__(self, path):
        self.path = path
        self.f = open(path, "rb")
        hdr = self.f.read(64)
        (magic, self.version, self.zoom_levels, chrom_tree, self.full_data, self.full_index, _fc, _dfc, _asql, _ts,
         self.uncompress_buf, _res) = struct.unpack("<IHHQQQHHQQIQ", hdr)
        if magic != BIGWIG_MAGIC:
            raise ValueError(f"{path}: not a little-endian bigWig file")
        self.chroms = {}
        self._read_chrom_tree(chrom_tree)
        self._leaves = None

    def close(self):
        self.f.close()

    # ---- chromosome B+ tree: name -> (id, size)
    def _read_chrom_tree(self, off):
        self.f.seek(off)
        magic, _bs, key_size, val_size, _count, _res = struct.unpack("<IIIIQQ", self.f.read(32))
        if magic != CHROM_TREE_MAGIC or val_size != 8:
            raise ValueError(f"{self.path}: bad chromosome tree")

        def node(at):
            self.f.seek(at)
            is_leaf, _r, count = struct.unpack("<BBH", self.f.read(4))
            if is_leaf:
                for _ in range(count):
                    key = self.f.read(key_size).rstrip(b"\0").decode()
                    cid, size = struct.unpack("<II", self.f.read(8))
                    self.chroms[key] = (cid, size)
            else:
                kids = []
                for _ in range(count):
                    self.f.read(key_size)
                    kids.append(struct.unpack("<Q", self.f.read(8))[0])
                for k in kids:
                    node(k)

        node(off + 32)

    # ---- R-tree: leaf blocks (start chrom, start base, end chrom, end base, data offset, data size)
    def leaves(self):
        if self._leaves is None:
            self.f.seek(self.full_index)
            magic, = struct.unpack("<I", self.f.read(4))
            if magic != RTREE_MAGIC:
                raise ValueError(f"{self.path}: bad R-tree index")
            self.f.read(44)
            out = []

            def node(at):
                self.f.seek(at)
                is_leaf, _r, count = struct.unpack("<BBH", self.f.read(4))
                if is_leaf:
                    for _ in range(count):
                        out.append(struct.unpack("<IIIIQQ", self.f.read(32)))
                else:
                    kids = [struct.unpack("<IIIIQ", self.f.read(24))[4] for _ in range(count)]
                    for k in kids:
                        node(k)

            node(self.full_index + 48)
            self._leaves = out
        return self._leaves

    def sections(self, chrom, starts=None, ends=None):
        """Inflated data sections of `chrom` that overlap any of the 1-based inclusive ranges (all of the
        chromosome's when no ranges are given): (blob bytes, int64 offsets [nsec + 1], chromosome id)."""
        cid, _size = self.chroms[chrom]
        st = None if starts is None else np.asarray(starts, dtype=np.int64)
        en = None if ends is None else np.asarray(ends, dtype=np.int64)
        parts, offs = [], [0]
        for sc, sb, ec, eb, off, size in self.leaves():
            if sc > cid or ec < cid:
                continue
            lo = sb if sc == cid else 0            # 0-based half open [lo, hi) on this chromosome
            hi = eb if ec == cid else 2 ** 32
            if st is not None:
                k = int(np.searchsorted(en, lo + 1, side="left"))  # first range ending at or after base lo + 1
                if k >= len(st) or st[k] > hi:
                    continue
            self.f.seek(off)
            raw = self.f.read(size)
            data = zlib.decompress(raw) if self.uncompress_buf else raw
            parts.append(data)
            offs.append(offs[-1] + len(data))
        return b"".join(parts), np.asarray(offs, dtype=np.int64), cid

    def read_rle(self, chrom):
        """Whole chromosome as an Rle (values float64, lengths int64) of the chromosome's length, 0 where the
        file has no data -- what `import(file, as = "RleList")[[chrom]]` returns."""
        blob, offs, cid = self.sections(chrom)
        size = self.chroms[chrom][1]
        s_all, e_all, v_all = [], [], []
        for i in range(len(offs) - 1):
            sec = blob[offs[i]:offs[i + 1]]
            scid, cstart, _cend, step, span, typ, _r, cnt = struct.unpack("<IIIIIBBH", sec[:24])
            if scid != cid:
                continue
            body = sec[24:]
            if typ == 1:
                a = np.frombuffer(body, dtype=np.dtype([("s", "<u4"), ("e", "<u4"), ("v", "<f4")]), count=cnt)
                s, e, v = a["s"].astype(np.int64), a["e"].astype(np.int64), a["v"].astype(np.float64)
            elif typ == 2:
                a = np.frombuffer(body, dtype=np.dtype([("s", "<u4"), ("v", "<f4")]), count=cnt)
                s, v = a["s"].astype(np.int64), a["v"].astype(np.float64)
                e = s + span
            elif typ == 3:
                v = np.frombuffer(body, dtype="<f4", count=cnt).astype(np.float64)
                s = cstart + np.arange(cnt, dtype=np.int64) * step
                e = s + span
            else:
                raise ValueError(f"{self.path}: unknown section type {typ}")
            s_all.append(s), e_all.append(e), v_all.append(v)
        if not s_all:
            return np.zeros(1), np.asarray([size], dtype=np.int64)
        s, e, v = np.concatenate(s_all), np.concatenate(e_all), np.concatenate(v_all)
        o = np.argsort(s, kind="stable")
        s, e, v = s[o], e[o], v[o]
        # interleave gaps (value 0) and items, then merge equal neighbours like Rle() does
        prev_end = np.concatenate(([0], e[:-1]))
        gap = s - prev_end
        vals = np.empty(2 * len(s) + 1)
        lens = np.empty(2 * len(s) + 1, dtype=np.int64)
        vals[0:-1:2], lens[0:-1:2] = 0.0, gap
        vals[1:-1:2], lens[1:-1:2] = v, e - s
        vals[-1], lens[-1] = 0.0, size - e[-1]
        keep = lens > 0
        vals, lens = vals[keep], lens[keep]
        new = np.concatenate(([True], vals[1:] != vals[:-1]))
        grp = np.cumsum(new) - 1
        out_l = np.zeros(int(grp[-1]) + 1, dtype=np.int64)
        np.add.at(out_l, grp, lens)
        return vals[new], out_l
