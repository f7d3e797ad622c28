"""Host-side BigWig plumbing (derfinder_b200/bigwig.py) against the fixture writer: chromosome tree, R-tree
look-up, zlib, the three section types, the Rle a whole-chromosome read returns.  No GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bigwig_writer import write_bigwig  # noqa: E402

from derfinder_b200.bigwig import BigWigFile  # noqa: E402


@pytest.mark.parametrize("compress", [True, False])
def test_bigwig_reader_round_trip(tmp_path, compress):
    rng = np.random.default_rng(3)
    items, pos = [], 7
    for _ in range(700):
        l, g = int(rng.integers(1, 60)), int(rng.integers(0, 30))
        pos += g
        items.append((pos, pos + l, float(np.float32(rng.gamma(2.0, 3.0)))))
        pos += l
    size = pos + 123
    path = str(tmp_path / "t.bw")
    write_bigwig(path, {"chr21": (size, "bedGraph", items), "chrA": (1000, "fixedStep", (10, 5, 5, [1.0, 2.0, 2.0, 3.0])),
                        "chrB": (2000, "varStep", (3, [(5, 1.5), (8, 1.5), (100, 4.0)]))}, items_per_section=37,
                 compress=compress)
    bw = BigWigFile(path)
    assert bw.chroms == {"chr21": (0, size), "chrA": (1, 1000), "chrB": (2, 2000)}
    v, l = bw.read_rle("chr21")
    ref = np.zeros(size)
    for s, e, val in items:
        ref[s:e] = val
    assert np.array_equal(np.repeat(v, l), ref) and not np.any(v[1:] == v[:-1])
    v, l = bw.read_rle("chrA")
    assert list(v) == [0.0, 1.0, 2.0, 3.0, 0.0] and list(l) == [10, 5, 10, 5, 970]
    v, l = bw.read_rle("chrB")
    assert list(v) == [0.0, 1.5, 0.0, 4.0, 0.0] and list(l) == [5, 6, 89, 3, 1897]
    # only the leaf blocks that overlap the requested ranges are read
    blob_all, offs_all, cid = bw.sections("chr21")
    blob, offs, cid2 = bw.sections("chr21", [items[100][0] + 1], [items[100][1]])
    assert cid == cid2 == 0 and 1 <= len(offs) - 1 <= 2 < len(offs_all) - 1
    blob, offs, _ = bw.sections("chr21", [size + 10], [size + 20])
    assert len(offs) == 1 and len(blob) == 0
    bw.close()
