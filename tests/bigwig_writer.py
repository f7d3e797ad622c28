"""Minimal bigWig writer for fixtures (tests only): one R-tree leaf level, optional zlib, the three section
types.  Follows the UCSC bigWig layout closely enough for any reader that uses the R-tree index."""
import struct
import zlib


def write_bigwig(path, chroms, items_per_section=64, compress=True):
    """chroms: {name: (size, kind, payload)} with kind 'bedGraph' (payload [(start, end, value)], 0-based half
    open), 'varStep' (payload (span, [(start, value)])) or 'fixedStep' (payload (start, step, span, [value]))."""
    names = sorted(chroms)
    key_size = max(len(n) for n in names)
    sections, leaves = [], []
    for cid, name in enumerate(names):
        size, kind, payload = chroms[name]
        if kind == "bedGraph":
            for i in range(0, len(payload), items_per_section):
                part = payload[i:i + items_per_section]
                body = b"".join(struct.pack("<IIf", s, e, v) for s, e, v in part)
                hdr = struct.pack("<IIIIIBBH", cid, part[0][0], part[-1][1], 0, 0, 1, 0, len(part))
                sections.append((cid, part[0][0], part[-1][1], hdr + body))
        elif kind == "varStep":
            span, pts = payload
            for i in range(0, len(pts), items_per_section):
                part = pts[i:i + items_per_section]
                body = b"".join(struct.pack("<If", s, v) for s, v in part)
                hdr = struct.pack("<IIIIIBBH", cid, part[0][0], part[-1][0] + span, 0, span, 2, 0, len(part))
                sections.append((cid, part[0][0], part[-1][0] + span, hdr + body))
        else:
            start, step, span, vals = payload
            for i in range(0, len(vals), items_per_section):
                part = vals[i:i + items_per_section]
                s0 = start + i * step
                body = b"".join(struct.pack("<f", v) for v in part)
                hdr = struct.pack("<IIIIIBBH", cid, s0, s0 + (len(part) - 1) * step + span, step, span, 3, 0, len(part))
                sections.append((cid, s0, s0 + (len(part) - 1) * step + span, hdr + body))
    with open(path, "wb") as f:
        f.write(b"\0" * 64)
        chrom_tree = f.tell()
        f.write(struct.pack("<IIIIQQ", 0x78CA8C91, len(names), key_size, 8, len(names), 0))
        f.write(struct.pack("<BBH", 1, 0, len(names)))
        for cid, name in enumerate(names):
            f.write(name.encode().ljust(key_size, b"\0") + struct.pack("<II", cid, chroms[name][0]))
        full_data = f.tell()
        f.write(struct.pack("<Q", len(sections)))
        max_unc = 0
        for cid, s, e, raw in sections:
            max_unc = max(max_unc, len(raw))
            data = zlib.compress(raw) if compress else raw
            leaves.append((cid, s, cid, e, f.tell(), len(data)))
            f.write(data)
        full_index = f.tell()
        f.write(struct.pack("<IIQIIIIQII", 0x2468ACE0, 256, len(leaves), leaves[0][0], leaves[0][1], leaves[-1][2],
                            leaves[-1][3], full_index, 1, 0))
        # root -> leaf nodes of up to 256 items
        groups = [leaves[i:i + 256] for i in range(0, len(leaves), 256)]
        root_at = f.tell()
        f.write(struct.pack("<BBH", 0, 0, len(groups)))
        child_slots = f.tell()
        f.write(b"\0" * (24 * len(groups)))
        child_at = []
        for g in groups:
            child_at.append(f.tell())
            f.write(struct.pack("<BBH", 1, 0, len(g)))
            for leaf in g:
                f.write(struct.pack("<IIIIQQ", *leaf))
        f.seek(child_slots)
        for g, at in zip(groups, child_at):
            f.write(struct.pack("<IIIIQ", g[0][0], g[0][1], g[-1][2], g[-1][3], at))
        assert root_at == full_index + 48
        f.seek(0)
        f.write(struct.pack("<IHHQQQHHQQIQ", 0x888FFC26, 4, 0, chrom_tree, full_data, full_index, 0, 0, 0, 0,
                            max_unc if compress else 0, 0))
