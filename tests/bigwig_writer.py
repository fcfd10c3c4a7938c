"""TEST INFRASTRUCTURE: a minimal bigWig writer (Kent et al. 2010 layout) so that the library's reader (csrc/bigwig.cu) can be
exercised without pyBigWig: dense per-base tracks with NaN gaps -> fixedStep / variableStep / bedGraph sections (cycled),
optionally zlib-deflated, a one-leaf chromosome B+ tree and a two-level R-tree."""
import struct
import zlib

import numpy as np


def write_bigwig(path, tracks, compress=True, items_per_section=512, leaf_fanout=8):
    """tracks: {chrom name: float32 array with NaN where there is no data} (insertion order = chromosome ids)."""
    names = list(tracks)
    key_size = max(len(n) for n in names)
    sections = []                                            # (chrom id, start, end, payload bytes)
    kind = 0
    for cid, name in enumerate(names):
        v = np.asarray(tracks[name], dtype=np.float32)
        ok = ~np.isnan(v)
        edges = np.flatnonzero(np.diff(np.r_[0, ok.astype(np.int8), 0]))
        for a, b in zip(edges[::2], edges[1::2]):            # maximal runs of covered bases
            for s in range(a, b, items_per_section):
                e = min(b, s + items_per_section)
                vals = v[s:e]
                t = (1, 2, 3)[kind % 3]
                kind += 1
                if t == 3:                                   # fixedStep, step = span = 1
                    body = vals.astype("<f4").tobytes()
                    hdr = struct.pack("<IIIIIBBH", cid, s, e, 1, 1, 3, 0, len(vals))
                elif t == 2:                                 # variableStep, span 1
                    body = b"".join(struct.pack("<If", s + i, float(x)) for i, x in enumerate(vals))
                    hdr = struct.pack("<IIIIIBBH", cid, s, e, 0, 1, 2, 0, len(vals))
                else:                                        # bedGraph: runs of equal values
                    cuts = np.r_[0, np.flatnonzero(np.diff(vals)) + 1, len(vals)]
                    body = b"".join(struct.pack("<IIf", s + c0, s + c1, float(vals[c0])) for c0, c1 in zip(cuts[:-1], cuts[1:]))
                    hdr = struct.pack("<IIIIIBBH", cid, s, e, 0, 0, 1, 0, len(cuts) - 1)
                sections.append((cid, s, e, hdr + body))
    raw_max = max(len(p) for *_, p in sections)
    out = bytearray(64)                                      # header, patched at the end
    chrom_tree = len(out)
    out += struct.pack("<IIIIQQ", 0x78CA8C91, len(names), key_size, 8, len(names), 0)
    out += struct.pack("<BBH", 1, 0, len(names))
    for cid, name in enumerate(names):
        out += name.encode().ljust(key_size, b"\0") + struct.pack("<II", cid, len(tracks[name]))
    full_data = len(out)
    out += struct.pack("<Q", len(sections))
    placed = []
    for cid, s, e, payload in sections:
        blob = zlib.compress(payload) if compress else payload
        placed.append((cid, s, e, len(out), len(blob)))
        out += blob
    full_index = len(out)
    out += struct.pack("<IIQIIIIQII", 0x2468ACE0, leaf_fanout, len(placed), placed[0][0], placed[0][1], placed[-1][0], placed[-1][2],
                       full_index, 1, 0)
    groups = [placed[i:i + leaf_fanout] for i in range(0, len(placed), leaf_fanout)]
    root_off = len(out)
    root_size = 4 + 24 * len(groups)
    leaf_offs, off = [], root_off + root_size
    for g in groups:
        leaf_offs.append(off)
        off += 4 + 32 * len(g)
    out += struct.pack("<BBH", 0, 0, len(groups))
    for g, lo in zip(groups, leaf_offs):
        out += struct.pack("<IIIIQ", g[0][0], g[0][1], g[-1][0], g[-1][2], lo)
    for g in groups:
        out += struct.pack("<BBH", 1, 0, len(g))
        for cid, s, e, o, n in g:
            out += struct.pack("<IIIIQQ", cid, s, cid, e, o, n)
    out[0:64] = struct.pack("<IHHQQQHHQQIQ", 0x888FFC26, 4, 0, chrom_tree, full_data, full_index, 0, 0, 0, 0, raw_max if compress else 0, 0)
    with open(path, "wb") as f:
        f.write(bytes(out))
