"""Continuous features from bigWig tracks: `get_mean_bw_for_bed` (MuRaL/data/preprocessing.py:725-750) without
pyBigWig — the C library's reader (csrc/bigwig.cu) decodes a track once per chromosome into prefix sums, a site's window mean
is two lookups."""
import ctypes as C

import numpy as np

from . import _lib


class BigWig:
    def __init__(self, path):
        h = C.c_void_p()
        _lib.check(_lib.lib().mural_bigwig_open(str(path).encode(), C.byref(h)))
        self._h = h
        L = _lib.lib()
        self.chroms = {L.mural_bigwig_chrom_name(h, i).decode(): int(L.mural_bigwig_chrom_len(h, i)) for i in range(L.mural_bigwig_n_chrom(h))}

    def window_means(self, chrom, lo, hi):
        """mean(nan_to_num(values(chrom, max(lo, 0), min(hi, len)))) per window (end exclusive)."""
        lo, hi = np.ascontiguousarray(lo, dtype=np.int64), np.ascontiguousarray(hi, dtype=np.int64)
        out = np.empty(len(lo), dtype=np.float64)
        _lib.check(_lib.lib().mural_bigwig_window_means(self._h, chrom.encode(), len(lo), _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().mural_bigwig_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mean_bw_for_sites(bw_files, bw_radii, sites, model_type="snv"):
    """float64 [n_sites, n_tracks] in FILE order == get_mean_bw_for_bed(...).values.  Like the reference, the window of track j
    is expanded from the window of track j-1 (its loop re-assigns `start, stop = get_expanded_region(start, stop, ...)`,
    preprocessing.py:741), i.e. the effective radius of track j is bw_radii[0] + ... + bw_radii[j]."""
    n = len(sites)
    out = np.zeros((n, len(bw_files)), dtype=np.float64)
    lo, hi = sites.start.astype(np.int64).copy(), sites.end.astype(np.int64).copy()
    for j, (path, R) in enumerate(zip(bw_files, bw_radii)):
        lo = lo - int(R) + (1 if model_type == "indel" else 0)             # extend_interval (preprocessing.py:559-567)
        hi = hi + int(R)
        bw = BigWig(path)
        try:
            for ci, name in enumerate(sites.chrom_names):
                m = sites.chrom == ci
                if m.any():
                    out[m, j] = bw.window_means(name, lo[m], hi[m])
        finally:
            bw.close()
    return out
