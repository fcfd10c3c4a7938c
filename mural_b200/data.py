"""Host side of the data path: BED sites -> (pos, meta) site records in the reference's sample order.

Mirrors MuRaL/data/preprocessing.py: bed_reader (:39-106, segment/strand batching that defines the
sample order), prepare_dataset_np (:828-848) and generate_data_batches (:1148-1177), but the dataset
yields *site records* (8 bytes/site) instead of one-hot tensors (32 KB/site); window extraction
happens on the GPU from the packed genome.
"""
import gzip

import numpy as np


def pack_meta(strand, label, chrom):
    """meta = strand | label<<1 | chrom<<8  (MURAL_META in include/mural_b200.h)."""
    return ((np.asarray(chrom, dtype=np.int64) << 8) | ((np.asarray(label, dtype=np.int64) & 0x7f) << 1)
            | (np.asarray(strand, dtype=np.int64) & 1)).astype(np.int32)


class SiteTable:
    """Column store of BED records in FILE order."""

    def __init__(self, chrom_names, chrom, start, end, strand, label):
        self.chrom_names = list(chrom_names)           # index -> name, order of first appearance
        self.chrom = np.asarray(chrom, dtype=np.int32)
        self.start = np.asarray(start, dtype=np.int64)
        self.end = np.asarray(end, dtype=np.int64)
        self.strand = np.asarray(strand, dtype=np.int8)   # 0 '+', 1 anything else (bed_reader :97-101)
        self.label = np.asarray(label, dtype=np.int64)

    def __len__(self):
        return len(self.start)

    @classmethod
    def from_bed(cls, path):
        """chrom start end name score strand (BED6; score = label, preprocessing.py:752-754), plain or .gz — parsed by the
        C library's streaming reader (mural_bed_read; replaces iterating BedTool(file), preprocessing.py:39-106)."""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.mural_bed_read(str(path).encode(), C.byref(h)))
        try:
            n = int(L.mural_bed_n(h))
            names = [L.mural_bed_chrom_name(h, i).decode() for i in range(L.mural_bed_n_chrom(h))]
            chrom, strand = np.empty(n, np.int32), np.empty(n, np.int8)
            start, end, label = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64)
            _lib.check(L.mural_bed_columns(h, _lib.ptr(chrom), _lib.ptr(start), _lib.ptr(end), _lib.ptr(strand), _lib.ptr(label)))
        finally:
            L.mural_bed_destroy(h)
        return cls(names, chrom, start, end, strand, label)


def segment_order(chrom, start, strand, segment_center):
    """Emission order of bed_reader (preprocessing.py:39-106): sites are grouped into windows of
    `segment_center` bp (anchored at the first site of the first chromosome, at 1 for later
    chromosomes), '+' batch before '-' batch inside each window.  Returns (perm, batch_sizes)."""
    chrom = np.asarray(chrom); start = np.asarray(start, dtype=np.int64); strand = np.asarray(strand, dtype=np.int64)
    n = len(start)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    brk = np.empty(n, dtype=bool); brk[0] = True
    brk[1:] = chrom[1:] != chrom[:-1]                   # a chromosome re-appearing later opens a new block
    block = np.cumsum(brk) - 1
    anchor = np.ones(block[-1] + 1, dtype=np.int64)
    anchor[0] = start[0]
    rel = start - anchor[block]
    win = np.maximum(0, (rel + segment_center - 1) // segment_center - 1)   # first window whose end >= start
    key = np.maximum.accumulate(block * (int(win.max()) + 2) + win)          # window end never moves backwards
    key = key * 2 + (strand != 0)
    perm = np.argsort(key, kind="stable")
    ks = key[perm]
    edges = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1], True])
    return perm, np.diff(edges)


class SiteBatch:
    """A batch of site records on the device (what the fast Network2.forward consumes)."""
    __slots__ = ("pos", "meta", "genome")

    def __init__(self, pos, meta, genome):
        self.pos, self.meta, self.genome = pos, meta, genome

    def __len__(self):
        return int(self.pos.numel())


class PackedSiteDataset:
    """Dataset over (segment, strand) batches like CombinedDatasetNP (preprocessing.py:850-954), holding
    site records only.  `pos`/`meta` are in emission order; `batch_sizes` are the segment batch sizes."""

    def __init__(self, sites, genome, segment_center, local_radius, local_order, distal_radius, model_type="snv"):
        self.model_type = model_type
        self.genome = genome
        self.sites = sites
        self.local_radius, self.local_order, self.distal_radius = local_radius, local_order, distal_radius
        self.central_radius = segment_center
        for nme in sites.chrom_names:
            if nme not in genome.chrom_index:
                raise KeyError(nme)                                   # seq_records[chrom] (preprocessing.py:458)
        gidx = np.array([genome.chrom_index[nme] for nme in sites.chrom_names], dtype=np.int64)
        self.perm, self.batch_sizes = segment_order(sites.chrom, sites.start, sites.strand, segment_center)
        self.pos = sites.start[self.perm].astype(np.int32)
        self.strand = sites.strand[self.perm].astype(np.int8)
        self.label = sites.label[self.perm]
        self.chrom = gidx[sites.chrom[self.perm]]
        self.meta = pack_meta(self.strand, self.label, self.chrom)
        self.batch_offsets = np.r_[0, np.cumsum(self.batch_sizes)]
        self.n = len(self.batch_sizes)
        self.distal_info = True
        self.cont_cols = []
        self.cont_X = np.zeros((self.n, 1))

    def __len__(self):
        return self.n

    def get_distal_encoding_infomation(self):   # name kept from the reference (preprocessing.py:946)
        self.distal_info = True

    def __getitem__(self, index):
        """(y [m,1] float32, pos int32 [m], meta int32 [m]) of segment batch `index`."""
        a, b = self.batch_offsets[index], self.batch_offsets[index + 1]
        return self.label[a:b].astype(np.float32).reshape(-1, 1), self.pos[a:b], self.meta[a:b]

    def get_labels(self):
        return self.label.astype(np.float32)

    def position_info(self):
        """get_position_info (preprocessing.py:108-121): chrom, start, end, strand in emission order."""
        names = np.array(self.sites.chrom_names, dtype=object)[self.sites.chrom[self.perm]]
        return names, self.sites.start[self.perm], self.sites.end[self.perm], np.where(self.strand == 0, "+", "-")


def generate_site_batches(dataset, sampled_segments, batch_size, shuffle=False, seed=None, segment_indices=None,
                          device=None):
    """generate_data_batches (preprocessing.py:1148-1177) over site records: pools `sampled_segments`
    segment batches, shuffles inside the pool when asked, yields SiteBatch-es of `batch_size`; a short tail is
    PREPENDED to the next pool (Create_DatasetSegment.merge, :1219-1226) and the final tail is emitted."""
    import torch
    rng = np.random.default_rng(seed)
    segs = np.arange(len(dataset)) if segment_indices is None else np.asarray(segment_indices)
    if shuffle:
        segs = rng.permutation(segs)
    dev = device if device is not None else dataset.genome.device
    carry_pos = np.zeros(0, np.int32); carry_meta = np.zeros(0, np.int32)
    for s0 in range(0, len(segs), sampled_segments):
        pool = segs[s0:s0 + sampled_segments]
        pos = np.concatenate([carry_pos] + [dataset[i][1] for i in pool])
        meta = np.concatenate([carry_meta] + [dataset[i][2] for i in pool])
        if shuffle:
            o = rng.permutation(len(pos)); pos, meta = pos[o], meta[o]
        last = s0 + sampled_segments >= len(segs)
        n_full = len(pos) // batch_size * batch_size
        for b0 in range(0, n_full, batch_size):
            yield SiteBatch(torch.from_numpy(pos[b0:b0 + batch_size]).to(dev), torch.from_numpy(meta[b0:b0 + batch_size]).to(dev),
                            dataset.genome)
        carry_pos, carry_meta = pos[n_full:], meta[n_full:]
        if last and len(carry_pos):
            yield SiteBatch(torch.from_numpy(carry_pos).to(dev), torch.from_numpy(carry_meta).to(dev), dataset.genome)
