"""Reference genome resident in HBM as a 2-bit packed array (+ non-ACGT run table).

Replaces `SeqIO.to_dict(SeqIO.parse(open(ref_genome), 'fasta'))` and the python-str genome the reference
keeps in RAM (MuRaL/data/preprocessing.py:836, 458, 964, 990).
"""
import ctypes as C
import gzip

import numpy as np
import torch

from . import _lib


def read_fasta(path):
    """Minimal FASTA reader -> {name: bytes} in file order (record id = first word of the header,
    like Bio.SeqIO's record.id used as dict key at preprocessing.py:836)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    seqs, name, parts = {}, None, []
    with opener(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    seqs[name] = b"".join(parts)
                name = line[1:].split()[0].decode() if len(line) > 1 and line[1:].split() else ""
                if name in seqs:
                    raise ValueError("Duplicate key '%s'" % name)      # SeqIO.to_dict raises ValueError too
                parts = []
            else:
                parts.append(line.strip())
    if name is not None:
        seqs[name] = b"".join(parts)
    return seqs


class PackedGenome:
    """Device-resident packed genome.  `seqs`: {name: str|bytes} (insertion order = chromosome index)."""

    def __init__(self, seqs, device=None):
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("mural_b200.PackedGenome needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.names = list(seqs)
        self.chrom_index = {n: i for i, n in enumerate(self.names)}
        bufs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs.values()]
        self.lengths = np.array([len(b) for b in bufs], dtype=np.int64)
        arr = (C.c_char_p * len(bufs))(*bufs)
        lens = (C.c_int64 * len(bufs))(*[len(b) for b in bufs])
        h = C.c_void_p()
        _lib.check(L.mural_genome_create(len(bufs), arr, lens, self.device.index or 0, C.byref(h)))
        self._h = h

    @classmethod
    def from_fasta(cls, path, device=None):
        """FASTA (plain or .gz) -> packed genome without a Python copy of the sequences: the C library's reader
        (mural_fasta_read) hands its buffers straight to mural_genome_create."""
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("mural_b200.PackedGenome needs a CUDA device (no CPU fallback)")
        f = C.c_void_p()
        _lib.check(L.mural_fasta_read(str(path).encode(), C.byref(f)))
        try:
            self = cls.__new__(cls)
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            n = int(L.mural_fasta_n(f))
            self.names = [L.mural_fasta_name(f, i).decode() for i in range(n)]
            self.chrom_index = {nm: i for i, nm in enumerate(self.names)}
            self.lengths = np.array([L.mural_fasta_len(f, i) for i in range(n)], dtype=np.int64)
            arr = (C.c_void_p * n)(*[L.mural_fasta_seq(f, i) for i in range(n)])
            lens = (C.c_int64 * n)(*[int(x) for x in self.lengths])
            h = C.c_void_p()
            _lib.check(L.mural_genome_create(n, C.cast(arr, C.POINTER(C.c_char_p)), lens, self.device.index or 0, C.byref(h)))
            self._h = h
        finally:
            L.mural_fasta_destroy(f)
        return self

    @property
    def handle(self):
        return self._h

    @property
    def device_bytes(self):
        return int(_lib.lib().mural_genome_device_bytes(self._h))

    @property
    def n_exception_runs(self):
        return int(_lib.lib().mural_genome_n_exception_runs(self._h))

    def exception_runs(self):
        """(chrom index, start, end) numpy arrays of the non-ACGT runs (end exclusive), sorted by chromosome then start."""
        L = _lib.lib()
        n = int(L.mural_genome_n_exception_runs(self._h))
        ch, st, en = np.empty(n, np.int32), np.empty(n, np.int64), np.empty(n, np.int64)
        _lib.check(L.mural_genome_exception_runs(self._h, _lib.ptr(ch), _lib.ptr(st), _lib.ptr(en)))
        return ch, st, en

    def windows_with_exceptions(self, chrom, pos, radius, model_type="snv"):
        """Boolean mask: the expanded window of the site contains a non-ACGT symbol or overhangs its chromosome (the reference
        imputes N there, preprocessing.py:791-800).  chrom: genome chromosome index per site; pos: BED start."""
        chrom, pos = np.asarray(chrom, dtype=np.int64), np.asarray(pos, dtype=np.int64)
        lo = pos - radius + (1 if model_type == "indel" else 0)
        hi = pos + radius + (0 if model_type == "indel" else 1)                     # exclusive
        out = (lo < 0) | (hi > self.lengths[chrom])
        ech, est, een = self.exception_runs()
        if len(est):
            # per chromosome: first run that ends after the window start; it intersects iff it starts before the window end
            big = int(self.lengths.max()) + 2 * int(radius) + 8
            key_runs_end = ech.astype(np.int64) * big + een
            key_lo = chrom * big + np.maximum(lo, 0)
            i = np.searchsorted(key_runs_end, key_lo, side="right")
            ok = i < len(est)
            ii = np.minimum(i, len(est) - 1)
            out |= ok & (ech[ii] == chrom) & (est[ii] < hi)
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().mural_genome_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- bit-exact encoders (parity surface; the network path never materialises these) ----------
    def encode_local(self, pos, meta, radius, order, model_type="snv"):
        """int64 [n, n_cat] k-mer indices == seq_digit_encoder (preprocessing.py:636-723)."""
        mt = _lib.MODEL_SNV if model_type == "snv" else _lib.MODEL_INDEL
        n_k = 2 * radius + (1 if model_type == "snv" else 0) - (order - 1)
        pos, meta = _dev_i32(pos, self.device), _dev_i32(meta, self.device)
        out = torch.empty((pos.numel(), n_k), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_encode_local(self._h, _lib.ptr(pos), _lib.ptr(meta), pos.numel(), radius, order, mt,
                                                      _lib.ptr(out), _lib.current_stream()))
        return out

    def encode_onehot(self, pos, meta, radius, model_type="snv"):
        """float32 [n, 4, W] == seq_ohe_encoder (preprocessing.py:756-816)."""
        mt = _lib.MODEL_SNV if model_type == "snv" else _lib.MODEL_INDEL
        W = 2 * radius + (1 if model_type == "snv" else 0)
        pos, meta = _dev_i32(pos, self.device), _dev_i32(meta, self.device)
        out = torch.empty((pos.numel(), 4, W), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_encode_onehot(self._h, _lib.ptr(pos), _lib.ptr(meta), pos.numel(), radius, mt,
                                                       _lib.ptr(out), _lib.current_stream()))
        return out


def _dev_i32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(device)
