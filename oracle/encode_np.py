"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy) of the reference's window/k-mer encoders.

Nothing under mural_b200/ may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do, and only as the checker / CPU baseline.

Pinned against the real reference (MuRaL/data/preprocessing.py imported from /root/reference) by
oracle/make_golden.py, which also writes tests/golden/encode_*.npz; tests/test_oracle_golden.py
re-checks this file against those fixtures on every run.

Each function cites the reference lines it restates.  Semantics are per *site*: the reference merges
overlapping windows of consecutive sites into runs before encoding (preprocessing.py:571-615) but the
run is only a cache — every site ends up with the window [start-R, stop+R) of its chromosome,
'N'-imputed outside [0, len) (:681-695, :790-804), upper-cased (:694, :802) and reverse-complemented
for '-' (:700, :813-814).
"""
import numpy as np

# Symbol alphabet shared by the whole project: 0..3 = A,C,G,T ; 4..14 = IUPAC (preprocessing.py:758-772)
SYMBOLS = "ACGTRYMSWKBDHVN"
SYM_N = 14
# one_hot_encoder, '+' strand (preprocessing.py:758-772); channel order A,C,G,T
ONEHOT = np.array([
    [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1],
    [.5, 0, .5, 0], [0, .5, 0, .5], [.5, .5, 0, 0], [0, .5, .5, 0], [.5, 0, 0, .5], [0, 0, .5, .5],
    [0, 1 / 3, 1 / 3, 1 / 3], [1 / 3, 0, 1 / 3, 1 / 3], [1 / 3, 1 / 3, 0, 1 / 3], [1 / 3, 1 / 3, 1 / 3, 0],
    [.25, .25, .25, .25]], dtype=np.float32)
# one_hot_encoder_rc (preprocessing.py:774-788) is exactly the channel-reversed table.
ONEHOT_RC = ONEHOT[:, ::-1].copy()

_ASCII2SYM = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(SYMBOLS):
    _ASCII2SYM[ord(_c)] = _i
    _ASCII2SYM[ord(_c.lower())] = _i       # .upper() at preprocessing.py:694,802


def seq_to_symbols(seq) -> np.ndarray:
    """str/bytes -> uint8 symbol codes.  Any other character is a KeyError in the reference
    (dict lookup at preprocessing.py:698,809); here it raises KeyError as well."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    sym = _ASCII2SYM[np.frombuffer(seq, dtype=np.uint8)]
    if (sym == 255).any():
        bad = np.frombuffer(seq, dtype=np.uint8)[sym == 255][0]
        raise KeyError(chr(bad))
    return sym


def window_length(radius: int, model_type: str = "snv") -> int:
    """calc_cat_n with local_order=1 (preprocessing.py:382-385): 2R+1 for SNV, 2R for INDEL."""
    return 2 * radius + (1 if model_type == "snv" else 0)


def window_symbols(sym: np.ndarray, pos: np.ndarray, strand: np.ndarray, radius: int,
                   model_type: str = "snv") -> np.ndarray:
    """Oriented symbol windows [n, W] for sites on ONE chromosome.

    pos = BED start (0-based); strand: 0 '+', 1 '-'.
    Window start: SNV start-R ; INDEL start-R+1 (extend_interval, preprocessing.py:559-567).
    Outside the chromosome -> 'N' (:681-695).  '-' strand: reversed, complemented: for ACGT the
    reference's digit_encoder_rc is 3-x (:668); for IUPAC the one-hot rc table is the channel flip
    which is handled in onehot_windows(), here the symbol is kept and only the order is reversed
    while ACGT are complemented (R<->Y etc. are *not* remapped to symbols because the reference never
    forms complemented IUPAC symbols, it looks vectors up in a second table).
    """
    W = window_length(radius, model_type)
    w0 = pos.astype(np.int64) - radius + (0 if model_type == "snv" else 1)
    q = w0[:, None] + np.arange(W, dtype=np.int64)[None, :]
    inside = (q >= 0) & (q < sym.shape[0])
    out = np.full(q.shape, SYM_N, dtype=np.uint8)
    out[inside] = sym[q[inside]]
    neg = strand.astype(bool)
    if neg.any():
        out[neg] = out[neg, ::-1]
    return out


def onehot_windows(sym, pos, strand, radius, model_type="snv") -> np.ndarray:
    """seq_ohe_encoder + distal_encoding_by_region (preprocessing.py:756-816, 978-999):
    float32 [n, 4, W], channel order A,C,G,T, '-' strand = reversed order + rc table."""
    ws = window_symbols(sym, pos, strand, radius, model_type)
    neg = strand.astype(bool)
    out = np.empty(ws.shape + (4,), dtype=np.float32)
    out[~neg] = ONEHOT[ws[~neg]]
    out[neg] = ONEHOT_RC[ws[neg]]
    return np.ascontiguousarray(out.transpose(0, 2, 1))


def kmer_windows(sym, pos, strand, radius, order, model_type="snv") -> np.ndarray:
    """seq_digit_encoder + process_local_seq_* (preprocessing.py:636-723, 479-522):
    int64 [n, W-(order-1)]; base codes A0 C1 G2 T3 (rc: 3-x), k-mer index most-significant base
    first (:710); any non-ACGT base in the k-mer -> 4**order (:707-708, :722).
    order == 1: the raw encoder gives 4 for non-ACGT (:722 with 4**1); afterwards
    process_local_seq_snv/indel (:486,490) maps only negatives to 0, so 4 stays 4."""
    ws = window_symbols(sym, pos, strand, radius, model_type).astype(np.int64)
    neg = strand.astype(bool)
    valid = ws < 4
    code = np.where(valid, ws, 0)
    code[neg] = 3 - code[neg]
    code = np.where(valid, code, -1)
    W = ws.shape[1]
    n_k = W - (order - 1)
    idx = np.zeros((ws.shape[0], n_k), dtype=np.int64)
    bad = np.zeros((ws.shape[0], n_k), dtype=bool)
    for d in range(order):
        c = code[:, d:d + n_k]
        bad |= c < 0
        idx = idx * 4 + np.where(c < 0, 0, c)
    return np.where(bad, 4 ** order, idx)


# ----------------------------------------------------------------------------------------------
# Sample order: bed_reader (preprocessing.py:39-106)
# ----------------------------------------------------------------------------------------------
def bed_batches(chrom_ids, starts, strands, central_bp):
    """Literal restatement of bed_reader's state machine.  Returns a list of (indices, strand) in
    yield order.  chrom_ids: any comparable per-site chromosome key; strands: 0 '+', 1 '-'.
    Pure-Python loop: use for small/medium inputs; order_sites() is the vectorised equivalent."""
    out = []
    pos_l, neg_l = [], []
    init = False
    chrom = None
    end0 = 0
    for i in range(len(starts)):
        c, s = chrom_ids[i], int(starts[i])
        if not init:                                  # :61-66
            init = True
            chrom = c
            end0 = s + central_bp
        if c != chrom:                                # :70-79
            if pos_l:
                out.append((pos_l, 0)); pos_l = []
            if neg_l:
                out.append((neg_l, 1)); neg_l = []
            chrom = c
            end0 = 1 + central_bp
        if s > end0:                                  # :83-94
            if pos_l:
                out.append((pos_l, 0)); pos_l = []
            if neg_l:
                out.append((neg_l, 1)); neg_l = []
            while s > end0:
                end0 += central_bp
        (pos_l if strands[i] == 0 else neg_l).append(i)   # :97-101
    if pos_l:
        out.append((pos_l, 0))
    if neg_l:
        out.append((neg_l, 1))
    return out


def order_sites(chrom_ids, starts, strands, central_bp):
    """Vectorised bed_reader: returns (perm, batch_sizes) with perm = site indices in emission
    order and batch_sizes = sizes of the successive (segment, strand) batches."""
    chrom_ids = np.asarray(chrom_ids)
    starts = np.asarray(starts, dtype=np.int64)
    strands = np.asarray(strands, dtype=np.int64)
    n = len(starts)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    newblk = np.ones(n, dtype=bool)
    newblk[1:] = chrom_ids[1:] != chrom_ids[:-1]
    blk = np.cumsum(newblk) - 1
    first_idx = np.flatnonzero(newblk)
    base = np.ones(len(first_idx), dtype=np.int64)          # start0 = 1 for later chroms (:78)
    base[0] = starts[0]                                     # first chrom anchors at first site (:63)
    b = base[blk]
    j = np.maximum(0, -((-(starts - b)) // central_bp) - 1)  # smallest j with start <= b+(j+1)c
    # end0 never moves backwards inside a block (stateful loop) -> running max per block
    key = blk * (j.max() + 2) + j
    key = np.maximum.accumulate(key)
    full = key * 2 + strands
    perm = np.argsort(full, kind="stable")
    fs = full[perm]
    cuts = np.flatnonzero(np.r_[True, fs[1:] != fs[:-1], True])
    return perm, np.diff(cuts)
