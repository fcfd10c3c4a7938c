"""GPU: window / k-mer encoders from the packed genome are bit-exact with the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from oracle import encode_np as E

pytestmark = pytest.mark.gpu


def _sites(z, genome, central):
    from mural_b200.data import pack_meta, segment_order
    perm, sizes = segment_order(z["chrom"], z["start"], z["strand"], central)
    pos = z["start"][perm].astype(np.int32)
    meta = pack_meta(z["strand"][perm], np.zeros(len(perm), int), z["chrom"][perm])
    return perm, pos, meta


@pytest.mark.parametrize("ci", range(6))
def test_encoders_bit_exact_vs_reference_fixture(kat, cuda_genome, ci):
    z, genome = kat
    mt, R_l, order, R_d, central = [int(v) for v in z["case%d_cfg" % ci]]
    mt = "snv" if mt == 0 else "indel"
    perm, pos, meta = _sites(z, genome, central)
    assert np.array_equal(perm, z["case%d_perm" % ci])
    cat = cuda_genome.encode_local(pos, meta, R_l, order, mt).cpu().numpy()
    assert cat.dtype == np.int64 and np.array_equal(cat, z["case%d_cat" % ci])
    rows = z["case%d_oh_rows" % ci]
    oh = cuda_genome.encode_onehot(pos[rows], meta[rows], R_d, mt).cpu().numpy()
    assert np.array_equal(oh.view(np.uint32), z["case%d_oh" % ci].view(np.uint32))


def test_encoders_vs_oracle_random_genome():
    """Larger seeded case incl. dense IUPAC noise, sizes the oracle finishes in seconds."""
    from mural_b200 import PackedGenome, pack_meta
    rng = np.random.default_rng(123)
    alpha = np.frombuffer(b"ACGTacgtNnRYMSWKBDHV", dtype=np.uint8)
    p = np.r_[np.full(8, 0.118), np.full(12, 0.056 / 12)]
    seqs = {"c%d" % i: alpha[rng.choice(len(alpha), n, p=p / p.sum())].tobytes() for i, n in enumerate((70001, 333, 64, 1))}
    g = PackedGenome(seqs)
    assert g.n_exception_runs > 100
    for ci, (name, s) in enumerate(seqs.items()):
        sym = E.seq_to_symbols(s)
        n = 400
        pos = rng.integers(0, len(s), n).astype(np.int32)
        strand = rng.integers(0, 2, n)
        meta = pack_meta(strand, rng.integers(0, 4, n), np.full(n, ci))
        for mt, R_l, order in (("snv", 7, 3), ("snv", 10, 1), ("indel", 5, 3), ("snv", 3, 6)):
            got = g.encode_local(pos, meta, R_l, order, mt).cpu().numpy()
            assert np.array_equal(got, E.kmer_windows(sym, pos, strand, R_l, order, mt)), (name, mt, R_l, order)
        for mt, R_d in (("snv", 150), ("indel", 64), ("snv", 1000)):
            got = g.encode_onehot(pos[:64], meta[:64], R_d, mt).cpu().numpy()
            exp = E.onehot_windows(sym, pos[:64], strand[:64], R_d, mt)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (name, mt, R_d)


def test_encode_edge_cases(cuda_genome):
    from mural_b200 import PackedGenome
    empty = np.zeros(0, np.int32)
    assert cuda_genome.encode_local(empty, empty, 7, 3).shape == (0, 13)          # empty input
    assert cuda_genome.encode_onehot(empty, empty, 100).shape == (0, 4, 201)
    with pytest.raises(KeyError):                                                   # reference: dict KeyError
        PackedGenome({"c": "ACGTX"})
    with pytest.raises(RuntimeError):
        cuda_genome.encode_local(np.zeros(1, np.int32), np.zeros(1, np.int32), 1, 5)   # window shorter than k-mer


def test_onehot_roundtrip_symbols(kat, cuda_genome):
    """one-hot -> symbols -> the symbols the oracle sees (the drop-in tensor path relies on it)."""
    import ctypes as C
    from mural_b200 import _lib
    z, genome = kat
    perm, pos, meta = _sites(z, genome, 5000)
    oh = cuda_genome.encode_onehot(pos[:200], meta[:200], 300)
    sym = torch.empty((200, 601), dtype=torch.uint8, device=oh.device)
    _lib.check(_lib.lib().mural_onehot_to_symbols(_lib.ptr(oh), 200, 601, _lib.ptr(sym), _lib.current_stream()))
    sym = sym.cpu().numpy()
    names = list(genome)
    comp = np.array([3, 2, 1, 0, 5, 4, 9, 7, 8, 6, 13, 12, 11, 10, 14], dtype=np.uint8)
    for i in range(200):
        c = int(meta[i]) >> 8
        ws = E.window_symbols(E.seq_to_symbols(genome[names[c]]), pos[i:i + 1], np.array([meta[i] & 1]), 300)[0]
        exp = comp[ws] if meta[i] & 1 else ws
        assert np.array_equal(sym[i], exp)
