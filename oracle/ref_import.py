"""TEST INFRASTRUCTURE ONLY (oracle/): import the *unmodified* reference from /root/reference.

Only usable inside the build container (the GPU box has no /root/reference).  It is used by
oracle/make_golden.py to (a) pin the numpy/torch restatement in oracle/ against the real reference
classes and (b) generate the committed fixtures under tests/golden/.

The reference imports a handful of I/O-only third-party modules that are not installed here
(SURVEY.md §8c).  None of them touches the arithmetic of the hot path, so they are replaced by
empty stub modules before `import MuRaL...`.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MURAL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "MuRaL"))


class _BedTool(list):
    """Duck-typed stand-in for pybedtools.BedTool: a list of region objects."""
    fn = "<memory>"


class Region:
    """Duck-typed pybedtools.Interval (fields used by preprocessing.py:60-101,586,753)."""
    __slots__ = ("chrom", "start", "stop", "end", "strand", "score", "name")

    def __init__(self, chrom, start, stop, strand, score=0, name="."):
        self.chrom, self.start, self.stop, self.end = chrom, int(start), int(stop), int(stop)
        self.strand, self.score, self.name = strand, score, name


class SeqRec:
    """Duck-typed Bio.SeqRecord: only `.seq` is read (preprocessing.py:458,964,990)."""
    def __init__(self, seq):
        self.seq = seq


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "pybedtools" not in sys.modules:
        _stub("pybedtools", BedTool=_BedTool)
    if "Bio" not in sys.modules:
        bio = _stub("Bio")
        bio.SeqIO = _stub("Bio.SeqIO")
        bio.Seq = _stub("Bio.Seq", Seq=str)
    for n in ("pyBigWig", "h5py", "prettytable", "pysam"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n, PrettyTable=object)
    if "jax" not in sys.modules:
        try:
            import jax  # noqa: F401
        except Exception:
            cfg = types.SimpleNamespace(update=lambda *a, **k: None)
            jx = _stub("jax", config=cfg, grad=None, hessian=None, jit=lambda f: f)
            jx.numpy = _stub("jax.numpy")
            _stub("jax.config", config=cfg, update=cfg.update)
    if "dirichletcal" not in sys.modules:
        class _Cal:  # unpickle target; only calibrator_.weights_ is read
            def __setstate__(self, st):
                self.__dict__.update(st)
        d = _stub("dirichletcal")
        d.calib = _stub("dirichletcal.calib")
        for sub, cls in (("fulldirichlet", "FullDirichletCalibrator"), ("multinomial", "MultinomialRegression"),
                         ("vectorscaling", "VectorScaling"), ("tempscaling", "TemperatureScaling")):
            m = _stub("dirichletcal.calib." + sub, **{cls: type(cls, (_Cal,), {})})
            setattr(d.calib, sub, m)
    if "ray" not in sys.modules:
        try:
            import ray  # noqa: F401
        except Exception:
            r = _stub("ray")
            r.tune = _stub("ray.tune")


def import_reference():
    """Returns (preprocessing, model_snv, model_indel, nn_utils) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    # nn_utils first: model_snv <-> nn_utils import each other and only this order resolves
    # (it is the order the reference's own pipelines use, scripts/run_predict.py:18-24).
    nnu = importlib.import_module("MuRaL.model.nn_utils")
    snv = importlib.import_module("MuRaL.model.model_snv")
    indel = importlib.import_module("MuRaL.model.model_indel")
    pre = importlib.import_module("MuRaL.data.preprocessing")
    return pre, snv, indel, nnu
