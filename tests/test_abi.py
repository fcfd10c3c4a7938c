"""CPU: the C-ABI library loads and exports exactly what include/mural_b200.h declares."""
import os
import re

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mural_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mural_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from mural_b200 import _lib
    L = _lib.lib()                       # raises if the .so is missing (no fallback)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmural_b200.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
    for n in _lib.PROTOTYPES:
        assert n in names, "%s is bound but not declared in include/mural_b200.h" % n
    assert L.mural_abi_version() == 2


def test_error_channel_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    import ctypes as C
    from mural_b200 import _lib
    L = _lib.lib()
    cfg = _lib.SnvConfig(7, 3, 50, 150, 75, 32, 3, 4)          # L = 101 <= 200
    h = C.c_void_p()
    assert L.mural_snv_model_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"distal seq len must be >200" in L.mural_last_error()
    cfg = _lib.SnvConfig(7, 3, 1000, 150, 75, 24, 3, 4)        # unsupported channel count
    assert L.mural_snv_model_create(C.byref(cfg), 0, C.byref(h)) != 0
    cfg = _lib.SnvConfig(10, 3, 1000, 150, 75, 32, 3, 4)
    assert L.mural_snv_model_create(C.byref(cfg), 0, C.byref(h)) == 0
    assert L.mural_snv_model_n_trainable(h) > 0
    names = []
    for i in range(L.mural_snv_model_n_tensors(h)):
        nm, off, num, buf = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
        assert L.mural_snv_model_tensor(h, i, C.byref(nm), C.byref(off), C.byref(num), C.byref(buf)) == 0
        names.append(nm.value.decode())
    assert "lin_layers.0.weight" in names and "RBs2_2.1.bn2.running_var" in names
    L.mural_snv_model_destroy(h)
