"""Phase-timing run of the stage kernels (scratch/libmural_timing.so built by scratch/build_timing.sh)."""
import ctypes as C, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mural_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "scratch", "libmural_timing.so")
sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--sites-per-step", "262144"]
import bench
import atexit
def dump():
    lib = L.lib()
    lib.mural_tc_timing_dump.restype = C.c_int
    lib.mural_tc_timing_dump()
atexit.register(dump)
bench.main()
