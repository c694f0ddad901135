"""TEST INFRASTRUCTURE ONLY: host emulation of the CUDA liftover pipeline (tests/emul).  The product's per-index device
bodies are compiled for the CPU under a one-lane SIMT shim and driven through the same C-ABI (prefix ptl_emul_), so the
CPU suite can compare the kernel logic with the oracle bit for bit.  Never imported by portello_b200/ or bench.py."""
import ctypes as C
import os
import subprocess

from portello_b200 import abi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
SO = os.path.join(HERE, "_build", "libptl_emul.so")


def build() -> str:
    if os.environ.get("PTL_EMUL_SO"):  # another build of the same sources (e.g. -fsanitize=address, see tools/fuzz/README.md)
        return os.environ["PTL_EMUL_SO"]
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"emulation build failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return SO


class EmulLib(abi.LiftLib):
    def __init__(self, dll):
        super().__init__(dll, "ptl_emul_")
        dll.ptl_emul_get_segment_table_gaps.restype = C.c_int
        dll.ptl_emul_get_segment_table_gaps.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, abi.u32p]
        dll.ptl_emul_slot_counters.restype = C.c_int
        dll.ptl_emul_slot_counters.argtypes = [C.c_void_p, C.c_int, abi.u64p]


_lib = None


def load() -> EmulLib:
    global _lib
    if _lib is None:
        _lib = EmulLib(C.CDLL(build()))
    return _lib
