"""ctypes binding of libcgb200.so (include/cgb200.h).  There is no fallback: if the CUDA library has not
been built (python __graft_entry__.py / make -C coulombgas_b200/csrc) importing this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CGB_LIB") or os.path.join(_HERE, "libcgb200.so")   # CGB_LIB: developer A/B builds only

NPARAMS = 1074
MAX_N = 64


class CgbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "coulombgas_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C coulombgas_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, f64, u64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_size_t
    pd2 = ctypes.POINTER(ctypes.c_double)
    lib.cgb_abi_version.restype = i32
    lib.cgb_error_string.restype = ctypes.c_char_p
    lib.cgb_error_string.argtypes = [i32]
    lib.cgb_ktable_twist_host.argtypes = [vp, i32, pd2]
    lib.cgb_ewald_prepare_host.argtypes = [vp, i32, f64, vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.cgb_logp_workspace_bytes.argtypes = [i32, i32, ctypes.POINTER(sz)]
    lib.cgb_mcmc_workspace_bytes.argtypes = [i32, i32, ctypes.POINTER(sz)]
    u32 = ctypes.c_uint
    lib.cgb_logp_f64io.argtypes = [vp, vp, vp, vp, i32, pd2, vp, i32, i32, f64, vp, vp, vp, vp, sz, u32]
    lib.cgb_mcmc_sweep_f64io.argtypes = [vp, vp, vp, vp, i32, pd2, vp, vp, vp, u64, vp, i32, f64, f64, i32, i32, i32,
                                         vp, vp, vp, vp, vp, vp, sz, u32]
    lib.cgb_ewald_f64io.argtypes = [vp, vp, i32, i32, f64, f64, f64, vp, i32, i32, vp, vp]
    lib.cgb_debug_rng.argtypes = [vp, u64, vp, i32, i32, i32, vp, vp, vp, vp, i32]
    lib.cgb_microbench.argtypes = [vp, i32, i32, i32, vp]
    lib.cgb_van_nparams.argtypes = [i32]
    lib.cgb_van_nparams.restype = sz
    lib.cgb_van_sample_f64io.argtypes = [vp, vp, sz, vp, i32, i32, i32, vp, u64, vp, vp, vp, vp]
    lib.cgb_van_logprob_f64io.argtypes = [vp, vp, sz, vp, i32, i32, i32, vp, vp, vp]
    for f in ("cgb_ktable_twist_host", "cgb_ewald_prepare_host", "cgb_logp_workspace_bytes",
              "cgb_mcmc_workspace_bytes", "cgb_logp_f64io", "cgb_mcmc_sweep_f64io", "cgb_ewald_f64io",
              "cgb_microbench", "cgb_van_sample_f64io", "cgb_van_logprob_f64io", "cgb_debug_rng"):
        getattr(lib, f).restype = i32
    return lib


lib = _load()

EXPORTS = ("cgb_abi_version", "cgb_error_string", "cgb_ktable_twist_host", "cgb_ewald_prepare_host",
           "cgb_logp_workspace_bytes", "cgb_mcmc_workspace_bytes", "cgb_logp_f64io", "cgb_mcmc_sweep_f64io",
           "cgb_ewald_f64io", "cgb_microbench", "cgb_van_nparams", "cgb_van_sample_f64io", "cgb_van_logprob_f64io",
           "cgb_debug_rng")

# per-call flags of include/cgb200.h
KERNELS = {None: 0, "auto": 0, "cta": 1, "cta4": 2, "cta8": 3, "warp": 4, "spec4": 5, "spec8": 6, "spec": 5}
FLAG_FORCE_PIVOT, FLAG_ADVANCE_KEY, FLAG_FORCE_F64_SLATER = 0x100, 0x200, 0x400
FLAG_FORCE_SEGMENTS, FLAG_NO_SEGMENTS = 0x800, 0x1000
STATS_LEN = 8


def check(rc, what):
    if rc != 0:
        raise CgbError("%s failed: %s (code %d)" % (what, lib.cgb_error_string(rc).decode(), rc))
