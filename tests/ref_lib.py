"""ctypes wrapper around oracle/_ref/libref_riccati.so -- the REFERENCE's own Riccati sources (/root/reference/src/riccati,
src/core), compiled unmodified against the Eigen / Robot stand-ins of oracle/shim (oracle/Makefile.ref).  TEST
INFRASTRUCTURE: it pins the oracle.  The library can only be (re)built where /root/reference exists; a prebuilt copy travels
with the tree (git-ignored, not gpurun-ignored)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libref_riccati.so")
REFERENCE = os.environ.get("ROBOTOC_REFERENCE", "/root/reference")

_lib = None


def available():
    return os.path.exists(LIB) or os.path.isdir(os.path.join(REFERENCE, "src", "riccati"))


def build():
    if os.path.isdir(os.path.join(REFERENCE, "src", "riccati")):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "-f", "Makefile.ref", f"REF={REFERENCE}"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    from robotoc_b200._lib import rbt_dims, rbt_stage_ctrl
    L = ctypes.CDLL(LIB)
    c_int, c_dbl, c_vp = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
    L.ref_riccati_batch.argtypes = [ctypes.POINTER(rbt_dims), ctypes.POINTER(rbt_stage_ctrl), c_int, c_dbl, c_int, c_vp, c_vp,
                                    c_vp, c_vp]
    L.ref_unconstr_batch.argtypes = [c_int, c_int, c_dbl, c_int, c_vp, c_vp, c_vp, c_vp]
    L.ref_version.restype = ctypes.c_char_p
    _lib = L
    return L


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def riccati_batch(dims, L, ctrl, kkt, dx0, max_dts0=0.1):
    """robotoc::RiccatiRecursion (reference code) over a batch, one OCP after the other.
    Returns (kkt mutated like the reference mutates it, ric, dir)."""
    lib = load()
    batch, n_grid = kkt.shape[0], kkt.shape[1]
    kk = kkt.copy()
    ric = np.zeros((batch, n_grid, L.r_stride))
    d = np.zeros((batch, n_grid, L.d_stride))
    cd = dims.c()
    rc = lib.ref_riccati_batch(ctypes.byref(cd), ctrl, n_grid, max_dts0, batch, _ptr(kk), _ptr(ric), _ptr(dx0), _ptr(d))
    assert rc == 0
    return kk, ric, d


def unconstr_batch(nv, UL, N, dt, kkt, dx0):
    lib = load()
    batch = kkt.shape[0]
    kk = kkt.copy()
    ric = np.zeros((batch, N + 1, UL.r_stride))
    d = np.zeros((batch, N + 1, UL.d_stride))
    rc = lib.ref_unconstr_batch(nv, N, dt, batch, _ptr(kk), _ptr(ric), _ptr(dx0), _ptr(d))
    assert rc == 0
    return kk, ric, d
