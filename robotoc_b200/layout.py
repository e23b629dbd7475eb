"""Record layouts (mirror of include/rbt_layout.h, queried from the C library so there is one source of truth)."""
import ctypes
from dataclasses import dataclass

from . import _lib

_FIELDS = ("nv nu nx ns_max k_Fxx k_Fvu k_Qxx k_Qxu k_Quu k_Fx k_lx k_lu k_Phix k_Phiu k_p k_fx k_hx k_hu k_Phit k_sc "
           "k_stage_size k_core_size k_extra_size k_stride r_P r_s r_K r_k r_M r_m r_Psi r_Phi r_T r_W r_psix r_psiu "
           "r_phix r_phiu r_mt r_mtn r_sc r_dtsdx r_stosc r_core_size r_extra_size r_stride f_F f_H f_G f_lu f_stride "
           "d_dx d_du d_dlmdgmm d_dxi d_dts d_stride").split()
_UFIELDS = ("nv nx k_Qxx k_Qxu k_Qaa k_Fx k_lx k_la k_stride r_P r_s r_K r_k r_stride f_F f_H f_G f_la f_stride "
            "d_dx d_da d_dlmdgmm d_stride").split()


@dataclass(frozen=True)
class Dims:
    """Robot dimensions (reference: src/robot/robot.cpp:33-60)."""
    nv: int
    nu: int
    ns_max: int
    n_passive: int = 0

    @property
    def nx(self):
        return 2 * self.nv

    def c(self):
        return _lib.rbt_dims(self.nv, self.nu, self.ns_max, self.n_passive)


class Layout:
    """Offsets (in doubles) of every block inside the KKT / Riccati / factorized-KKT / direction records."""

    def __init__(self, dims: Dims, getter=None):
        self.dims = dims
        cd = dims.c()
        get = getter or _lib.lib().rbt_layout_get
        for f in _FIELDS:
            v = get(ctypes.byref(cd), f.encode())
            if v < 0:
                raise RuntimeError(f"layout field {f} unknown to the library")
            setattr(self, f, v)


class ULayout:
    def __init__(self, nv: int, getter=None):
        get = getter or _lib.lib().rbt_ulayout_get
        for f in _UFIELDS:
            v = get(nv, f.encode())
            if v < 0:
                raise RuntimeError(f"ulayout field {f} unknown to the library")
            setattr(self, f, v)
