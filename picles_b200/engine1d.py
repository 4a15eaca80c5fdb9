"""ctypes front of the picles1d_* entry points (include/picles_b200.h): one handle = the one-dimensional model
on one GPU.  No CPU fallback: construction raises when the library or a B200 is missing."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import PiclesCounters, PiclesError, load_library


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class B200Engine1D:
    def __init__(self, Nx, xmin, dx, x_nodes, params, device=0):
        self.lib = load_library()
        self.Nx = int(Nx)
        self.h = C.c_void_p()
        rc = self.lib.picles1d_create(C.byref(self.h), int(device))
        if rc != 0:
            raise PiclesError(f"picles1d_create failed ({rc}): {self.lib.picles1d_last_error(None).decode()}")
        xn = np.ascontiguousarray(x_nodes, dtype=np.float64)
        if xn.shape != (self.Nx,):
            raise ValueError("x_nodes must have Nx elements")
        self._ck(self.lib.picles1d_set_grid(self.h, self.Nx, float(xmin), float(dx), _ptr(xn)))
        self._ck(self.lib.picles1d_set_params(self.h, C.byref(params)))

    def _ck(self, rc):
        if rc != 0:
            raise PiclesError(f"picles1d call failed ({rc}): {self.lib.picles1d_last_error(self.h).decode()}")

    def _plane(self, a):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (self.Nx,)))
        return a

    def seed(self, u0):
        u0 = self._plane(u0)
        self._ck(self.lib.picles1d_seed(self.h, _ptr(u0)))

    def step(self, t, DT, u_t, u_t1):
        a, b = self._plane(u_t), self._plane(u_t1)
        self._ck(self.lib.picles1d_step(self.h, float(t), float(DT), _ptr(a), _ptr(b)))

    def state(self):
        """State as planes: shape (3, Nx) = [e, m_x, 0]."""
        S = np.empty((3, self.Nx), np.float64)
        self._ck(self.lib.picles1d_get_state(self.h, _ptr(S)))
        return S

    def particles(self):
        z = np.empty((3, self.Nx), np.float64)
        t = np.empty(self.Nx, np.float64)
        dt = np.empty(self.Nx, np.float64)
        flags = np.empty(self.Nx, np.uint8)
        status = np.empty(self.Nx, np.int32)
        self._ck(self.lib.picles1d_get_particles(self.h, _ptr(z), _ptr(t), _ptr(dt), _ptr(flags), _ptr(status)))
        return dict(z=z, t=t, dt=dt, flags=flags, status=status)

    def counters(self):
        c = PiclesCounters()
        self._ck(self.lib.picles1d_get_counters(self.h, C.byref(c)))
        return {name: getattr(c, name) for name, _ in c._fields_}

    def close(self):
        if self.h:
            self.lib.picles1d_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
