"""ctypes driver of the host build of waveome_b200/csrc/wv_lbfgsb.h (test infrastructure)."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "waveome_b200", "_lib",
                     "libwv_lbfgsb_host.so")
TASKS = {0: "FG", 1: "CONV_PG", 2: "CONV_F", 3: "ABNORMAL", 4: "MAXITER", 5: "MAXFUN"}


def _lib():
    lib = C.CDLL(_PATH)
    lib.wvh_lb_create.restype = C.c_void_p
    lib.wvh_lb_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.wvh_lb_destroy.argtypes = [C.c_void_p]
    lib.wvh_lb_start.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.wvh_lb_step.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    lib.wvh_lb_step.restype = C.c_int
    lib.wvh_lb_x.argtypes = [C.c_void_p]; lib.wvh_lb_x.restype = C.POINTER(C.c_double)
    lib.wvh_lb_f.argtypes = [C.c_void_p]; lib.wvh_lb_f.restype = C.c_double
    lib.wvh_lb_iter.argtypes = [C.c_void_p]; lib.wvh_lb_iter.restype = C.c_int
    lib.wvh_lb_nfev.argtypes = [C.c_void_p]; lib.wvh_lb_nfev.restype = C.c_int
    return lib


def minimize(fun, x0, maxcor=10, maxiter=15000, maxfun=15000, maxls=20, ftol=2.220446049250313e-09, gtol=1e-5,
             trace=None):
    """Drive the state machine exactly as wv_batch_fit_lbfgs does on the device."""
    lib = _lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    P = x0.size
    h = lib.wvh_lb_create(P, maxcor, maxiter, maxfun, maxls, ftol, gtol)
    try:
        lib.wvh_lb_start(h, x0.ctypes.data_as(C.POINTER(C.c_double)))
        task = 0
        while task == 0:
            x = np.ctypeslib.as_array(lib.wvh_lb_x(h), shape=(P,)).copy()
            f, g = fun(x)
            if trace is not None:
                trace.append((x.copy(), float(f)))
            g = np.ascontiguousarray(g, dtype=np.float64)
            task = lib.wvh_lb_step(h, float(f), g.ctypes.data_as(C.POINTER(C.c_double)))
        x = np.ctypeslib.as_array(lib.wvh_lb_x(h), shape=(P,)).copy()
        return dict(x=x, f=lib.wvh_lb_f(h), nit=lib.wvh_lb_iter(h), nfev=lib.wvh_lb_nfev(h), task=TASKS.get(task, task))
    finally:
        lib.wvh_lb_destroy(h)
