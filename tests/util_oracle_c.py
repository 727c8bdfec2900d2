"""ctypes access to the C restatement of the sampler (oracle/sampler_ref.c, TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ROOT, "oracle", "_build", "libsampler_ref.so")
        # always ask make: a no-op when oracle/_build is newer than the sources, a rebuild after an edit
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
        _lib = C.CDLL(so)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def upsample_round(o, d, z, sdf, n_new, inv_s):
    R, m = z.shape
    o, d, z, sdf = (np.ascontiguousarray(x, np.float32) for x in (o, d, z, sdf))
    cdf = np.zeros((R, m), np.float32)
    z_new = np.zeros((R, n_new), np.float32)
    zm = np.zeros((R, m + n_new), np.float32)
    inds = np.zeros((R, n_new), np.int32)
    order = np.zeros((R, m + n_new), np.int32)
    lib().nrw_ref_upsample_round(C.c_int(R), C.c_int(m), C.c_int(n_new), C.c_float(inv_s), _p(o), _p(d), _p(z), _p(sdf),
                                 _p(cdf), _p(z_new), _p(zm), _p(inds), _p(order))
    return z_new, zm, inds, order, cdf


def coarse(n_samples, n_outside, near, far, s_near=None, s_far=None, u_ray=None, u_out=None):
    R = len(near)
    near, far = np.ascontiguousarray(near, np.float32), np.ascontiguousarray(far, np.float32)
    c = lambda x: None if x is None else np.ascontiguousarray(x, np.float32)
    s_near, s_far, u_ray, u_out = c(s_near), c(s_far), c(u_ray), c(u_out)
    z = np.zeros((R, n_samples), np.float32)
    zo = np.zeros((R, max(n_outside, 1)), np.float32)
    sd = np.zeros(R, np.float32)
    lib().nrw_ref_coarse(C.c_int(R), C.c_int(n_samples), C.c_int(n_outside), C.c_int(int(u_ray is not None)), _p(near), _p(far),
                         _p(s_near), _p(s_far), _p(u_ray), _p(u_out), _p(z), _p(zo), _p(sd))
    return z, zo[:, :n_outside], sd


def boundary(near, far, z, nb):
    """renderer.py:546-566 -> sorted [R, S0+nb]"""
    R, S0 = z.shape
    near, far, z = (np.ascontiguousarray(x, np.float32) for x in (near, far, z))
    out = np.zeros((R, S0 + nb), np.float32)
    lib().nrw_ref_boundary(C.c_int(R), C.c_int(S0), C.c_int(nb), _p(near), _p(far), _p(z), _p(out))
    return out
