"""ORACLE -- TEST INFRASTRUCTURE ONLY: numpy front-end of oracle/pointnet2_oracle.c.

Compiles the C restatement with gcc (-O2 -ffp-contract=off) into oracle/_build/ on first use and
exposes one function per reference kernel, taking/returning numpy arrays with the reference's
shapes and dtypes (float32 / int32).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pointnet2_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libpointnet2_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11", _SRC, "-o", _OUT, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_emd_forward.restype = ctypes.c_int
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, dtype):
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n):
    return lib().oracle_opt_n_threads(int(n))


def gather_points(points, idx):
    points, pp = _f(points); idx, ip = _i(idx)
    b, c, n = points.shape; m = idx.shape[1]
    out, op = _out((b, c, m), np.float32)
    lib().oracle_gather_points(b, c, n, m, pp, ip, op)
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out); idx, ip = _i(idx)
    b, c, m = grad_out.shape
    out, op = _out((b, c, n), np.float32)
    lib().oracle_gather_points_grad(b, c, n, m, gp, ip, op)
    return out


def furthest_point_sampling(xyz, m, return_temp=False):
    xyz, xp = _f(xyz)
    b, n, _ = xyz.shape
    idx, ip = _out((b, m), np.int32)
    temp, tp = _out((b, n), np.float32)
    lib().oracle_furthest_point_sampling(b, n, m, xp, tp, ip)
    return (idx, temp) if return_temp else idx


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, qp = _f(new_xyz); xyz, xp = _f(xyz)
    b, m, _ = new_xyz.shape; n = xyz.shape[1]
    idx, ip = _out((b, m, nsample), np.int32)
    lib().oracle_query_ball_point(b, n, m, ctypes.c_float(radius), nsample, qp, xp, ip)
    return idx


def group_points(points, idx):
    points, pp = _f(points); idx, ip = _i(idx)
    b, c, n = points.shape; _, npoints, nsample = idx.shape
    out, op = _out((b, c, npoints, nsample), np.float32)
    lib().oracle_group_points(b, c, n, npoints, nsample, pp, ip, op)
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out); idx, ip = _i(idx)
    b, c, npoints, nsample = grad_out.shape
    out, op = _out((b, c, n), np.float32)
    lib().oracle_group_points_grad(b, c, n, npoints, nsample, gp, ip, op)
    return out


def three_nn(unknown, known):
    unknown, up = _f(unknown); known, kp = _f(known)
    b, n, _ = unknown.shape; m = known.shape[1]
    d2, dp = _out((b, n, 3), np.float32)
    idx, ip = _out((b, n, 3), np.int32)
    lib().oracle_three_nn(b, n, m, up, kp, dp, ip)
    return d2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points); idx, ip = _i(idx); weight, wp = _f(weight)
    b, c, m = points.shape; n = idx.shape[1]
    out, op = _out((b, c, n), np.float32)
    lib().oracle_three_interpolate(b, c, m, n, pp, ip, wp, op)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, gp = _f(grad_out); idx, ip = _i(idx); weight, wp = _f(weight)
    b, c, n = grad_out.shape
    out, op = _out((b, c, m), np.float32)
    lib().oracle_three_interpolate_grad(b, c, n, m, gp, ip, wp, op)
    return out


def chamfer_forward(xyz1, xyz2):
    xyz1, p1 = _f(xyz1); xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape; m = xyz2.shape[1]
    d1, d1p = _out((b, n), np.float32); d2, d2p = _out((b, m), np.float32)
    i1, i1p = _out((b, n), np.int32); i2, i2p = _out((b, m), np.int32)
    lib().oracle_chamfer_forward(b, n, p1, m, p2, d1p, d2p, i1p, i2p)
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, p1 = _f(xyz1); xyz2, p2 = _f(xyz2); idx1, i1p = _i(idx1); idx2, i2p = _i(idx2)
    g1, g1p = _f(g1); g2, g2p = _f(g2)
    b, n, _ = xyz1.shape; m = xyz2.shape[1]
    o1, o1p = _out((b, n, 3), np.float32); o2, o2p = _out((b, m, 3), np.float32)
    lib().oracle_chamfer_backward(b, n, p1, m, p2, i1p, i2p, g1p, g2p, o1p, o2p)
    return o1, o2


def emd_forward(xyz1, xyz2, eps, iters):
    xyz1, p1 = _f(xyz1); xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    dist, dp = _out((b, n), np.float32)
    ass, ap = _out((b, n), np.int32)
    rounds = lib().oracle_emd_forward(b, n, p1, p2, dp, ap, ctypes.c_float(eps), int(iters))
    return dist, ass, rounds
