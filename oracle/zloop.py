"""ctypes binding of oracle/zloop.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libzloop.so')


def build():
    src = os.path.join(_HERE, 'zloop.c')
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def _lib():
    lib = ctypes.CDLL(build())
    fp = ctypes.POINTER(ctypes.c_float)
    lib.zl_zigap_z.argtypes = [fp] * 7 + [ctypes.c_long] * 3 + [ctypes.c_int]
    lib.zl_gap_z.argtypes = [fp] * 5 + [ctypes.c_long] * 3
    lib.zl_sparse_z.argtypes = [fp] * 9 + [ctypes.c_long] * 3
    return lib


def _p(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def zigap_z(log_U_hat, log_V_hat, D_hat, X, quirk=True, third=False):
    """zigap.py:79-95 as a sequential float32 loop; returns (DZ_hat_i, DZ_hat_j[, DZ_exp_logsum_hat])."""
    n, K = log_U_hat.shape; p = log_V_hat.shape[0]
    Zi = np.empty((n, K), np.float32); Zj = np.empty((p, K), np.float32)
    Z3 = np.empty((p, K), np.float32) if third else None
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (log_U_hat, log_V_hat, D_hat, X)]
    _lib().zl_zigap_z(_p(Zi), _p(Zj), _p(Z3) if third else None, *[_p(x) for x in a], n, p, K, int(bool(quirk)))
    return (Zi, Zj, Z3) if third else (Zi, Zj)


def gap_z(log_U_hat, log_V_hat, X):
    """gap.py:67-80 as a sequential float32 loop; returns (Z_hat_i, Z_hat_j)."""
    n, K = log_U_hat.shape; p = log_V_hat.shape[0]
    Zi = np.empty((n, K), np.float32); Zj = np.empty((p, K), np.float32)
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (log_U_hat, log_V_hat, X)]
    _lib().zl_gap_z(_p(Zi), _p(Zj), *[_p(x) for x in a], n, p, K)
    return Zi, Zj


def sparse_z(log_U_hat, log_Vp_hat, S_tilde, S_hat, D_hat, X):
    """sparse_zigap.py:100-116 as a sequential float32 loop; returns (DSZ_hat, DZ_hat, DZ_exp_logsum_hat)."""
    n, K = log_U_hat.shape; p = log_Vp_hat.shape[0]
    DSZ = np.empty((n, K), np.float32); DZ = np.empty((p, K), np.float32); DZl = np.empty((p, K), np.float32)
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (log_U_hat, log_Vp_hat, S_tilde, S_hat, D_hat, X)]
    _lib().zl_sparse_z(_p(DSZ), _p(DZ), _p(DZl), *[_p(x) for x in a], n, p, K)
    return DSZ, DZ, DZl
