"""GPU: the operator seam exercised INSIDE the reference's own `step()`.

The unmodified reference (installed under baseline/_ref by `__graft_entry__.build()`, see oracle/refshim.py) is
imported, and `ZIGaP.compute_Z_q_expectations` / `GaP.compute_Z_q_expectations` (zigap.py:79-95, gap.py:67-80) are
replaced by the ctypes stubs of INTEGRATION.md section 1, i.e. by `ori_*_compute_Z_q_expectations_host` of
liboriana_b200.so.  The reference's `update_variational_parameters` (zigap.py:97-141) then runs unchanged on host numpy
arrays with the O(n p K) loop on the B200, and is stepped against its un-patched twin (the numba kernel) from the same
constructed state.

Tolerance: the reference accumulates the latent-count sums sequentially in float32; the device sums them in tiles.
Over 5 steps the parameters agree to 2e-4 relative (floor 1e-6 max), D_hat to 2e-5 absolute on the CUDA-core kernels
(small problems); a slab of >= 2^21 entries runs the tcgen05 kernels with TF32 operands: 2e-3 / 2e-4 after 2 steps.
"""
import ctypes
import warnings

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _stubs(lib):
    fp = ctypes.c_void_p

    keep = []

    def ptr(a, out=False):
        if a.dtype != np.float32 or a.ndim != 2:                   # the numba signature f4[:, :] raises TypeError too
            raise TypeError('2-D float32 array expected (zigap.py:79)')
        if not a.flags.c_contiguous:
            # f4[:, :] accepts any layout, and the reference does pass one: X[:] comes out of a DataFrame column-major
            # (cmatrix.py:31-37) and .astype keeps that order (zigap.py:112).  The C ABI takes row-major arrays.
            assert not out, 'outputs are np.empty arrays (zigap.py:102-104): always C-contiguous'
            a = np.ascontiguousarray(a)
            keep.append(a)
        return fp(a.ctypes.data)

    def check(rc):
        if rc:
            buf = ctypes.create_string_buffer(512)
            lib.ori_last_error(buf, 512)
            raise RuntimeError(buf.value.decode())

    calls = {'zigap': 0, 'gap': 0}

    def zigap_z(DZ_hat_i, DZ_hat_j, DZ_exp_logsum_hat, log_U_hat, log_V_hat, D_hat, X):   # zigap.py:80
        n, p = X.shape
        K = log_U_hat.shape[1]
        calls['zigap'] += 1
        check(lib.ori_zigap_compute_Z_q_expectations_host(
            ptr(DZ_hat_i, True), ptr(DZ_hat_j, True), ptr(DZ_exp_logsum_hat, True), ptr(log_U_hat), ptr(log_V_hat),
            ptr(D_hat), ptr(X), n, p, K, 1))            # 1 = keep the D_hat[i, k] indexing of zigap.py:94
        keep.clear()

    def gap_z(Z_hat_i, Z_hat_j, log_U_hat, log_V_hat, X):                                  # gap.py:68
        n, p = X.shape
        K = log_U_hat.shape[1]
        calls['gap'] += 1
        check(lib.ori_gap_compute_Z_q_expectations_host(ptr(Z_hat_i, True), ptr(Z_hat_j, True), ptr(log_U_hat),
                                                        ptr(log_V_hat), ptr(X), n, p, K))
        keep.clear()
    return zigap_z, gap_z, calls


@pytest.mark.parametrize('model_name,shape', [('ZIGaP', (300, 220, 5)), ('GaP', (257, 190, 7)),
                                              ('ZIGaP', (2100, 1100, 12))])     # the last: tensor-path sized slab
def test_reference_step_with_the_b200_operator(cuda_lib, model_name, shape):
    from oracle import refshim, cavi_numpy as cn
    if not refshim.available():
        pytest.skip('the reference install (baseline/_ref) did not travel: run __graft_entry__.build() in the container')
    n, p, K = shape
    X = cn.synth_counts(n, p, K, seed=11)
    refshim.import_reference()
    try:
        import oriana.models as ref_models
        from oriana.singlecell import CountMatrix
        cls = getattr(ref_models, model_name)
        original = cls.__dict__['compute_Z_q_expectations']
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            np.random.seed(3)
            twin = cls(CountMatrix(X.copy()), k=K, use_factors=False)      # un-patched: numba kernel
            np.random.seed(3)
            ours = cls(CountMatrix(X.copy()), k=K, use_factors=False)
            for k in ('a1', 'a2', 'b1', 'b2'):
                assert np.array_equal(getattr(twin, k)[:], getattr(ours, k)[:])
            zigap_z, gap_z, calls = _stubs(cuda_lib)
            steps = 5 if n * p < 1_000_000 else 2
            for _ in range(steps):
                twin.step()
            try:
                cls.compute_Z_q_expectations = staticmethod(zigap_z if model_name == 'ZIGaP' else gap_z)
                for _ in range(steps):
                    ours.step()
            finally:
                cls.compute_Z_q_expectations = original
        assert calls['zigap' if model_name == 'ZIGaP' else 'gap'] == steps
        names = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2') + (('pi_d',) if model_name == 'ZIGaP' else ())
        tensor = n * p >= (1 << 21)          # slabs of this size run the tcgen05 kernels: TF32 operands (measured 3.5e-4)
        for k in names:
            e = relerr(getattr(ours, k)[:], getattr(twin, k)[:])
            assert e < (2e-3 if tensor else 2e-4), (model_name, k, e)
        if model_name == 'ZIGaP':
            assert np.max(np.abs(ours.D_hat - twin.D_hat)) < (2e-4 if tensor else 2e-5)
    finally:
        refshim.release_reference()
