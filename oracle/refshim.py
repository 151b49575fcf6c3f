"""Import the UNMODIFIED reference in the build container (TEST INFRASTRUCTURE).

`/root/reference` exists only in the build container, never on the GPU box: nothing that runs
under `pytest -m gpu`, `smoke()` or `bench.py` may import this module.  It is used by
`oracle/make_golden.py` to produce the fixtures under `tests/golden/` and by the optional
container-only test `tests/test_oracle.py::test_port_matches_live_reference`.

The reference pins numpy 1.13 and uses `np.float` / `np.int` (`oriana/parameters.py:11`), removed
in numpy >= 1.24; the two-line alias below is the only change needed.
"""
import os, sys
import numpy as np

REFERENCE_ROOT = os.environ.get('ORIANA_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'oriana'))


def import_reference():
    if not available():
        raise ImportError('reference checkout not present at %s' % REFERENCE_ROOT)
    np.float = float  # noqa: alias shim
    np.int = int
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own repo ships an `oriana` alias package: make sure the reference's wins here
    for k in [k for k in sys.modules if k == 'oriana' or k.startswith('oriana.')]:
        del sys.modules[k]
    import oriana  # noqa
    assert os.path.realpath(oriana.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), oriana.__file__
    return oriana


def release_reference():
    """Undo import_reference(): drop the reference modules and its sys.path entry, so that the repo's own
    `oriana` alias package is importable again in the same process."""
    for k in [k for k in sys.modules if k == 'oriana' or k.startswith('oriana.')]:
        del sys.modules[k]
    while REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)


STATE_KEYS_ZIGAP = ('a1', 'a2', 'b1', 'b2', 'p_d', 'pi_d', 'alpha1', 'alpha2', 'beta1', 'beta2')
STATE_KEYS_GAP = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2')
STATE_KEYS_SPARSE = STATE_KEYS_ZIGAP + ('p_s', 'pi_s')


def snapshot(model):
    """Copy the state vector out of a constructed reference model."""
    keys = STATE_KEYS_SPARSE if hasattr(model, 'p_s') else (STATE_KEYS_ZIGAP if hasattr(model, 'p_d') else STATE_KEYS_GAP)
    s = {k: np.array(getattr(model, k)[:], dtype=np.float64, copy=True) for k in keys}
    s['X'] = np.array(model.X[:], copy=True)
    return s
