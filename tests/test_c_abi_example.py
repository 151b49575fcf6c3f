"""The drop-in boundary from plain C (examples/c_abi_zigap.c): no Python, no torch on the caller's side.
CPU: the example compiles as C99 against include/oriana_b200.h and links against the library.
GPU: it runs a ZIGaP construction + steps through the C ABI and its ELBO trace is monotone."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


def _build(tmp_path):
    from oriana_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / 'c_abi_zigap')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(ROOT, 'include'),
                    '-I', os.path.join(CUDA, 'include'), os.path.join(ROOT, 'examples', 'c_abi_zigap.c'),
                    '-L', libdir, '-l:' + os.path.basename(_lib.LIB_PATH), '-L', os.path.join(CUDA, 'lib64'), '-lcudart', '-lm',
                    '-o', exe], check=True)
    return exe, libdir


def test_c_example_compiles_and_links(tmp_path):
    exe, _ = _build(tmp_path)
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [('4096', '1024', '10', '5'), ('300', '200', '3', '4'), ('2500', '900', '40', '3')])
def test_c_example_runs_on_the_gpu(cuda_lib, tmp_path, shape):
    exe, libdir = _build(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ':' + os.path.join(CUDA, 'lib64') + ':' + os.environ.get('LD_LIBRARY_PATH', ''))
    r = subprocess.run([exe, *shape], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'ELBO monotone: yes' in r.stdout and ('ELBO[%s]' % shape[3]) in r.stdout
