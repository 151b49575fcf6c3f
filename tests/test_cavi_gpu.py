"""GPU: the device CAVI iteration (`model.step()`, base.py:54-56) against

  * the golden trajectories recorded from the UNMODIFIED reference (compat_quirk=True reproduces zigap.py:94),
  * the float32/float64 oracle port for the de-quirked update and for the ELBO (new; no reference value),
  * size-independent properties at a larger size.

Stated tolerances (float32 device arithmetic vs the reference's mixed float32/float64):
  parameters a1,a2,b1,b2,alpha,beta,pi : 2e-5 after 1 step, 2e-4 after 50 steps (relative, floor 1e-6*max)
  D_hat                                : 2e-5 absolute
  ELBO                                 : 1e-5 relative (north_star asks 1e-4)
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_state, load_golden, relerr, relerr_quantile

pytestmark = pytest.mark.gpu
PARAMS = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2')


def make_model(s, quirk, **kw):
    from oriana.models import GaP, ZIGaP
    from oriana.singlecell import CountMatrix
    cls = ZIGaP if ('p_d' in s or 'pi_d' in s) else GaP
    K = s['a1'].shape[1]
    return cls(CountMatrix(s['X']), k=K, use_factors=False, state=s, compat_quirk=quirk, **kw)


def tol(t):
    return 2e-5 if t <= 1 else (6e-5 if t <= 10 else 2e-4)


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_trajectory_matches_reference(cuda_lib, name):
    g = load_golden(name)
    s = golden_state(g, 0)
    m = make_model(s, quirk=True)
    for k in PARAMS:                                   # the hand-off itself
        assert relerr(getattr(m, k).asarray(), s[k]) < 1e-6, k
    if 'pi_d' in s:
        assert relerr(m.pi_d.asarray(), s['pi_d']) < 1e-12
        assert np.array_equal(m.D_hat, (s['X'] > 0).astype(np.float32))
    steps = [int(t) for t in g['steps']]
    for t in range(1, max(steps) + 1):
        m.step()
        if t in steps:
            r = golden_state(g, t)
            for k in PARAMS + (('pi_d',) if 'pi_d' in s else ()):
                e = relerr(getattr(m, k).asarray(), r[k])
                assert e < tol(t), (name, t, k, e)
            assert relerr(m.U_hat, g['s%d_U_hat' % t]) < tol(t)
            assert relerr(m.V_hat, g['s%d_V_hat' % t]) < tol(t)
            assert np.max(np.abs(m.log_U_hat - g['s%d_log_U_hat' % t])) < 50 * tol(t)
            if 'p_d' in s:
                assert np.max(np.abs(m.D_hat - r['p_d'])) < 2e-5, (name, t)


@pytest.mark.parametrize('name', ['zigap_ragged', 'gap_ragged', 'zigap_k10'])
def test_dequirked_update_and_elbo_match_oracle(cuda_lib, name):
    from oracle import cavi_numpy as cn
    g = load_golden(name)
    s = golden_state(g, 0)
    m = make_model(s, quirk=False)
    ref = {k: v.copy() for k, v in s.items()}
    e0 = cn.elbo(ref, guard32=True)     # random Gamma(1) start: fp32 exp underflow + den guard (zigap.py:90)
    assert abs(m.elbo() - e0) <= 1e-6 * abs(e0), (m.elbo(), e0)
    want = [e0]
    for t in range(1, 9):
        m.step()
        cn.step(ref, quirk=False)
        want.append(cn.elbo(ref))
    for k in PARAMS + (('pi_d',) if 'pi_d' in s else ()):
        e = relerr(getattr(m, k).asarray(), ref[k])
        assert e < 6e-5, (k, e)
    got = m.elbo_trace
    assert got.shape == (9,)
    assert np.max(np.abs(got - np.asarray(want)) / np.abs(want)) < 1e-5, (got, want)
    assert np.all(np.diff(got) >= -1e-7 * np.abs(got[:-1]))          # valid CAVI bound: monotone


def test_split_step_equals_step_and_state_roundtrip(cuda_lib):
    g = load_golden('zigap_k10')
    s = golden_state(g, 0)
    a = make_model(s, quirk=False); b = make_model(s, quirk=False)
    for _ in range(3):
        a.step()
        b.update_variational_parameters(); b.update_prior_hyper_parameters()
    for k in PARAMS + ('pi_d',):
        assert relerr(a.state_dict()[k], b.state_dict()[k]) < 5e-6      # float atomics: order differs run to run
    # mid-run snapshot -> new model continues identically (needs pi_prev)
    snap = a.state_dict()
    snap['X'] = s['X']
    c = make_model(snap, quirk=False)
    a.step(); c.step()
    for k in PARAMS + ('pi_d',):
        assert relerr(c.state_dict()[k], a.state_dict()[k]) < 5e-6, k
    bad = dict(golden_state(g, 1))                                  # soft p_d without pi_prev must be refused
    with pytest.raises(ValueError):
        make_model(bad, quirk=False)


def test_state_roundtrip_without_a_step_in_between(cuda_lib):
    """load -> state_dict() -> load with no step in between keeps the generating pi (pi_prev): the checkpoint of a freshly
    resumed model continues like the model it was saved from."""
    g = load_golden('zigap_k10')
    s = golden_state(g, 0)
    a = make_model(s, quirk=False)
    for _ in range(3):
        a.step()
    snap = a.state_dict(); snap['X'] = s['X']
    b = make_model(snap, quirk=False)
    snap2 = b.state_dict(); snap2['X'] = s['X']             # no step between load and save
    assert np.array_equal(snap2['pi_prev'], snap['pi_prev'])
    assert snap2['iterations'] == snap['iterations'] == 3
    c = make_model(snap2, quirk=False)
    a.step(); c.step()
    for k in PARAMS + ('pi_d',):
        assert relerr(c.state_dict()[k], a.state_dict()[k]) < 5e-6, k


def test_host_streamed_resume_from_mid_run_state(cuda_lib):
    """`HostStreamedCAVI` resumed from its own state_dict() (or a device model's) follows the run it was saved from: the
    first step rebuilds D_hat from pi_prev, not from the indicator of a fresh model."""
    import torch
    from oriana_b200.host_step import HostStreamedCAVI
    g = load_golden('zigap_ragged')
    s = golden_state(g, 0)
    K = s['a1'].shape[1]
    Xh = torch.as_tensor(s['X'].astype(np.float32)).pin_memory()
    h = HostStreamedCAVI(Xh, K, s, dropout=True, slab_rows=64)
    for _ in range(3):
        h.step()
    snap = h.state_dict()
    h2 = HostStreamedCAVI(Xh, K, snap, dropout=True, slab_rows=96)
    assert h2.iterations == 3
    e1, e2 = h.step(), h2.step()
    assert abs(e1 - e2) < 1e-6 * abs(e1)
    a, b = h.state_dict(), h2.state_dict()
    for k in PARAMS:
        assert relerr(b[k], a[k]) < 5e-6, k
    m = make_model(s, quirk=False)                           # the device model's snapshot seeds the host-streamed run too
    for _ in range(3):
        m.step()
    h3 = HostStreamedCAVI(Xh, K, m.state_dict(), dropout=True, slab_rows=64)
    e3 = h3.step()                                           # = the ELBO of the state it started from
    assert abs(e3 - m.elbo()) < 2e-6 * abs(e3), (e3, m.elbo())
    c = h3.state_dict()
    for k in PARAMS:
        assert relerr(c[k], a[k]) < 2e-5, k
    bad = dict(snap); bad.pop('pi_prev')
    with pytest.raises(ValueError):
        HostStreamedCAVI(Xh, K, bad, dropout=True)


def test_all_zero_gene_and_cell(cuda_lib):
    """Columns with pi = 0 take the 1e-10 override (zigap.py:133); all-zero rows and genes stay finite."""
    from oracle import cavi_numpy as cn
    X = cn.synth_counts(150, 70, 3, seed=2)
    X[:, 5] = 0; X[:, 64] = 0; X[17, :] = 0
    s = cn.init_state(X, 3, np.random.default_rng(0), 'zigap')
    m = make_model(s, quirk=False)
    ref = {k: v.copy() for k, v in s.items()}
    for _ in range(4):
        m.step(); cn.step(ref, quirk=False)
    for k in PARAMS + ('pi_d',):
        got = getattr(m, k).asarray()
        assert np.isfinite(got).all()
        assert relerr(got, ref[k]) < 6e-5, k
    assert abs(m.pi_d[5] - ref['pi_d'][5]) < 1e-12 and m.pi_d[5] < 1e-9


def test_fresh_init_runs_and_elbo_increases(cuda_lib):
    """`use_factors=False` bootstrap on the device model's own RNG draws (zigap.py:55-77, base.py:43-52)."""
    from oracle import cavi_numpy as cn
    from oriana.models import GaP, ZIGaP
    X = cn.synth_counts(700, 300, 6, seed=9)
    for cls in (ZIGaP, GaP):
        np.random.seed(1)
        m = cls(X, k=6, use_factors=False)
        for _ in range(12):
            m.step()
        tr = m.elbo_trace
        assert np.isfinite(tr).all() and np.all(np.diff(tr) >= -1e-6 * np.abs(tr[:-1])), tr
        s = m.state_dict(); s['X'] = X
        if cls is ZIGaP:
            D = m.D_hat
            assert np.all(D[X != 0] == 1.0) and np.all((D >= 0) & (D <= 1))
            s['p_d'] = np.where(X != 0, 1 - 1e-10, D.astype(np.float64))
        ref_elbo = cn.elbo(s)
        assert abs(tr[-1] - ref_elbo) < 1e-5 * abs(ref_elbo)


def test_full_size_properties_config2(cuda_lib):
    """BASELINE.json configs[1] (10k x 2k, K=10): properties that do not need the oracle at this size."""
    import torch
    from oriana.models import ZIGaP
    from oriana.singlecell import synth_counts_device
    n, p, K = 10_000, 2_000, 10
    X = synth_counts_device(n, p, K, seed=3)
    Xp = X[:, :p]
    zf = float((Xp == 0).float().mean())
    assert 0.35 < zf < 0.75 and float(Xp.max()) < 5000 and float(Xp.mean()) > 1.0
    np.random.seed(0)
    m = ZIGaP(X[:, :p], k=K, use_factors=False)
    for _ in range(6):
        m.step()
    tr = m.elbo_trace
    assert np.isfinite(tr).all() and np.all(np.diff(tr) >= -1e-6 * np.abs(tr[:-1]))
    st = m.device_state()
    # sum_i Zi = sum_j Zj = sum X  => sum(a1) - n*sum(alpha1_prev) bookkeeping: check the identity on one pass
    # pi is a column mean of probabilities that are 1 on non-zeros
    pi = m.pi_d.asarray()
    nzfrac = (Xp != 0).double().mean(0).cpu().numpy()
    assert np.all(pi >= nzfrac - 1e-9) and np.all(pi <= 1 + 1e-12)
    assert torch.isfinite(st['U_hat']).all() and torch.isfinite(st['V_hat']).all()
    # rows of a permuted problem give the permuted answer (the row pass has no cross-row state)
    s0 = m.state_dict(); s0['X'] = X[:, :p]
    perm = torch.randperm(n, device=X.device)
    pn = perm.cpu().numpy()
    s1 = dict(s0); s1['X'] = X[perm][:, :p].contiguous(); s1['a1'] = s0['a1'][pn]; s1['a2'] = s0['a2'][pn]
    ma = ZIGaP(s0['X'], k=K, use_factors=False, state=s0)
    mb = ZIGaP(s1['X'], k=K, use_factors=False, state=s1)
    for _ in range(2):
        ma.step(); mb.step()
    # this size takes the tensor path: the gene sums are accumulated tile by tile in a different order, and
    # their fp32 / TF32-operand rounding feeds the second step (stated tolerance of that path: 1e-3)
    assert ma.uses_tensor_path and mb.uses_tensor_path
    assert relerr(mb.a1.asarray(), ma.a1.asarray()[pn]) < 2e-4
    assert relerr(mb.b1.asarray(), ma.b1.asarray()) < 2e-4 and relerr(mb.pi_d.asarray(), ma.pi_d.asarray()) < 1e-5
    assert abs(ma.elbo() - mb.elbo()) < 1e-6 * abs(ma.elbo())


def test_full_size_properties_config3(cuda_lib):
    """BASELINE.json configs[2] (100k x 20k, K=20, dropout on; 8 GB of counts) at full size on the tensor path:
    monotone ELBO, pi bounds, sum-preservation of the latent counts, and the two-block property -- the gene-side
    sums of the whole matrix equal the sums of its two halves (what row sharding over ranks relies on)."""
    import ctypes
    import torch
    from oriana.models import ZIGaP
    from oriana.singlecell import synth_counts_device
    from oriana_b200 import _lib
    n, p, K = 100_000, 20_000, 20
    X = synth_counts_device(n, p, K, seed=3)
    np.random.seed(0)
    m = ZIGaP(X[:, :p], k=K, use_factors=False)
    assert m.uses_tensor_path
    # one step by hand: the latent counts of every cell / gene add up to its counts (sum_k r_ijk = 1, zigap.py:91-94)
    m._call('ori_zero_accumulators', 1)
    m._call('ori_pass_rows', m._gen)
    Zi = (m._Zi * m._eU[m._gen]).sum(1)
    rows = X[:, :p].sum(1)
    assert float(((Zi - rows).abs() / rows.clamp(min=1)).max()) < 2e-3
    m._call('ori_row_update', m._gen, 1)
    m._call('ori_pass_genes', m._gen)
    whole = m._red32.clone()
    Zj = (whole[0] * m._eV).sum(1)
    cols = X[:, :p].sum(0)
    assert float(((Zj - cols).abs() / cols.clamp(min=1)).max()) < 2e-3
    # the same gene-side sums from two row blocks, each through its own problem description
    halves = torch.zeros_like(whole)
    for r0, r1 in ((0, n // 2 + 64), (n // 2 + 64, n)):
        P = _lib.OriProblem.from_buffer_copy(m._P)
        P.n_rows = r1 - r0
        P.X = m._X[r0:r1].data_ptr()
        for name in ('a1', 'a2', 'Zi', 'a2s'):
            setattr(P, name, getattr(m, '_' + name)[r0:r1].data_ptr())
        for g in (0, 1):
            P.U_hat[g] = m._Uhat[g][r0:r1].data_ptr(); P.eU[g] = m._eU[g][r0:r1].data_ptr()
        P.xrow = m._xrow[r0:r1].data_ptr()
        part = torch.zeros_like(whole)
        scratch64 = torch.zeros_like(m._red64)
        P.red32, P.red64 = part.data_ptr(), scratch64.data_ptr()
        # the block's own local half-step (the tensor path lays its workspace out per problem size)
        m._Zi[r0:r1].zero_(); m._a2s[r0:r1].zero_()
        for call, args in (('ori_pass_rows', (m._gen,)), ('ori_row_update', (m._gen, 1)), ('ori_pass_genes', (m._gen,))):
            _lib.check(getattr(m._lib, call)(ctypes.byref(P), *args, _lib.stream_ptr()))
        halves += part
    scale = whole.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((halves - whole).abs() / scale).max()) < 1e-4        # fp32 sums over 100k cells in a different order
    m._call('ori_gene_update', 1); m._pending_mstep = True
    m.update_prior_hyper_parameters()
    for _ in range(3):
        m.step()
    tr = m.elbo_trace
    assert np.isfinite(tr).all() and np.all(np.diff(tr) >= -1e-6 * np.abs(tr[:-1]))
    pi = m.pi_d.asarray()
    nzfrac = (X[:, :p] != 0).double().mean(0).cpu().numpy()
    assert np.all(pi >= nzfrac - 1e-9) and np.all(pi <= 1 + 1e-12)


@pytest.mark.parametrize('name', ['zigap_ragged', 'gap_ragged'])
def test_host_streamed_step_matches_device_model(cuda_lib, name):
    """The host-buffer entry (what bench.py's e2e times) gives the device model's results, slab by slab."""
    import torch
    from oriana_b200.host_step import HostStreamedCAVI
    g = load_golden(name)
    s = golden_state(g, 0)
    m = make_model(s, quirk=False)
    Xh = torch.as_tensor(s['X'].astype(np.float32)).pin_memory()
    h = HostStreamedCAVI(Xh, s['a1'].shape[1], s, dropout='p_d' in s, slab_rows=64)   # 4 ragged slabs
    elbos = []
    for _ in range(4):
        m.step(); elbos.append(h.step())
    hs = h.state_dict()
    for k in PARAMS:
        assert relerr(hs[k], getattr(m, k).asarray()) < 5e-6, k
    want = m.elbo_trace[:4]
    assert np.max(np.abs(np.asarray(elbos) - want) / np.abs(want)) < 1e-6
    assert h.h2d_bytes > 4 * s['X'].size * 4 and h.d2h_bytes > 0


def test_host_streamed_step_with_compact_counts(cuda_lib):
    """X kept on the host as uint16 / uint8 counts (widened on the device) gives the float32-host results and
    moves half / a quarter of the bytes; the widening kernel is exact on ragged shapes."""
    import ctypes
    import torch
    from oriana_b200.host_step import HostStreamedCAVI
    g = load_golden('zigap_ragged')
    s = golden_state(g, 0)
    K = s['a1'].shape[1]
    X32 = torch.as_tensor(s['X'].astype(np.float32)).pin_memory()
    ref = HostStreamedCAVI(X32, K, s, dropout=True, slab_rows=64)
    for _ in range(3):
        ref.step()
    for dt in (torch.uint16, torch.uint8):
        assert s['X'].max() < (1 << (8 * torch.empty((), dtype=dt).element_size()))
        Xq = torch.as_tensor(s['X'].astype(np.float32)).to(dt).pin_memory()
        h = HostStreamedCAVI(Xq, K, s, dropout=True, slab_rows=64)
        for _ in range(3):
            h.step()
        for k in PARAMS:
            assert relerr(h.state_dict()[k], ref.state_dict()[k]) < 5e-6, (dt, k)
        assert h.h2d_bytes < ref.h2d_bytes
    # saturating uint8 + escapes: counts >= 255 survive
    from oriana_b200.host_step import CompactCounts
    s2 = dict(s); X2 = s['X'].copy(); X2[3, 7] = 255; X2[100, 330] = 70000; X2[192, 0] = 256; s2['X'] = X2
    X2f = torch.as_tensor(X2.astype(np.float32))
    cc = CompactCounts.from_tensor(X2f, chunk_rows=50)
    assert cc.row.numel() == 3 and torch.equal(cc.dense(), X2f)
    ref2 = HostStreamedCAVI(X2f.pin_memory(), K, s2, dropout=True, slab_rows=64)
    h2 = HostStreamedCAVI(cc, K, s2, dropout=True, slab_rows=64)
    for _ in range(2):
        ref2.step(); h2.step()
    for k in PARAMS:
        assert relerr(h2.state_dict()[k], ref2.state_dict()[k]) < 5e-6, ('u8esc', k)
    assert h2.h2d_bytes < 0.5 * ref2.h2d_bytes       # X: 1 byte instead of 4 per entry (+ a1, a2, escapes)
    # the kernel itself, odd sizes and strides
    rng = np.random.default_rng(0)
    for rows, p, lds in ((7, 13, 13), (33, 130, 131), (5, 64, 64)):
        for dt, hi in ((np.uint16, 65535), (np.uint8, 255)):
            a = rng.integers(0, hi + 1, size=(rows, lds)).astype(dt)
            src = torch.as_tensor(a.view(np.int16) if dt is np.uint16 else a).cuda()
            ldd = (p + 3) // 4 * 4
            dst = torch.zeros((rows, ldd), dtype=torch.float32, device='cuda')
            rc = cuda_lib.ori_widen_counts_f32(src.data_ptr(), a.itemsize, lds, dst.data_ptr(), ldd, rows, p,
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0
            assert np.array_equal(dst.cpu().numpy()[:, :p], a[:, :p].astype(np.float32))


def test_host_streamed_step_with_sparse_counts(cuda_lib):
    """X kept on the host as bitmap + non-zero bytes (SparseCounts; cmatrix.py:100-104's sparse view as a streaming
    format): the device expansion is exact (ragged widths, rows without a non-zero, rows longer than one block of bitmap
    words, counts >= 255 through the escape list) and the host-streamed step gives the float32-host results for fewer bytes."""
    import ctypes
    import torch
    from oriana_b200.host_step import HostStreamedCAVI, SparseCounts
    rng = np.random.default_rng(5)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rows, p, keep in ((7, 13, 0.5), (33, 130, 0.3), (5, 64, 1.0), (3, 1, 0.5), (4, 40000, 0.2), (300, 977, 0.05)):
        X = (rng.poisson(9., size=(rows, p)) * (rng.random((rows, p)) < keep)).astype(np.float32)
        X[rows // 2] = 0                                           # a cell without a single count
        X[0, p - 1] = 255; X[rows - 1, 0] = 70000
        sp = SparseCounts.from_tensor(torch.as_tensor(X), chunk_rows=5, pin=False)
        assert np.array_equal(sp.dense().numpy(), X)
        ldd = (p + 3) // 4 * 4
        dst = torch.full((rows, ldd), -1., dtype=torch.float32, device='cuda')
        bm, nz, off = sp.bitmap.cuda(), sp.nz.cuda(), sp.rowoff.cuda()
        # a slab in the middle of the stream: rows [r0, rows), base = position of its first byte
        for r0 in (0, rows // 3):
            lo, hi = sp.byte_range(r0, rows)
            rc = cuda_lib.ori_expand_bitmap_counts_f32(bm[r0:].data_ptr(), bm.shape[1], nz[lo:].data_ptr() if hi > lo else nz.data_ptr(),
                                                       off[r0:].data_ptr(), lo, dst[r0:].data_ptr(), ldd, rows - r0, p, st())
            assert rc == 0
            a = sp.escapes(r0, rows)
            cnt = a[1] - a[0]
            if cnt:
                er, ec, ev = sp.row[a[0]:a[1]].cuda(), sp.col[a[0]:a[1]].cuda(), sp.val[a[0]:a[1]].cuda()
                assert cuda_lib.ori_scatter_counts_f32(dst[r0:].data_ptr(), ldd, r0, rows - r0, p, er.data_ptr(), ec.data_ptr(),
                                                       ev.data_ptr(), cnt, st()) == 0
            assert np.array_equal(dst.cpu().numpy()[r0:, :p], X[r0:]), (rows, p, r0)
    g = load_golden('zigap_ragged')
    s = dict(golden_state(g, 0))
    X2 = s['X'].copy(); X2[3, 7] = 255; X2[100, 330] = 70000; X2[64] = 0; s['X'] = X2
    K = s['a1'].shape[1]
    X2f = torch.as_tensor(X2.astype(np.float32))
    sp = SparseCounts.from_tensor(X2f, chunk_rows=50)
    assert sp.row.numel() == 2 and torch.equal(sp.dense(), X2f)
    ref = HostStreamedCAVI(X2f.pin_memory(), K, s, dropout=True, slab_rows=64)
    h = HostStreamedCAVI(sp, K, s, dropout=True, slab_rows=64)
    for _ in range(3):
        e_ref, e = ref.step(), h.step()
    for k in PARAMS:
        assert relerr(h.state_dict()[k], ref.state_dict()[k]) < 5e-6, ('sparse', k)
    assert abs(e - e_ref) < 1e-6 * abs(e_ref)
    zeros = float((X2 == 0).mean())
    assert h.h2d_bytes < ref.h2d_bytes * (0.25 * (0.125 + 1 - zeros) + 0.1)


def test_count_matrix_to_device_narrow_upload(cuda_lib):
    """CountMatrix.to_device: uint8 / uint16 / float32 uploads give the same float32 matrix in HBM, with the 16-byte
    row pitch the kernels need (ragged p), uploaded in several slabs."""
    import torch
    from oriana.singlecell import CountMatrix
    rng = np.random.default_rng(3)
    X = rng.poisson(2., size=(301, 77)).astype(np.int64)
    for top, dt in ((None, torch.uint8), (300, torch.uint16), (70000, None)):
        Y = X.copy()
        if top:
            Y[17, 5] = top
        c = CountMatrix(Y)
        assert c.narrow_dtype() == dt
        d = c.to_device('cuda', slab_rows=64)
        assert d.shape == (301, 77) and d.dtype == torch.float32 and d.stride(0) == 80 and d.data_ptr() % 16 == 0
        np.testing.assert_array_equal(d.cpu().numpy(), Y.astype(np.float32))
    Z = torch.as_tensor(X, device='cuda')
    np.testing.assert_array_equal(CountMatrix(Z).to_device().cpu().numpy(), X.astype(np.float32))


def test_nmf_warm_start_on_device(cuda_lib):
    """`use_factors=True` (base.py:38-40) with the factorisation done in HBM: the factors are non-negative, reconstruct
    X about as well as sklearn's NMF does, seed a1 / b1 exactly like the host path, and the CAVI steps run from them."""
    from oracle import cavi_numpy as cn
    from oriana.models import ZIGaP
    from oriana.singlecell import CountMatrix
    from sklearn.decomposition import NMF
    X = cn.synth_counts(400, 300, 5, seed=9)
    np.random.seed(0)
    m = ZIGaP(CountMatrix(X), k=5, use_factors=True, nmf='device')
    U, V = m._nmf_U.cpu().numpy().astype(np.float64), m._nmf_V.cpu().numpy().astype(np.float64)
    assert (U >= 0).all() and (V >= 0).all()
    err_dev = np.linalg.norm(X - U @ V.T)
    sk = NMF(n_components=5, max_iter=400)
    W = sk.fit_transform(X.astype(np.float64))
    err_host = np.linalg.norm(X - W @ sk.components_)
    assert err_dev < 1.05 * err_host, (err_dev, err_host)
    assert relerr(m.a1.asarray(), np.maximum(U, 1e-15)) < 1e-6 and relerr(m.b1.asarray(), np.maximum(V, 1e-15)) < 1e-6
    for _ in range(3):
        m.step()
    tr = m.elbo_trace
    assert np.isfinite(tr).all() and (np.diff(tr) > -1e-6 * np.abs(tr[:-1])).all()
    np.random.seed(0)
    mh = ZIGaP(CountMatrix(X), k=5, use_factors=True, nmf='host')      # the reference's own route (sklearn)
    mh.step()
    assert np.isfinite(mh.elbo_trace).all()


@pytest.mark.parametrize('name', ['zigap_nmf', 'gap_nmf'])
@pytest.mark.parametrize('tensor', [False, True])
def test_nmf_initialised_trajectory_matches_reference(cuda_lib, tensor, name):
    """The reference's default construction path (`use_factors=True`, base.py:38-40) from its recorded
    post-construction state (tests/golden/zigap_nmf.npz): E[log U], E[log V] between -100 and -1e15, where
    exp(lU) * exp(lV) leaves float32 although the reference's exp(lU + lV) does not.  Both kernel families stay finite
    and on the reference's trajectory (per-row centred exponentials, csrc/special.cuh)."""
    from oriana.models import GaP, ZIGaP
    from oriana.singlecell import CountMatrix
    g = load_golden(name)
    s = golden_state(g, 0)
    cls = ZIGaP if 'p_d' in s else GaP
    m = cls(CountMatrix(s['X']), k=s['a1'].shape[1], use_factors=False, state=s, compat_quirk=True, tensor=tensor)
    assert m.uses_tensor_path == tensor
    steps = [int(t) for t in g['steps']]
    # tensor path: TF32 operand rounding is amplified by the near-dead components of this regime
    ftol, htol, dtol = (1e-2, 3e-4, 1e-3) if tensor else (3e-4, 1e-5, 2e-5)
    for t in range(1, max(steps) + 1):
        m.step()
        if t in steps:
            r = golden_state(g, t)
            for k in ('a1', 'a2', 'b1', 'b2'):
                got = getattr(m, k).asarray()
                assert np.isfinite(got).all(), (t, k)
                q, worst = relerr_quantile(got, r[k])      # all but the entries fed by denormal-range terms (conftest.py)
                assert q < ftol and worst < 0.5, (t, k, q, worst)
            for k in ('alpha1', 'alpha2', 'beta1', 'beta2') + (('pi_d',) if 'p_d' in s else ()):
                assert relerr(getattr(m, k).asarray(), r[k]) < htol, (t, k)
            if 'p_d' in s:
                assert np.max(np.abs(m.D_hat - r['p_d'])) < dtol, t
    assert np.isfinite(m.elbo_trace).all()


@pytest.mark.parametrize('case', ['zigap_simt', 'gap_simt', 'zigap_tensor', 'sparse'])
def test_graph_replayed_steps_match_eager_steps(cuda_lib, case):
    """`graphs=True`: step() replays a captured CUDA graph (one per generation parity, iteration count on the device).
    Same states and the same ELBO trace as separate launches, up to the order of the float atomics."""
    from oracle import cavi_numpy as cn, sparse_numpy as sn
    from oriana.models import GaP, SparseZIGaP, ZIGaP
    from oriana.singlecell import CountMatrix
    if case == 'sparse':
        X = cn.synth_counts(300, 260, 5, seed=4)
        s = sn.init_state(X, 5, np.random.default_rng(1))
        mk = lambda **kw: SparseZIGaP(CountMatrix(X), k=5, use_factors=False, state={k: np.array(v) for k, v in s.items()}, **kw)
        keys = ('a1', 'a2', 'b1', 'b2', 'p_s', 'pi_s', 'pi_d', 'alpha1', 'beta2')
    else:
        n, p, K = (4000, 1800, 12) if case == 'zigap_tensor' else (300, 260, 5)
        X = cn.synth_counts(n, p, K, seed=4)
        kind = 'gap' if case.startswith('gap') else 'zigap'
        s = cn.init_state(X, K, np.random.default_rng(1), kind)
        cls = GaP if kind == 'gap' else ZIGaP
        mk = lambda **kw: cls(CountMatrix(X), k=K, use_factors=False, state={k: np.array(v) for k, v in s.items()},
                              tensor=(case == 'zigap_tensor'), **kw)
        keys = ('a1', 'a2', 'b1', 'b2', 'alpha1', 'beta2') + (('pi_d',) if kind == 'zigap' else ())
    me, mg = mk(), mk(graphs=True)
    for _ in range(7):
        me.step(); mg.step()
    assert mg.graph_replays == 5 and mg.graph_kernel_launches > 0 and mg.iterations == me.iterations == 7
    tol = 2e-3 if case in ('sparse', 'zigap_tensor') else 1e-5     # TF32 operands: a re-rounded term moves a sum by up to 2^-12
    for k in keys:
        assert relerr(getattr(mg, k).asarray(), getattr(me, k).asarray()) < tol, (case, k)
    if case != 'sparse':
        te, tg = me.elbo_trace, mg.elbo_trace
        assert te.shape == tg.shape == (8,) and np.max(np.abs(te - tg) / np.abs(te)) < 1e-5
    else:
        assert abs(mg.reconstruction_deviance() - me.reconstruction_deviance()) <= 1e-5 * abs(me.reconstruction_deviance())
    # an edit of the parameters falls back to eager launches for that step and the graphs stay valid afterwards
    mg.a1[:] = mg.a1.asarray(); me.a1[:] = me.a1.asarray()
    for _ in range(2):
        me.step(); mg.step()
    assert relerr(mg.b1.asarray(), me.b1.asarray()) < tol


def test_block_generator_on_device(cuda_lib):
    """The reference's block generator (generation.py:8-86) drawn and multiplied in HBM: same block structure, labels and
    distributions as the host version (which is pinned bit for bit to the reference, tests/test_host_logic.py) -- other random
    stream, so the comparison is statistical; row blocks generated per rank tile the same genes."""
    from oriana.singlecell import generate_factor_matrices, generate_factor_matrices_device
    from oriana.models import SparseZIGaP
    n, m, k = 900, 700, 8
    np.random.seed(4)
    Xh, Uh, Vh, lh = generate_factor_matrices(n, m, k, sparsity_degree_in_v=0.5, n_groups=2, zero_inflation_level=0.4)
    X, U, V, lab = generate_factor_matrices_device(n, m, k, sparsity_degree_in_v=0.5, n_groups=2, zero_inflation_level=0.4, seed=4)
    assert X.is_cuda and X.shape == (n, m) and U.shape == (n, k) and V.shape == (m, k)
    assert np.array_equal(lab.cpu().numpy(), lh)
    Xd, Ud, Vd = X.cpu().numpy(), U.cpu().numpy(), V.cpu().numpy()
    assert (Xd >= 0).all() and np.array_equal(Xd, np.floor(Xd))
    assert abs((Xd == 0).mean() - (Xh == 0).mean()) < 0.08                   # zero inflation level (pi_j is random per gene)
    m0 = int(round(m * 0.5))
    for A, B in ((Vd, Vh),):                                                 # on-block scale beta, background (1 - theta) beta
        assert abs(A[:m0 // 2, :k // 2].mean() / B[:m0 // 2, :k // 2].mean() - 1) < 0.1
        assert abs(A[m0:, :].mean() / B[m0:, :].mean() - 1) < 0.1
    # per-rank generation: same genes, own cells
    Xa, Ua, Va, la = generate_factor_matrices_device(n, m, k, sparsity_degree_in_v=0.5, n_groups=2, zero_inflation_level=0.4,
                                                     seed=4, row0=0, row1=450)
    Xb, Ub, Vb, lb = generate_factor_matrices_device(n, m, k, sparsity_degree_in_v=0.5, n_groups=2, zero_inflation_level=0.4,
                                                     seed=4, row0=450, row1=n)
    assert torch_equal(Va, V) and torch_equal(Vb, V) and Xa.shape == (450, m) and Xb.shape == (450, m)
    assert np.array_equal(np.concatenate([la.cpu().numpy(), lb.cpu().numpy()]), lh)
    # and the model the reference's driver runs on such data steps on it (experiments/clustering.py:18-38)
    mdl = SparseZIGaP(X, k=k, use_factors=False)
    d0 = mdl.reconstruction_deviance()
    for _ in range(3):
        mdl.step()
    assert np.isfinite(mdl.a1.asarray()).all() and mdl.reconstruction_deviance() < d0


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))


@pytest.mark.parametrize('name', ['gap_c1', 'gap_ragged'])
def test_model_loglikelihood_matches_reference(cuda_lib, name):
    """FactorModel.loglikelihood() (base.py:89-95) against the values the reference's GaP printed along the recorded
    trajectories (tests/golden/loglik.npz, oracle/make_golden.py loglik); ZIGaP has none in the reference either."""
    import os
    from conftest import GOLDEN
    ll = np.load(os.path.join(GOLDEN, 'loglik.npz'))
    g = load_golden(name)
    s = golden_state(g, 0)
    m = make_model(s, quirk=True)
    steps = [int(t) for t in g['steps']]
    for t in range(1, max(steps) + 1):
        m.step()
        if t in steps:
            want = float(ll['%s_s%d' % (name, t)])
            got = m.loglikelihood()
            assert abs(got - want) < 2e-4 * abs(want) + 1.0, (name, t, got, want)
    assert bool(ll['zigap_raises'])
    z = make_model(golden_state(load_golden('zigap_ragged'), 0), quirk=True)
    with pytest.raises(AttributeError):
        z.loglikelihood()


def test_node_logp_matches_reference_including_its_broadcasting(cuda_lib):
    """Gamma / Bernoulli / Poisson `logp()` against arrays the reference's own nodes returned (tests/golden/loglik.npz):
    same formulas (gamma.py:63-68 uses beta as a scale; poisson.py:64-73 has no log-factorial term and the -1000 rule)
    and the same result shapes ((n, m, m) from the flat parameter broadcast; flat for Poisson)."""
    import os
    from conftest import GOLDEN
    from oriana import Dimensions, Parameter
    from oriana.nodes import Bernoulli, Gamma, Poisson
    f = np.load(os.path.join(GOLDEN, 'loglik.npz'))
    dims = Dimensions({'n': 6, 'k': 3, 'm': 4})
    g = Gamma(Parameter(f['node_gamma_alpha']), Parameter(f['node_gamma_beta']), dims('n,k ~ s,d'))
    g.buffer = f['node_gamma_samples']
    got = g.logp()
    assert got.shape == f['node_gamma_logp'].shape == (6, 3, 3)
    assert np.allclose(got, f['node_gamma_logp'], rtol=1e-12, atol=1e-12)
    assert abs(g.loglikelihood() - f['node_gamma_logp'].sum()) < 1e-9
    b = Bernoulli(Parameter(f['node_bern_pi']), dims('n,m ~ s,d'))
    b.buffer = f['node_bern_samples']
    assert b.logp().shape == (6, 4, 4) and np.allclose(b.logp(), f['node_bern_logp'], rtol=1e-12, atol=1e-12)
    po = Poisson(Parameter(f['node_pois_lambda']), dims('n,m ~ d,d'))
    po.buffer = f['node_pois_samples']
    got = po.logp()
    assert got.shape == (24,) and np.allclose(got, f['node_pois_logp'], rtol=1e-12, atol=1e-12)
    assert got[0] == 0.0 and got[1 * 4 + 2] == -1000.0
