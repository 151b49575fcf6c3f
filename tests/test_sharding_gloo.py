"""CPU, world_size 2 over gloo: the row-sharding protocol of the CAVI iteration (oriana_b200/sharding.py).

Each rank owns a contiguous block of cells (X, a1, a2, p_d rows) and the replicated gene side.  Per
iteration exactly two sum-allreduces are issued, the same two buffers the device path reduces over NCCL
(models/base.py: `_red32` = [Zj | D_hat^T U_hat], `_red64` = [colsum p_d | sum_i log U_hat | sum_i U_hat]).
The sharded iteration, written here with the oracle's primitives and `RowSharding.allreduce_sum`, must
reproduce the unsharded oracle step: that pins WHAT is reduced and that nothing else crosses ranks.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def sharded_step(s, shard, n_total, quirk):
    """One CAVI iteration on this rank's row block `s` (reference order zigap.py:97-158)."""
    from oracle import cavi_numpy as cn
    zig = 'p_d' in s
    e = cn.expectations(s)
    Zi, Zj = cn.z_expectations(e['log_U_hat'], e['log_V_hat'], s['X'], e.get('D_hat'), quirk)
    s['a1'] = cn.clamp(s['alpha1'][None, :] + Zi)                                    # row side: local
    if zig:
        s['a2'] = cn.clamp(s['alpha2'] + e['D_hat'] @ e['V_hat'])
    else:
        s['a2'] = cn.clamp(np.broadcast_to(s['alpha2'] + e['V_hat'].sum(axis=0), s['a1'].shape).copy())
    U_hat = cn.gamma_mean(s['a1'], s['a2'])
    b2s = (e['D_hat'].T @ U_hat) if zig else np.broadcast_to(U_hat.sum(axis=0), Zj.shape)
    red32 = torch.as_tensor(np.stack([Zj.astype(np.float64), np.asarray(b2s, dtype=np.float64)]))
    shard.allreduce_sum(red32)                                                       # allreduce #1
    Zj, b2s = red32[0].numpy(), red32[1].numpy()
    s['b1'] = cn.clamp(s['beta1'][None, :] + Zj)
    s['b2'] = cn.clamp(s['beta2'] + b2s)
    V_hat = cn.gamma_mean(s['b1'], s['b2'])
    p = s['b1'].shape[0]
    if zig:
        pi = s['pi_d']
        p_d = cn.sigmoid(cn.logit(pi)[None, :] - U_hat @ V_hat.T)
        p_d[:, pi <= 0] = 1e-10
        p_d[:, pi >= 1] = 1. - 1e-10
        p_d[s['X'] != 0] = 1. - 1e-10
        s['p_d'] = p_d
    e = cn.expectations(s)
    red64 = torch.as_tensor(np.concatenate([s['p_d'].sum(axis=0) if zig else np.zeros(p),
                                            e['log_U_hat'].astype(np.float64).sum(axis=0), e['U_hat'].sum(axis=0)]))
    shard.allreduce_sum(red64)                                                       # allreduce #2
    r = red64.numpy()
    K = s['a1'].shape[1]
    s['alpha1'] = cn.clamp(cn.inverse_digamma(np.log(s['alpha2']) + r[p:p + K] / n_total))
    s['alpha2'] = cn.clamp(s['alpha1'] / (r[p + K:p + 2 * K] / n_total))
    s['beta1'] = cn.clamp(cn.inverse_digamma(np.log(s['beta2']) + np.mean(e['log_V_hat'], axis=0)))
    s['beta2'] = cn.clamp(s['beta1'] / np.mean(e['V_hat'], axis=0))
    if zig:
        s['pi_d'] = r[:p] / n_total
    return s


def sharded_sparse_step(s, shard, n_total, tau):
    """One SparseZIGaP iteration on this rank's row block (reference order sparse_zigap.py:118-196): the device path
    reduces red32 = [R^T eU | D_hat^T U_hat | R^T (eU log U)] and red64 = [colsum p_d | sum log U_hat | sum U_hat];
    the S update, pi_s and the V' side are recomputed on every rank from the reduced sums."""
    from oracle import cavi_numpy as cn, sparse_numpy as sn
    dt = np.float32
    e = sn.expectations(s)
    X = s['X'].astype(dt); D = e['D_hat'].astype(dt); S_hat = e['S_hat'].astype(dt)
    lU = e['log_U_hat'].astype(dt); lV = e['log_Vprime_hat'].astype(dt)
    eU = cn.centred_exp(lU, dt)
    eV = cn.centred_exp(lV, dt) * (s['p_s'] > tau).astype(dt)
    den = eU @ eV.T
    R = X * D / np.where(den > 0, den, dt(1))
    V_eff = S_hat * e['Vprime_hat']
    s['a1'] = cn.clamp(s['alpha1'][None, :] + (R @ (eV * S_hat)) * eU)              # row side: local
    s['a2'] = cn.clamp(s['alpha2'] + e['D_hat'] @ V_eff)
    U_hat = cn.gamma_mean(s['a1'], s['a2'])
    red32 = torch.as_tensor(np.stack([(R.T @ eU).astype(np.float64), np.asarray(e['D_hat'].T @ U_hat, dtype=np.float64),
                                      (R.T @ (eU * lU)).astype(np.float64)]))
    shard.allreduce_sum(red32)                                                       # allreduce #1 (three blocks)
    RtU, DtU, Zl = (red32[i].numpy() for i in range(3))
    DZ = (RtU.astype(dt) * eV); DZl = ((Zl.astype(dt) + lV * RtU.astype(dt)) * eV)
    s['b1'] = cn.clamp(s['beta1'][None, :] + S_hat * DZ)
    s['b2'] = cn.clamp(s['beta2'] + S_hat * DtU)
    Vp = cn.gamma_mean(s['b1'], s['b2'])
    pi_s = s['pi_s']
    p_s = np.nan_to_num(cn.sigmoid(cn.logit(pi_s)[:, None] - (-DZl + np.nan_to_num(DtU * Vp))))
    p_s[pi_s <= 0] = 1e-10; p_s[pi_s >= 1] = 1. - 1e-10
    s['p_s'] = p_s
    pi = s['pi_d']
    p_d = cn.sigmoid(cn.logit(pi)[None, :] - U_hat @ V_eff.T)                        # OLD effective V_hat, NEW U_hat
    p_d[:, pi <= 0] = 1e-10; p_d[:, pi >= 1] = 1. - 1e-10; p_d[s['X'] != 0] = 1. - 1e-10
    s['p_d'] = p_d
    e = sn.expectations(s)
    p, K = s['b1'].shape
    red64 = torch.as_tensor(np.concatenate([p_d.sum(axis=0), e['log_U_hat'].astype(np.float64).sum(axis=0),
                                            e['U_hat'].sum(axis=0)]))
    shard.allreduce_sum(red64)                                                       # allreduce #2
    r = red64.numpy()
    s['alpha1'] = cn.clamp(cn.inverse_digamma(np.log(s['alpha2']) + r[p:p + K] / n_total))
    s['alpha2'] = cn.clamp(s['alpha1'] / (r[p + K:p + 2 * K] / n_total))
    s['beta1'] = cn.clamp(cn.inverse_digamma(np.log(s['beta2']) + np.mean(e['log_Vprime_hat'], axis=0)))
    s['beta2'] = cn.clamp(s['beta1'] / np.mean(e['Vprime_hat'], axis=0))
    s['pi_d'] = r[:p] / n_total
    s['pi_s'] = np.mean(s['p_s'], axis=1)
    return s


def _sparse_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import warnings
        from oracle import cavi_numpy as cn, sparse_numpy as sn
        from oriana_b200.sharding import RowSharding
        n, p, K = 61, 40, 3
        X = cn.synth_counts(n, p, K, seed=11)
        full = sn.init_state(X, K, np.random.default_rng(5))
        shard = RowSharding(enabled=True)
        r0, r1 = RowSharding.row_block(n, rank, world)
        mine = {k: (v[r0:r1].copy() if k in ('X', 'a1', 'a2', 'p_d') else v.copy()) for k, v in full.items()}
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for _ in range(3):
                sn.step(full, tau=0.5)
                sharded_sparse_step(mine, shard, n, 0.5)
        err = {}
        for k in ('a1', 'a2'):
            err[k] = float(np.max(np.abs(mine[k] - full[k][r0:r1]) / (np.abs(full[k][r0:r1]) + 1e-9)))
        for k in ('b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2', 'pi_d', 'pi_s'):
            err[k] = float(np.max(np.abs(mine[k] - full[k]) / (np.abs(full[k]) + 1e-9)))
        err['p_s'] = float(np.max(np.abs(mine['p_s'] - full['p_s']) / (np.abs(full['p_s']) + 1e-6)))
        out[rank] = err
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_of_the_sparse_model():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sparse_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert sorted(out.keys()) == [0, 1]
        for rank in (0, 1):
            for k, e in out[rank].items():
                assert e < (2e-3 if k in ('p_s', 'pi_s') else 2e-5), (rank, k, e)   # the S-step amplifies summation order


def _worker(rank, world, port, model, quirk, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import cavi_numpy as cn
        from oriana_b200.sharding import RowSharding
        n, p, K = 61, 40, 3                                   # 61 cells: blocks of 31 and 30
        X = cn.synth_counts(n, p, K, seed=11)
        full = cn.init_state(X, K, np.random.default_rng(5), model)
        shard = RowSharding(enabled=True)
        assert shard.world == world and shard.rank == rank and shard.enabled
        r0, r1 = RowSharding.row_block(n, rank, world)
        assert shard.total_rows(r1 - r0) == n
        t = torch.full((3,), float(rank))
        assert shard.broadcast(t).tolist() == [0., 0., 0.]          # replicated state: rank 0's values win
        mine = {k: (v[r0:r1].copy() if k in ('X', 'a1', 'a2', 'p_d') else v.copy()) for k, v in full.items()}
        for _ in range(4):
            cn.step(full, quirk=quirk)
            sharded_step(mine, shard, n, quirk)
        err = {}
        for k in ('a1', 'a2'):
            err[k] = float(np.max(np.abs(mine[k] - full[k][r0:r1]) / (np.abs(full[k][r0:r1]) + 1e-9)))
        for k in ('b1', 'b2', 'alpha1', 'alpha2', 'beta1', 'beta2') + (('pi_d',) if model == 'zigap' else ()):
            err[k] = float(np.max(np.abs(mine[k] - full[k]) / (np.abs(full[k]) + 1e-9)))
        out[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('model,quirk', [('zigap', False), ('zigap', True), ('gap', False)])
def test_two_rank_row_sharding_reproduces_the_unsharded_step(model, quirk):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), model, quirk, out), nprocs=world, join=True)
        assert sorted(out.keys()) == [0, 1]
        for rank in (0, 1):
            for k, e in out[rank].items():
                assert e < 5e-6, (rank, k, e)      # float32 Z sums are added in a different order


def test_row_blocks_partition_cells():
    from oriana_b200.sharding import RowSharding
    for n, w in ((1_000_000, 8), (61, 2), (7, 8), (100_000, 3)):
        blocks = [RowSharding.row_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
