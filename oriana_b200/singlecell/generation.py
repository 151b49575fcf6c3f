"""The reference's block-structured synthetic data (oriana/singlecell/generation.py:8-86), used by its drivers
(`main.py`, `experiments/clustering.py:47`).  Host numpy, global `np.random` stream consumed in the reference's order,
so a seeded call reproduces the reference's matrices bit for bit (tests/golden/generator.npz, recorded from the
reference; tests/test_host_logic.py).  Bench inputs do NOT come from here (these counts are D * floor(U V^T)
with rates in the thousands, not Poisson draws); see `synth_counts_device`.
"""
import numpy as np


def _block_edges(total, n_groups):
    """Start offsets of `n_groups` equal blocks plus the end: the last block absorbs the remainder (:9-12).  Fewer
    items than groups is an error in the reference too (`range()` with a zero step)."""
    width = total // n_groups
    if width == 0:
        raise ValueError('cannot split %d items into %d groups' % (total, n_groups))
    return [g * width for g in range(n_groups)] + [total]


def _blocked_gamma(rows, cols, row_edges, col_edges, scales, background_scale):
    """Gamma(1, scale_g) on the g-th diagonal block, Gamma(1, background_scale) elsewhere.  The background is drawn as
    one full matrix AFTER the blocks and only its off-block entries are kept (generation.py:33-36, :63-65)."""
    out = np.empty((rows, cols), dtype=np.float64)
    on_block = np.zeros((rows, cols), dtype=bool)
    for g, scale in enumerate(scales):
        r0, r1, c0, c1 = row_edges[g], row_edges[g + 1], col_edges[g], col_edges[g + 1]
        out[r0:r1, c0:c1] = np.random.gamma(1., scale, size=(r1 - r0, c1 - c0))
        on_block[r0:r1, c0:c1] = True
    fill = np.random.gamma(1., background_scale, size=(rows, cols))
    out[~on_block] = fill[~on_block]
    return out


def generate_u(n, k, n_groups=3, theta=0.5):
    """Cell factors with `n_groups` groups of cells, each loading on its own block of components (:8-37).
    Returns (U [n, k], labels [n])."""
    row_edges, col_edges = _block_edges(n, n_groups), _block_edges(k, n_groups)
    alpha = np.random.choice([100, 250], size=n_groups) / k
    labels = np.empty(n, dtype=np.int64)
    for g in range(n_groups):
        labels[row_edges[g]:row_edges[g + 1]] = g
    U = _blocked_gamma(n, k, row_edges, col_edges, alpha, (1. - theta) * np.mean(alpha))
    return U, labels


def generate_v(m, k, sparsity_degree=0.2, beta=80, theta=0.8, n_groups=2):
    """Gene factors: the first round(m * sparsity_degree) genes carry the block structure (:40-66)."""
    m0 = int(np.round(m * sparsity_degree))
    row_edges, col_edges = _block_edges(m0, n_groups), _block_edges(k, n_groups)
    return _blocked_gamma(m, k, row_edges, col_edges, [beta] * n_groups, (1. - theta) * beta)


def generate_factor_matrices(n, m, k, sparsity_degree_in_v=0.5, beta=80, theta=0.8, n_groups=2,
                             zero_inflation_level=0.5):
    """(X, U, V, labels) with X = D * floor(U V^T), D_ij ~ Bernoulli(pi_j), pi_j ~ Beta(1, 1/z - 1) (:68-86)."""
    U, labels = generate_u(n, k, n_groups=n_groups, theta=theta)
    V = generate_v(m, k, sparsity_degree=sparsity_degree_in_v, beta=beta, theta=theta, n_groups=n_groups)
    rate = U @ V.T
    pi_d = np.random.beta(1., (1. / zero_inflation_level) - 1., size=m)
    D = np.random.binomial(np.ones(m, dtype=np.int64), pi_d, size=(n, m))
    return (D * rate).astype(np.int64), U, V, labels


# ---------------------------------------------------------------------------------------------------------------------
# The same generator with the matrices drawn and multiplied in HBM: nothing O(n m) ever
# exists on the host, rows can be generated per rank.  Same block structure and distributions as above; the random
# stream is torch's device generator, so the VALUES differ from the numpy path (which is the one pinned bit for bit to
# the reference) -- tests compare the two statistically.
def _blocked_gamma_device(rows, cols, row_edges, col_edges, scales, background_scale, gen, dev, row0=0, row1=None):
    import torch
    row1 = rows if row1 is None else row1
    shape_one = torch.ones((), device=dev, dtype=torch.float64)

    def gamma(shape, scale):
        # Gamma(1, scale) = scale * Exponential(1): inverse-cdf draws from the seeded device generator
        u = torch.rand(shape, generator=gen, device=dev, dtype=torch.float64)
        return -torch.log1p(-u) * (float(scale) * shape_one)
    out = gamma((row1 - row0, cols), background_scale)
    for g, scale in enumerate(scales):
        r0, r1, c0, c1 = max(row_edges[g], row0), min(row_edges[g + 1], row1), col_edges[g], col_edges[g + 1]
        if r1 > r0:
            out[r0 - row0:r1 - row0, c0:c1] = gamma((r1 - r0, c1 - c0), scale)
    return out


def generate_factor_matrices_device(n, m, k, sparsity_degree_in_v=0.5, beta=80, theta=0.8, n_groups=2,
                                    zero_inflation_level=0.5, seed=0, device=None, row0=0, row1=None):
    """(X, U, V, labels) as CUDA tensors: X = D * floor(U V^T) float32 [rows, m] for the cells [row0, row1) (all by default),
    U float64 [rows, k], V float64 [m, k], labels int64 [rows].  The gene-side draws (V, pi_d) depend on `seed` only, the
    cell-side draws on (seed, row0): every rank of a sharded run generates its own row block against the same genes."""
    import torch
    from .. import _lib
    dev = device or _lib.require_cuda()
    row1 = n if row1 is None else row1
    ggen = torch.Generator(device=dev); ggen.manual_seed(int(seed))
    cgen = torch.Generator(device=dev); cgen.manual_seed(int(seed) * 1_000_003 + 17 + int(row0))
    # generate_u (:8-37)
    row_edges, col_edges = _block_edges(n, n_groups), _block_edges(k, n_groups)
    pick = torch.randint(0, 2, (n_groups,), generator=ggen, device=dev)
    alpha = (torch.where(pick == 0, 100., 250.) / k).double()
    labels = torch.empty((row1 - row0,), dtype=torch.int64, device=dev)
    for g in range(n_groups):
        lo, hi = max(row_edges[g], row0), min(row_edges[g + 1], row1)
        if hi > lo:
            labels[lo - row0:hi - row0] = g
    U = _blocked_gamma_device(n, k, row_edges, col_edges, alpha.tolist(), (1. - theta) * float(alpha.mean()), cgen, dev, row0, row1)
    # generate_v (:40-66)
    m0 = int(round(m * sparsity_degree_in_v))
    V = _blocked_gamma_device(m, k, _block_edges(m0, n_groups), _block_edges(k, n_groups), [beta] * n_groups,
                              (1. - theta) * beta, ggen, dev)
    # X = D * floor(U V^T), D_ij ~ Bernoulli(pi_j), pi_j ~ Beta(1, 1/z - 1) (:68-86)
    b = (1. / zero_inflation_level) - 1.
    pi_d = 1. - torch.rand((m,), generator=ggen, device=dev, dtype=torch.float64) ** (1. / b)      # Beta(1, b) by inverse cdf
    X = torch.empty((row1 - row0, m), dtype=torch.float32, device=dev)
    step = max(1, (1 << 24) // max(1, m))
    for r in range(0, row1 - row0, step):
        rate = torch.floor(U[r:r + step] @ V.T)
        keep = torch.rand(rate.shape, generator=cgen, device=dev, dtype=torch.float64) < pi_d[None, :]
        X[r:r + step] = (rate * keep).to(torch.float32)
    return X, U, V, labels
