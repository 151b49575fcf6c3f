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
