"""Small seeded test graphs as numpy arrays (built with sparsebase_b200.synth on the CPU)."""
import numpy as np
import torch

from sparsebase_b200 import synth


def _np(t):
    return t.cpu().numpy()


def csr_of(n, row, col, nnz_dtype=np.int32):
    rp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rp, row.astype(np.int64) + 1, 1)
    return np.cumsum(rp).astype(nnz_dtype)


def vals_for(nnz, seed=1, dtype=np.float32):
    return _np(synth.hash_vals(nnz, seed)).astype(dtype)


def rmat(scale, edge_factor=16, seed=42):
    n, r, c = synth.rmat(scale, edge_factor, seed)
    return n, _np(r), _np(c)


def er(n, ppv=8, seed=43):
    n, r, c = synth.erdos_renyi(n, ppv, seed)
    return n, _np(r), _np(c)


def band(n, hb=31, density=0.5, seed=45, shuffle_seed=46):
    n, r, c = synth.band(n, hb, density, seed, shuffle_seed)
    return n, _np(r), _np(c)


def poisson(nx, ny):
    n, rp, col, vals = synth.poisson2d(nx, ny)
    return n, _np(rp), _np(col), _np(vals)


def random_rect(n, m, nnz, seed):
    """unique (row, col) pairs of an n x m matrix, sorted by (row, col)."""
    rng = np.random.default_rng(seed)
    key = np.unique(rng.integers(0, n * m, size=nnz, dtype=np.int64))
    return (key // m).astype(np.int32), (key % m).astype(np.int32)


def multi_component(seed=7):
    """several components of different shapes + isolated vertices + a self-loop-only vertex."""
    rng = np.random.default_rng(seed)
    parts = []
    off = 0
    # path of 50, cycle of 33, star of 40, 2 random blobs, grid 7x9
    def add(edges, cnt):
        nonlocal off
        e = np.asarray(edges, dtype=np.int64).reshape(-1, 2) + off
        parts.append(e)
        off += cnt
    add([(i, i + 1) for i in range(49)], 50)
    off += 3  # isolated
    add([(i, (i + 1) % 33) for i in range(33)], 33)
    add([(0, i) for i in range(1, 40)], 40)
    for cnt in (120, 77):
        e = rng.integers(0, cnt, size=(cnt * 3, 2))
        add(e, cnt)
        off += 2
    add([(y * 7 + x, y * 7 + x + 1) for y in range(9) for x in range(6)]
        + [(y * 7 + x, (y + 1) * 7 + x) for y in range(8) for x in range(7)], 63)
    add([(0, 0)], 1)  # vertex with only a self loop
    off += 1
    n = off
    e = np.concatenate(parts)
    perm = rng.permutation(n)
    r, c = perm[e[:, 0]], perm[e[:, 1]]
    rr, cc = np.concatenate([r, c]), np.concatenate([c, r])
    key = np.unique(rr * n + cc)
    return n, (key // n).astype(np.int32), (key % n).astype(np.int32)
