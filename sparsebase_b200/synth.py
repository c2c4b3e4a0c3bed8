"""Seeded synthetic inputs of the BASELINE.json shapes.

All graphs are square, 0-based, structurally symmetric and free of duplicate (row, col)
entries (duplicates make the reference's own result unspecified).  Generators are
written with torch ops so that the large configurations can be built directly in HBM; they are
input plumbing, not part of the measured path.

Every generator returns ``(n, row, col)`` as a (row, col)-sorted COO with int32 indices
(``poisson2d`` returns CSR directly) on ``device``.
"""
import torch


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def hash_vals(count, seed=1, device="cpu", dtype=torch.float32):
    """vals[k] = float(hash32(seed, k) & 0xFFFF) + 0.5 -- distinct-ish, exactly representable."""
    k = torch.arange(count, device=device, dtype=torch.int64)
    x = (k * 0x9E3779B1 + seed * 0x85EBCA77) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & 0xFFFFFFFF
    x = ((x ^ (x >> 15)) * 0x846CA68B) & 0xFFFFFFFF
    x = x ^ (x >> 16)
    return ((x & 0xFFFF).to(torch.float64) + 0.5).to(dtype)


def _finish(n, r, c, symmetrise=True, drop_self=True, id_dtype=torch.int32):
    """symmetrise + dedupe + sort by (row, col).  On a CUDA device this is the library's own
    edge-list -> COO path (sb200_edges_to_coo: EdgeListReader semantics, radix sort + unique in
    libsb200.so, no library sort); on the CPU (test plumbing without a GPU) unique() on
    row*n + col gives the same arrays."""
    if r.is_cuda:
        from . import lib
        _, _, row, col, _ = lib.edges_to_coo(r.to(id_dtype), c.to(id_dtype), None,
                                             remove_duplicates=True, remove_self_edges=drop_self,
                                             read_undirected=symmetrise, square=True)
        return n, row.clone(), col.clone()
    if drop_self:
        keep = r != c
        r, c = r[keep], c[keep]
    if symmetrise:
        r, c = torch.cat([r, c]), torch.cat([c, r])
    key = torch.unique(r.to(torch.int64) * n + c.to(torch.int64))  # sorted, dedup'd
    return n, (key // n).to(id_dtype), (key % n).to(id_dtype)


def rmat(scale, edge_factor=16, seed=42, device="cpu", a=0.57, b=0.19, c=0.19, symmetrise=True):
    """R-MAT (a,b,c,d) graph with 2^scale vertices and edge_factor*2^scale generated edges."""
    n = 1 << scale
    e = edge_factor * n
    g = _gen(seed, device)
    r = torch.zeros(e, dtype=torch.int64, device=device)
    cc = torch.zeros(e, dtype=torch.int64, device=device)
    for _ in range(scale):
        u = torch.rand(e, generator=g, device=device)
        rbit = (u >= a + b).to(torch.int64)
        cbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)
        r = (r << 1) | rbit
        cc = (cc << 1) | cbit
    return _finish(n, r, cc, symmetrise=symmetrise)


def erdos_renyi(n, pairs_per_vertex=8, seed=43, device="cpu"):
    """G(n, M): pairs_per_vertex*n uniformly random undirected pairs, symmetrised, dedup'd."""
    g = _gen(seed, device)
    e = pairs_per_vertex * n
    r = torch.randint(0, n, (e,), generator=g, device=device, dtype=torch.int64)
    c = torch.randint(0, n, (e,), generator=g, device=device, dtype=torch.int64)
    return _finish(n, r, c)


def band(n, half_bandwidth=31, density=0.5, seed=45, shuffle_seed=46, device="cpu"):
    """|i-j| <= half_bandwidth, diagonal always present, each off-diagonal symmetric pair kept
    with probability `density`; both axes relabelled by one random permutation when
    shuffle_seed is not None."""
    g = _gen(seed, device)
    rows, cols = [torch.arange(n, device=device, dtype=torch.int64)], [
        torch.arange(n, device=device, dtype=torch.int64)]
    for d in range(1, half_bandwidth + 1):
        keep = torch.rand(n - d, generator=g, device=device) < density
        i = torch.nonzero(keep).flatten()
        rows += [i, i + d]
        cols += [i + d, i]
    r, c = torch.cat(rows), torch.cat(cols)
    if shuffle_seed is not None:
        perm = torch.randperm(n, generator=_gen(shuffle_seed, device), device=device)
        r, c = perm[r], perm[c]
    return _finish(n, r, c, symmetrise=False, drop_self=False)


def poisson2d(nx, ny, device="cpu", id_dtype=torch.int32, nnz_dtype=torch.int32,
              val_dtype=torch.float32):
    """5-point Laplacian on an nx x ny grid (row-major ids, diagonal included) as CSR with
    ascending columns: returns (n, row_ptr, col, vals) with vals 4 / -1."""
    n = nx * ny
    v = torch.arange(n, device=device, dtype=torch.int64)
    x, y = v % nx, v // nx
    cand = torch.stack([v - nx, v - 1, v, v + 1, v + nx], dim=1)
    ok = torch.stack([y > 0, x > 0, torch.ones_like(x, dtype=torch.bool), x < nx - 1,
                      y < ny - 1], dim=1)
    col = cand[ok].to(id_dtype)
    deg = ok.sum(dim=1)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptr[1:] = torch.cumsum(deg, 0)
    vals = torch.where(cand == v[:, None], 4.0, -1.0)[ok].to(val_dtype)
    return n, row_ptr.to(nnz_dtype), col, vals


def csr_from_sorted_coo(n, row, nnz_dtype=torch.int32):
    """row_ptr of a (row, col)-sorted COO -- input plumbing for benches/tests that need CSR."""
    counts = torch.bincount(row.to(torch.int64), minlength=n)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=row.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return row_ptr.to(nnz_dtype)
