"""Support construction: drop-in for ``GCN.Adj_Preprocessor`` (reference ``GCN.py:50-135``) plus the
sparse-native path of SURVEY.md section 8(f)-1.

``process(adj)`` returns the same dense ``(K+1, N, N)`` stack the reference returns (so ``Main.py:49-55``
runs unchanged); ``process_sparse(adj)`` returns a :class:`~stmgcn_b200.graph.ChebSupports` holding only
the rescaled Laplacian as CSR -- no ``N x N`` polynomial is ever built (the reference needs K dense
``N^3`` products and 19 GB of supports at 16384 regions).

``lambda_max``: the reference calls ``torch.eig`` (``GCN.py:117``), which no longer exists in torch >= 1.13;
its bare ``except`` then uses 2 (``GCN.py:119-121``).  ``lambda_max="reference"`` (default) reproduces that
behaviour, a float fixes the value, ``"power"`` estimates the largest eigenvalue by power iteration.
"""
from __future__ import annotations

from typing import Union

import torch

from .graph import ChebSupports


class Adj_Preprocessor(object):
    def __init__(self, kernel_type: str, K: int, lambda_max: Union[str, float] = "reference"):
        if kernel_type not in ("chebyshev", "localpool", "random_walk_diffusion"):
            raise ValueError('Invalid kernel_type. Must be one of [chebyshev, localpool, random_walk_diffusion].')
        self.kernel_type = kernel_type
        self.K = K if kernel_type != "localpool" else 1      # GCN.py:54
        self.lambda_max = lambda_max

    # ---- normalisations (GCN.py:99-111), written with broadcasting instead of diag/mm -------------------
    @staticmethod
    def symmetric_normalize(A: torch.Tensor) -> torch.Tensor:
        d = A.sum(dim=1).pow(-0.5)
        return d.unsqueeze(1) * A * d.unsqueeze(0)

    @staticmethod
    def random_walk_normalize(A: torch.Tensor) -> torch.Tensor:
        d = A.sum(dim=1).pow(-1)
        d = torch.where(torch.isinf(d), torch.zeros_like(d), d)
        return d.unsqueeze(1) * A

    def _lambda(self, lap: torch.Tensor) -> float:
        if isinstance(self.lambda_max, (int, float)):
            return float(self.lambda_max)
        if self.lambda_max == "reference":
            return 2.0
        if self.lambda_max == "power":
            v = torch.ones(lap.shape[0], dtype=lap.dtype, device=lap.device)
            lam = 2.0
            for _ in range(200):
                w = lap @ v
                nrm = float(w.norm())
                if nrm == 0.0:
                    break
                lam, v = nrm / max(float(v.norm()), 1e-30), w / nrm
            return lam
        raise ValueError(f"lambda_max={self.lambda_max!r}")

    def rescale_laplacian(self, L: torch.Tensor) -> torch.Tensor:
        eye = torch.eye(L.shape[0], dtype=L.dtype, device=L.device)
        return (2.0 / self._lambda(L)) * L - eye

    def _polynomials(self, x: torch.Tensor):
        """``T_0 .. T_K`` of the matrix ``x`` (GCN.py:125-135)."""
        polys = [torch.eye(x.shape[0], dtype=x.dtype, device=x.device)]
        if self.K >= 1:
            polys.append(x)
        while len(polys) < self.K + 1:
            polys.append(2 * (x @ polys[-1]) - polys[-2])
        return polys

    def process(self, adj: torch.Tensor) -> torch.Tensor:
        """(N,N) adjacency -> (K_supports, N, N) dense stack, as ``GCN.py:57-97``."""
        if self.kernel_type == "localpool":
            a_norm = self.symmetric_normalize(adj)
            kernels = [torch.eye(adj.shape[0], dtype=adj.dtype, device=adj.device) + a_norm]
        elif self.kernel_type == "chebyshev":
            a_norm = self.symmetric_normalize(adj)
            lap = torch.eye(adj.shape[0], dtype=adj.dtype, device=adj.device) - a_norm
            kernels = self._polynomials(self.rescale_laplacian(lap))
        else:   # random_walk_diffusion: K+1 polynomials of P^T (the reference's forward-only variant)
            kernels = self._polynomials(self.random_walk_normalize(adj).T)
        return torch.stack(kernels, dim=0)

    def process_sparse(self, adj: torch.Tensor) -> ChebSupports:
        """Chebyshev supports without dense polynomials: only ``L~`` is formed, as CSR.

        ``adj`` may be dense ``(N,N)`` or a sparse COO/CSR tensor.  The result is accepted wherever the
        dense stack is (``GCN.forward``, ``ST_MGCN.forward``'s ``sta_adj_list``)."""
        if self.kernel_type != "chebyshev":
            raise ValueError("process_sparse is defined for kernel_type='chebyshev' only")
        coo = adj.to_sparse_coo().coalesce() if adj.layout != torch.sparse_coo else adj.coalesce()
        n = coo.shape[0]
        row, col = coo.indices()
        val = coo.values().to(torch.float32)
        deg = torch.zeros(n, dtype=torch.float32, device=val.device).index_add_(0, row, val)
        d = deg.pow(-0.5)
        a_norm = d[row] * val * d[col]
        if self.lambda_max == "reference":
            lam = 2.0
        elif isinstance(self.lambda_max, (int, float)):
            lam = float(self.lambda_max)
        else:
            lam = self._lambda_sparse(n, row, col, a_norm)
        scale = 2.0 / lam
        # L~ = scale * (I - A_norm) - I  => off-diagonal -scale*A_norm, diagonal (scale - 1) - scale*A_norm_ii
        diag_val = scale - 1.0
        idx_r, idx_c, vals = row, col, -scale * a_norm
        if diag_val != 0.0:
            ar = torch.arange(n, device=val.device)
            idx_r, idx_c = torch.cat([idx_r, ar]), torch.cat([idx_c, ar])
            vals = torch.cat([vals, torch.full((n,), diag_val, dtype=torch.float32, device=val.device)])
        lt = torch.sparse_coo_tensor(torch.stack([idx_r, idx_c]), vals, (n, n)).coalesce()
        keep = lt.values() != 0
        r, c, v = lt.indices()[0][keep], lt.indices()[1][keep], lt.values()[keep]
        counts = torch.bincount(r, minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=v.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        return ChebSupports(n, self.K + 1, rowptr.to(torch.int32), c.to(torch.int32), v)

    @staticmethod
    def _lambda_sparse(n, row, col, a_norm) -> float:
        v = torch.ones(n, dtype=torch.float32, device=a_norm.device)
        lam = 2.0
        for _ in range(200):
            w = v - torch.zeros_like(v).index_add_(0, row, a_norm * v[col])      # (I - A_norm) v
            nrm = float(w.norm())
            if nrm == 0.0:
                break
            lam, v = nrm / max(float(v.norm()), 1e-30), w / nrm
        return lam
