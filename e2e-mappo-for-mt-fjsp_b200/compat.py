"""Glue for feeding the B200 environment's native observation to the UNMODIFIED reference networks.

The reference actors take a dense `adj [B,N,N]` (float64 numpy), turn it into a sparse tensor and re-index it into one
block-diagonal `[B*N, B*N]` COO matrix (`aggr_obs`, model/gcn_mlp.py:305-320; called at model/actor_critic.py:139-140 and
155-156).  The environment here emits that adjacency in compact ELL form -- per destination op the job-predecessor
weight, the machine-predecessor weight and its source (`adj_w [B,N,2] f32`, `adj_src [B,N] i16`; the diagonal is 1).
`ell_to_block_sparse` builds the very same COO matrix from the ELL arrays directly, on whatever device they live, so a
dense [B,N,N] float64 array (10.4 KB per env at J6M6, 2.9 MB at J30M20) never exists and nothing crosses PCIe.
INTEGRATION.md shows the three-line change in the reference's `forward` that accepts it.
"""
from __future__ import annotations

import torch


def ell_to_block_sparse(adj_w: torch.Tensor, adj_src: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """adj_w [B,N,2], adj_src [B,N] -> coalesced sparse COO [B*N, B*N] with A[b*N + dst, b*N + src] = weight: what
    `aggr_obs(dense_adj.to_sparse(), N)` returns for the dense matrix the same observation stands for."""
    B, N, _ = adj_w.shape
    dev = adj_w.device
    row = torch.arange(B * N, device=dev)
    wj = adj_w[..., 0].reshape(-1)
    wm = adj_w[..., 1].reshape(-1)
    src = adj_src.reshape(-1).long()
    has_j = wj != 0
    has_m = src >= 0
    base = (row // N) * N
    rows = torch.cat((row, row[has_j], row[has_m]))
    cols = torch.cat((row, row[has_j] - 1, base[has_m] + src[has_m]))
    vals = torch.cat((torch.ones(B * N, device=dev, dtype=dtype), wj[has_j].to(dtype), wm[has_m].to(dtype)))
    return torch.sparse_coo_tensor(torch.stack((rows, cols)), vals, (B * N, B * N)).coalesce()
