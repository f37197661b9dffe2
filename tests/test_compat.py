"""compat.ell_to_block_sparse builds the block-diagonal sparse adjacency of the reference networks (aggr_obs,
model/gcn_mlp.py:305-320) from the environment's ELL arrays: checked against `dense.to_sparse()` + the reference's
re-indexing on the adjacencies of a reference replay dump."""
import importlib
import os

import numpy as np
import torch

from tests.test_ppo import dense_to_ell

GOLD = os.path.join(os.path.dirname(__file__), "golden")
compat = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.compat")


def _aggr_obs(obs_mb, n_node):
    """The reference's aggr_obs, restated (model/gcn_mlp.py:305-320)."""
    idxs, vals = obs_mb.coalesce().indices(), obs_mb.coalesce().values()
    idx = torch.stack((idxs[1] + idxs[0] * n_node, idxs[2] + idxs[0] * n_node))
    return torch.sparse_coo_tensor(idx, vals, (obs_mb.shape[0] * n_node,) * 2).coalesce()


def test_block_sparse_from_ell_equals_reference_aggr_obs():
    for name in ("replay_j6m6_ls_fin_lowest.npz", "replay_j10m10e3_ls_esa_mixed.npz", "replay_j3m4_ls_fin_mixed.npz"):
        g = np.load(os.path.join(GOLD, name))
        M = int(g["M"])
        adj = g["adj"][0].astype(np.float64)                       # [steps, B, N, N] dense, as the reference emits it
        S, B, N, _ = adj.shape
        for s in (0, S // 3, S - 1):
            aw, asrc = dense_to_ell(adj[s], M)
            mine = compat.ell_to_block_sparse(torch.tensor(aw), torch.tensor(asrc))
            ref = _aggr_obs(torch.from_numpy(adj[s]).to_sparse(), N)
            assert torch.equal(mine.indices(), ref.indices()), (name, s)
            assert torch.equal(mine.values(), ref.values()), (name, s)
            h = torch.randn(B * N, 5, dtype=torch.float64)
            assert torch.equal(torch.mm(mine, h), torch.mm(ref, h))
