"""MT-FJSP instance data: the reference's distributions, two generators.

An *instance* is what the reference calls an "ability" sample
(instance/generate_allsize_mofjsp_dataset.py:126-296):

    t      [S, N, M] float64  processing time of op i on machine k   (negative = machine infeasible)
    p      [S, N, M] float64  processing power                        (negative where t is negative)
    transT [S, M, M] float64  transport time between machines, symmetric, zero diagonal
    edge   [S, E, W] int32    machine ids of each edge group, padded with -1 (W = largest group)

`reference_stream_instances` follows the reference generator's ``np.random`` call sequence, so
``seed=3`` / ``seed=1`` reproduce the shipped test / eval pickles bit for bit (checked in
tests/test_instances.py against hashes of those pickles).  `synthetic_instances` draws the same
distributions from a counter-based Philox stream keyed by (seed, first env index) in fixed blocks, so
any slice of a large batch can be regenerated on any rank without generating the rest.
"""
from __future__ import annotations

import numpy as np

ABILITY_SCOPE = dict(t_low=1, t_high=99, p_low=1, p_high=20, transT_in_low=1, transT_in_high=10,
                     transT_out_low=1, transT_out_high=20, weight_low=0.8, weight_high=1.2)  # instance/config_ins.json


def edge_groups(n_machine: int, n_edge: int) -> np.ndarray:
    """Contiguous split, last group takes the remainder (generate_...py:318-334, equal_edge=true).
    Returns int32 [E, W] padded with -1."""
    avg = n_machine // n_edge
    sizes = [avg] * (n_edge - 1) + [n_machine - avg * (n_edge - 1)]
    out = np.full((n_edge, max(sizes)), -1, dtype=np.int32)
    start = 0
    for g, s in enumerate(sizes):
        out[g, :s] = np.arange(start, start + s)
        start += s
    return out


def edge_of_machine(edge: np.ndarray, n_machine: int) -> np.ndarray:
    """[.., E, W] groups -> [.., M] 0-based group id of each machine."""
    edge = np.asarray(edge)
    lead = edge.shape[:-2]
    out = np.zeros(lead + (n_machine,), dtype=np.int32)
    E, W = edge.shape[-2:]
    flat = edge.reshape((-1, E, W))
    o = out.reshape((-1, n_machine))
    for g in range(E):
        for k in range(W):
            m = flat[:, g, k]
            ok = m >= 0
            o[np.nonzero(ok)[0], m[ok]] = g
    return out


def reference_stream_instances(samples: int, n_job: int, n_machine: int, n_edge: int, seed: int,
                               scope: dict = ABILITY_SCOPE) -> dict:
    """Same legacy ``np.random`` draw order as the reference generator (generate_...py:161-273)."""
    S, M, N = samples, n_machine, n_job * n_machine
    rs = np.random.RandomState(seed)
    avg_t = rs.uniform(scope["t_low"], scope["t_high"], size=(S, N))
    avg_p = rs.uniform(scope["p_low"], scope["p_high"], size=(S, N))
    w_t = rs.uniform(scope["weight_low"], scope["weight_high"], size=(S, N, M))
    w_p = rs.uniform(scope["weight_low"], scope["weight_high"], size=(S, N, M))
    rs.uniform(1, 5, size=(S, 1, M))  # standby power draw: consumed, never used by the env (singlestep.py:371)
    t = avg_t[:, :, None] * w_t
    p = avg_p[:, :, None] * w_p
    for s in range(S):  # generate_...py:204-210
        for i in range(N):
            k = rs.randint(0, M)
            idx = rs.choice(M, size=k, replace=False)
            t[s, i, idx] *= -1
    p = np.where(t < 0, -p, p)
    eg = edge_groups(M, n_edge)
    gid = edge_of_machine(eg, M)
    tt = np.zeros((S, M, M))
    for s in range(S):  # generate_...py:243-262
        raw = np.zeros((M, M))
        for i in range(M):
            for j in range(M):
                if i == j:
                    continue
                d = abs(int(gid[i]) - int(gid[j]))
                if d == 0:
                    raw[i, j] = rs.uniform(scope["transT_in_low"], scope["transT_in_high"], size=1).item()
                else:
                    raw[i, j] = rs.uniform(scope["transT_in_high"] * d, scope["transT_out_high"] * d, size=1).item()
        U = np.triu(raw, k=1)
        tt[s] = U + U.T - np.diag(np.diag(raw))
    edge = np.broadcast_to(eg, (S,) + eg.shape).copy()
    return dict(t=t, p=p, transT=tt, edge=edge)


def save_instances(path: str, data: dict) -> None:
    """The reference's wire format (generate_...py:286-296): a pickled list of four arrays
    ``[t [S,N,M] f64, p [S,N,M] f64, transT [S,M,M] f64, edge [S,E,W] int64]``.  Ragged edge groups (M % E != 0) are
    written padded with -1, which the reference generator cannot express at all ("J10M10E3")."""
    import pickle

    arr = [np.asarray(data["t"], dtype=np.float64), np.asarray(data["p"], dtype=np.float64),
           np.asarray(data["transT"], dtype=np.float64), np.asarray(data["edge"], dtype=np.int64)]
    with open(path, "wb") as f:
        pickle.dump(arr, f)


def load_instances(path: str) -> dict:
    """Reads a reference instance pickle (generate_...py:298-316: ``t, p, transT, edge = pickle.load(f)``)."""
    import pickle

    with open(path, "rb") as f:
        t, p, tt, edge = pickle.load(f)
    t, p, tt = (np.ascontiguousarray(x, dtype=np.float64) for x in (t, p, tt))
    edge = np.asarray(edge)
    if edge.dtype == object:  # ragged lists written by old numpy: pad with -1
        W = max(len(g) for inst in edge for g in inst)
        pad = np.full((len(edge), len(edge[0]), W), -1, dtype=np.int32)
        for s_, inst in enumerate(edge):
            for g_, grp in enumerate(inst):
                pad[s_, g_, :len(grp)] = grp
        edge = pad
    return dict(t=t, p=p, transT=tt, edge=np.ascontiguousarray(edge, dtype=np.int32))


_BLOCK = 1024


def synthetic_instances(first_env: int, count: int, n_job: int, n_machine: int, n_edge: int, seed: int,
                        scope: dict = ABILITY_SCOPE) -> dict:
    """Envs [first_env, first_env+count) of the synthetic batch `seed`: same
    distributions as the reference generator, Philox stream per block of 1024 envs."""
    M, N = n_machine, n_job * n_machine
    eg = edge_groups(M, n_edge)
    gid = edge_of_machine(eg, M)
    dist = np.abs(gid[:, None] - gid[None, :]).astype(np.float64)
    lo = np.where(dist == 0, scope["transT_in_low"], scope["transT_in_high"] * dist)
    hi = np.where(dist == 0, scope["transT_in_high"], scope["transT_out_high"] * dist)
    iu = np.triu_indices(M, k=1)
    ts, ps, tts = [], [], []
    b0 = (first_env // _BLOCK) * _BLOCK
    end = first_env + count
    while b0 < end:
        g = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, b0 // _BLOCK]))
        avg_t = g.uniform(scope["t_low"], scope["t_high"], size=(_BLOCK, N))
        avg_p = g.uniform(scope["p_low"], scope["p_high"], size=(_BLOCK, N))
        t = avg_t[:, :, None] * g.uniform(scope["weight_low"], scope["weight_high"], size=(_BLOCK, N, M))
        p = avg_p[:, :, None] * g.uniform(scope["weight_low"], scope["weight_high"], size=(_BLOCK, N, M))
        k = g.integers(0, M, size=(_BLOCK, N))  # number of infeasible machines, 0..M-1
        rank = np.argsort(np.argsort(g.random((_BLOCK, N, M)), axis=-1), axis=-1)
        neg = rank < k[:, :, None]
        t = np.where(neg, -t, t)
        p = np.where(neg, -p, p)
        tt = np.zeros((_BLOCK, M, M))
        u = g.random((_BLOCK, len(iu[0])))
        tt[:, iu[0], iu[1]] = lo[iu] + (hi[iu] - lo[iu]) * u
        tt = tt + np.transpose(tt, (0, 2, 1))
        s0, s1 = max(first_env, b0) - b0, min(end, b0 + _BLOCK) - b0
        ts.append(t[s0:s1]); ps.append(p[s0:s1]); tts.append(tt[s0:s1])
        b0 += _BLOCK
    t = np.concatenate(ts); p = np.concatenate(ps); tt = np.concatenate(tts)
    edge = np.broadcast_to(eg, (count,) + eg.shape).copy()
    return dict(t=t, p=p, transT=tt, edge=edge)


def random_weights(first_env: int, count: int, seed: int, episode: int = 0) -> np.ndarray:
    """Per-env reward weights, three U(0,1) draws normalised to sum 1 (singlestep.py:1255-1259),
    from a counter-based stream so shards agree with the unsharded batch."""
    out = np.empty((count, 3))
    b0 = (first_env // _BLOCK) * _BLOCK
    end = first_env + count
    pos = 0
    while b0 < end:
        g = np.random.Generator(np.random.Philox(key=[(seed ^ 0x5DEECE66D) & 0xFFFFFFFFFFFFFFFF,
                                                      (episode << 32) | (b0 // _BLOCK)]))
        w = g.random((_BLOCK, 3))
        w = w / np.sum(w, axis=-1, keepdims=True)
        s0, s1 = max(first_env, b0) - b0, min(end, b0 + _BLOCK) - b0
        out[pos:pos + s1 - s0] = w[s0:s1]
        pos += s1 - s0
        b0 += _BLOCK
    return out


def device_instances(first_env: int, count: int, n_job: int, n_machine: int, n_edge: int, seed: int, device=None) -> dict:
    """Envs [first_env, first_env + count) of the device-generated synthetic batch `seed`: the reference's distributions
    drawn by `mtfjsp_generate_instances` (csrc/mtfjsp_env.cu `instance_gen_kernel`) straight into device tensors -- a
    65,536-env J6M6 batch is 226 MB of tables that never cross PCIe.  Counter-based like `synthetic_instances` (a slice
    equals the same rows of the whole batch) but a different stream: the two generators are not interchangeable."""
    import ctypes as C

    import torch

    from . import _lib

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    N, M, E = n_job * n_machine, n_machine, n_edge
    W = edge_groups(M, E).shape[1]
    t = torch.empty((count, N, M), dtype=torch.float64, device=dev)
    p = torch.empty_like(t)
    tt = torch.empty((count, M, M), dtype=torch.float64, device=dev)
    edge = torch.empty((count, E, W), dtype=torch.int32, device=dev)
    ptr = lambda x: C.c_void_p(x.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mtfjsp_generate_instances(count, n_job, M, E, seed & 0xFFFFFFFFFFFFFFFF, first_env, ptr(t), ptr(p),
                                                        ptr(tt), ptr(edge), W, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "mtfjsp_generate_instances")
    return dict(t=t, p=p, transT=tt, edge=edge)
