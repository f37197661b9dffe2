"""ctypes front-end of the C restatement ``oracle/mtfjsp_oracle.c``.

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mtfjsp_oracle.c")
_LIB = os.path.join(_HERE, "libmtfjsp_oracle.so")


def build(force: bool = False) -> str:
    """gcc -O2, no fast-math, no FMA contraction (the reference is plain IEEE double arithmetic)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
               "-o", _LIB, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_create.argtypes = [C.c_int] * 5
        _lib.oracle_destroy.argtypes = [C.c_void_p]
        _lib.oracle_set_params.argtypes = [C.c_void_p] + [C.c_double] * 5 + [C.c_int]
        _lib.oracle_load.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        _lib.oracle_scaler_init.argtypes = [C.c_void_p]
        _lib.oracle_scaler_reset.argtypes = [C.c_void_p]
        _lib.oracle_reset.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        _lib.oracle_obs.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
        _lib.oracle_mfea1.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        _lib.oracle_dense_adj.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_costs.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_export_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.oracle_export_scaler.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.oracle_rollout_random.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_int] + [C.c_void_p] * 9
        _lib.oracle_np_sum.restype = C.c_double
        _lib.oracle_np_sum.argtypes = [C.c_void_p, C.c_long]
        _lib.oracle_rand_u32.restype = C.c_uint32
        _lib.oracle_rand_u32.argtypes = [C.c_uint64] * 4
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pad_edge(edge, B, E):
    """ragged per-env edge groups -> int32 [B,E,W] padded with -1."""
    if isinstance(edge, np.ndarray) and edge.ndim == 3 and edge.dtype != object:
        return np.ascontiguousarray(edge.astype(np.int32))
    W = max(len(g) for e in edge for g in e)
    out = np.full((B, E, W), -1, dtype=np.int32)
    for b in range(B):
        for g in range(E):
            row = list(edge[b][g])
            out[b, g, : len(row)] = row
    return out


class OracleEnv:
    """Batch of reference-semantics environments on the CPU."""

    def __init__(self, B, J, M, E, left_shift=True, weights=(0.4, 0.4, 0.2), divisor=1.0, gamma=0.99, nthreads=1):
        self.B, self.J, self.M, self.E, self.N = B, J, M, E, J * M
        self._h = lib().oracle_create(B, J, M, E, int(bool(left_shift)))
        lib().oracle_set_params(self._h, weights[0], weights[1], weights[2], divisor, gamma, nthreads)

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def load(self, t, p, tt, edge):
        t = np.ascontiguousarray(t, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        tt = np.ascontiguousarray(tt, dtype=np.float64)
        assert t.shape == (self.B, self.N, self.M) and tt.shape == (self.B, self.M, self.M)
        e = pad_edge(edge, self.B, self.E)
        lib().oracle_load(self._h, _p(t), _p(p), _p(tt), _p(e), e.shape[2])
        self._keep = (t, p, tt, e)

    def scaler_init(self):
        lib().oracle_scaler_init(self._h)

    def scaler_reset(self):
        lib().oracle_scaler_reset(self._h)

    def reset(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (self.B, 3)
        lib().oracle_reset(self._h, _p(w))

    def step(self, op, mach):
        op = np.ascontiguousarray(op, dtype=np.int32)
        mach = np.ascontiguousarray(mach, dtype=np.int32)
        r5 = np.zeros((self.B, 5)); s4 = np.zeros((self.B, 4))
        done = np.zeros(self.B, dtype=np.uint8); inv = np.zeros(self.B, dtype=np.uint8)
        lib().oracle_step(self._h, _p(op), _p(mach), _p(r5), _p(s4), _p(done), _p(inv))
        return r5, s4, done, inv

    def obs(self, mask_mode=1):
        B, N, M, J = self.B, self.N, self.M, self.J
        tfea = np.zeros((B, N, 12)); mfea = np.zeros((B, M, 8))
        ix = np.zeros((B, N, 3), dtype=np.int32); w = np.zeros((B, N, 3))
        jm = np.zeros((B, J), dtype=np.uint8); cand = np.zeros((B, J), dtype=np.int32)
        lib().oracle_obs(self._h, _p(tfea), _p(mfea), _p(ix), _p(w), _p(jm), _p(cand), mask_mode)
        return dict(task_fea=tfea, mach_fea=mfea, ell_idx=ix, ell_w=w, job_mask=jm, candidate=cand)

    def mfea1(self, op):
        op = np.ascontiguousarray(op, dtype=np.int32)
        out = np.zeros((self.B, self.M, 6)); mm = np.zeros((self.B, self.M), dtype=np.uint8)
        lib().oracle_mfea1(self._h, _p(op), _p(out), _p(mm))
        return out, mm

    def dense_adj(self):
        adj = np.zeros((self.B, self.N, self.N))
        lib().oracle_dense_adj(self._h, _p(adj))
        return adj

    def costs(self):
        c = np.zeros((self.B, 4))
        lib().oracle_costs(self._h, _p(c))
        return c

    def export_state(self):
        B, N, M = self.B, self.N, self.M
        mach = np.zeros((B, N), dtype=np.int32); st = np.zeros((B, N)); ft = np.zeros((B, N))
        routes = np.zeros((B, M, N), dtype=np.int32)
        lib().oracle_export_state(self._h, _p(mach), _p(st), _p(ft), _p(routes))
        return dict(mach=mach, st=st, ft=ft, routes=routes)

    def export_scaler(self):
        B = self.B
        R = np.zeros((B, 4)); mean = np.zeros((B, 4)); S = np.zeros((B, 4)); n = np.zeros(B, dtype=np.int64)
        lib().oracle_export_scaler(self._h, _p(R), _p(mean), _p(S), _p(n))
        return dict(R=R, mean=mean, S=S, n=n)

    def rollout_random(self, steps, seed, env_offset=0, mask_mode=1, record_actions=False):
        B, N, M = self.B, self.N, self.M
        acts = np.full((steps, B, 2), -1, dtype=np.int32) if record_actions else None
        tfea = np.zeros((B, N, 12)); mfea = np.zeros((B, M, 8))
        ix = np.zeros((B, N, 3), dtype=np.int32); w = np.zeros((B, N, 3)); mfea1 = np.zeros((B, M, 6))
        r5 = np.zeros((B, 5)); s4 = np.zeros((B, 4)); done = np.zeros(B, dtype=np.uint8)
        lib().oracle_rollout_random(self._h, steps, seed, env_offset, mask_mode, _p(acts), _p(tfea), _p(mfea), _p(ix),
                                    _p(w), _p(mfea1), _p(r5), _p(s4), _p(done))
        return dict(actions=acts, task_fea=tfea, mach_fea=mfea, ell_idx=ix, ell_w=w, mfea1=mfea1, reward5=r5,
                    scaled4=s4, done=done)


def np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().oracle_np_sum(_p(a), a.size)


def rand_u32(seed, env, step, stream):
    return lib().oracle_rand_u32(seed, env, step, stream)


def max_threads():
    return lib().oracle_max_threads()
