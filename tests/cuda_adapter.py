"""Test helper: presents the CUDA environment through the numpy-in / numpy-out interface of the oracle, so the
same parity routine (tests/test_oracle_golden.py:check_replay) drives both."""
import importlib

import numpy as np
import torch

_env = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
BatchedMTFJSPEnv = _env.BatchedMTFJSPEnv
MASK_ESA = _env.MASK_ESA


class NumpyEnvAdapter:
    def __init__(self, B, J, M, E, left_shift=True, fused=False, **kw):
        self.env = BatchedMTFJSPEnv(B, J, M, E, left_shift=left_shift, obs_dtype=torch.float64, **kw)
        self.B, self.J, self.M, self.E, self.N = B, J, M, E, J * M
        self.fused = fused
        self._obs_cache = None

    def load(self, t, p, tt, edge):
        self.env.load(t, p, tt, edge)

    def scaler_init(self):
        self.env.scaler_init()

    def scaler_reset(self):
        self.env.scaler_reset()

    def reset(self, weights):
        self.env.reset(weights)
        self._obs_cache = None

    def _i32(self, x):
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int32)).to(self.env.device)

    def _collect_obs(self):
        e = self.env
        N = self.N
        aw = e.adj_w.cpu().numpy().astype(np.float64)
        src = e.adj_src.cpu().numpy().astype(np.int32)
        idx = np.full((self.B, N, 3), -1, dtype=np.int32)
        w = np.zeros((self.B, N, 3))
        idx[:, :, 0] = np.arange(N)[None, :]
        w[:, :, 0] = 1.0
        has_j = aw[:, :, 0] != 0
        idx[:, :, 1] = np.where(has_j, np.arange(N)[None, :] - 1, -1)
        w[:, :, 1] = aw[:, :, 0]
        idx[:, :, 2] = src
        w[:, :, 2] = np.where(src >= 0, aw[:, :, 1], 0.0)
        return dict(task_fea=e.task_fea.cpu().numpy(), mach_fea=e.mach_fea.cpu().numpy(), ell_idx=idx, ell_w=w,
                    job_mask=e.job_mask.cpu().numpy(), candidate=e.candidate.cpu().numpy())

    def step(self, op, mach):
        e = self.env
        if self.fused:
            e.step_obs(self._i32(op), self._i32(mach), mask_mode=self._mm)
            self._obs_cache = self._collect_obs()
        else:
            e.step(self._i32(op), self._i32(mach))
            self._obs_cache = None
        return (e.reward5.cpu().numpy(), e.scaled4.cpu().numpy(), e.done.cpu().numpy(), e.invalid.cpu().numpy())

    _mm = MASK_ESA

    def obs(self, mask_mode=1):
        self._mm = mask_mode
        if self.fused and self._obs_cache is not None:
            return self._obs_cache
        self.env.obs(mask_mode)
        return self._collect_obs()

    def mfea1(self, op):
        m1, mm = self.env.mfea1(self._i32(op))
        return m1.cpu().numpy(), mm.cpu().numpy()

    def dense_adj(self):
        return self.env.dense_adj(torch.float64).cpu().numpy()

    def costs(self):
        return self.env.costs().cpu().numpy()

    def export_state(self):
        return {k: v.cpu().numpy() for k, v in self.env.export_state().items()}

    def export_scaler(self):
        return {k: v.cpu().numpy() for k, v in self.env.export_scaler().items()}
