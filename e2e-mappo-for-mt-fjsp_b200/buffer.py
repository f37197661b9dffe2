"""Device-resident rollout buffer with sparse (ELL) adjacency and the 4-stream GAE.

The reference's replay buffer keeps two dense float64 adjacencies per buffered step (trainer/replaybuffer.py:31, 36):
29.9 MB each at B = 16 and 122 GB each at B = 65,536 (N = 36, 180 steps).  Here a step stores the env's native
observation: F32 feature rows plus 10 bytes per node of ELL adjacency, i.e. 2.3 KB per env-step at J6M6 instead of
24.4 KB, and everything stays on the GPU.  `gae4` replaces the python loops of
algorithm/ppo_algorithm.py:438-536; advantage normalisation uses sums that are allreduced across ranks, so sharded
runs normalise over the same [steps, B_total] block a single GPU would."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gae4(r, v, v_next, done, gamma=0.99, lam=0.98, normalize=True):
    """r, v, v_next [T,B,4] f32 (mk, pt, tt, idle), done [T,B] -> advantages [T,B,4] f32.
    With torch.distributed initialised the normalisation statistics are summed over all ranks."""
    T, B, K = r.shape
    assert K == 4
    r, v, v_next = r.contiguous().float(), v.contiguous().float(), v_next.contiguous().float()
    done = done.contiguous().float()
    adv = torch.empty_like(r)
    stats = torch.zeros(8, dtype=torch.float64, device=r.device)
    check(_lib.lib().mtfjsp_gae4(_ptr(r), _ptr(v), _ptr(v_next), _ptr(done), _ptr(adv), _ptr(stats), T, B, gamma, lam,
                                 _stream()), "mtfjsp_gae4")
    if normalize:
        count = float(T * B)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            cnt = torch.tensor([count], dtype=torch.float64, device=r.device)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            count = float(cnt.item())
        check(_lib.lib().mtfjsp_adv_normalize(_ptr(adv), _ptr(stats), count, T, B, _stream()), "mtfjsp_adv_normalize")
    return adv


class RolloutBuffer:
    """Pre-allocated [T, B, ...] device tensors for one PPO buffer (reference fields: trainer/replaybuffer.py:18-204)."""

    def __init__(self, T, env, hidden=128):
        B, N, M, J, dev = env.B, env.N, env.M, env.J, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.T, self.t = T, 0
        self.task_fea = torch.empty((T, B, N, 12), **f32)
        self.adj_w = torch.empty((T, B, N, 2), **f32)
        self.adj_src = torch.empty((T, B, N), dtype=torch.int16, device=dev)
        self.mach_fea1 = torch.empty((T, B, M, 6), **f32)
        self.mach_fea2 = torch.empty((T, B, M, 8), **f32)
        self.candidate = torch.empty((T, B, J), dtype=torch.int32, device=dev)
        self.job_mask = torch.empty((T, B, J), dtype=torch.uint8, device=dev)
        self.mach_mask = torch.empty((T, B, M), dtype=torch.uint8, device=dev)
        self.op = torch.empty((T, B), dtype=torch.int32, device=dev)
        self.mach = torch.empty((T, B), dtype=torch.int32, device=dev)
        self.log_a = torch.empty((T, B), **f32)
        self.m_log_a = torch.empty((T, B), **f32)
        self.v = torch.empty((T, B, 4), **f32)        # (mk, pt, tt, idle) = job_v[0], mch_v[0], mch_v[1], job_v[1]
        self.reward4 = torch.empty((T, B, 4), **f32)  # scaled (mk, pt, tt, idle)
        self.done = torch.empty((T, B), **f32)
        self.h_mch_in = torch.empty((T, B, hidden), **f32)

    def bytes_per_env_step(self):
        tot = sum(x.element_size() * x[0, 0].numel() for x in
                  (self.task_fea, self.adj_w, self.adj_src, self.mach_fea1, self.mach_fea2, self.candidate, self.job_mask,
                   self.mach_mask, self.op, self.mach, self.log_a, self.m_log_a, self.v, self.reward4, self.done, self.h_mch_in))
        return tot

    def store_pre(self, env, rollout):
        """Call BEFORE rollout.step(): the observation the actors are about to see."""
        t = self.t
        self.task_fea[t].copy_(env.task_fea); self.adj_w[t].copy_(env.adj_w); self.adj_src[t].copy_(env.adj_src)
        self.mach_fea2[t].copy_(env.mach_fea); self.candidate[t].copy_(env.candidate); self.job_mask[t].copy_(env.job_mask)
        self.h_mch_in[t].copy_(rollout.h_mch)

    def store_post(self, env, rollout):
        """Call AFTER rollout.step(): actions, log-probs, values, scaled rewards (order mk, idle, pt, tt in the env)."""
        t = self.t
        self.mach_fea1[t].copy_(env.mfea1_buf); self.mach_mask[t].copy_(env.mach_mask)
        self.op[t].copy_(env.op); self.mach[t].copy_(env.mach)
        self.log_a[t].copy_(rollout.log_a); self.m_log_a[t].copy_(rollout.m_log_a)
        self.v[t, :, 0].copy_(rollout.job_v[:, 0]); self.v[t, :, 1].copy_(rollout.mch_v[:, 0])
        self.v[t, :, 2].copy_(rollout.mch_v[:, 1]); self.v[t, :, 3].copy_(rollout.job_v[:, 1])
        s4 = env.scaled4
        self.reward4[t, :, 0].copy_(s4[:, 0]); self.reward4[t, :, 1].copy_(s4[:, 2])
        self.reward4[t, :, 2].copy_(s4[:, 3]); self.reward4[t, :, 3].copy_(s4[:, 1])
        self.done[t].copy_(env.done)
        self.t = t + 1

    def advantages(self, v_last, gamma=0.99, lam=0.98):
        """v_last [B,4]: bootstrap values after the final stored step (Run.py:455-475)."""
        T = self.t
        v_next = torch.cat((self.v[1:T], v_last.unsqueeze(0)), dim=0)
        return gae4(self.reward4[:T], self.v[:T], v_next, self.done[:T], gamma, lam)
