#!/usr/bin/env python
"""bench.py -- MT-FJSP env-steps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the environment hot path over the whole env batch: random valid action ->
candidate-machine features -> transition + reward + reward scaling + observation + job mask, i.e.
everything the environment side of Run.py's rollout loop does per step,
one launch of the fused kernel for the sizes that have a specialised one.  Episodes wrap inside the timed
region: every N = J*M steps the batch is reset (two more launches), as the reference's loop does.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload A|B|C]

N > 1 is launched by the driver with torch.distributed.run; each rank owns its own env slice (weak
scaling, no data-path collective; NCCL is used only for the barrier and the max-over-ranks time).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1..3]
    "A": dict(name="J6M6E2 synthetic, 65536 envs/GPU, random-action rollouts", J=6, M=6, E=2, B=65536, seed=1002),
    "B": dict(name="J10M10E3 synthetic, 16384 envs/GPU, env step + mask/feature build", J=10, M=10, E=3, B=16384, seed=1003),
    "C": dict(name="J30M20E5 synthetic, 4096 envs/GPU", J=30, M=20, E=5, B=4096, seed=1004),
}
METRIC = "MT-FJSP env-steps/sec"
UNIT = "env-steps/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_rate(wl, nthreads, target_seconds):
    """Times the CPU restatement (oracle/) on a bounded sample of the same workload: whole random rollouts,
    every step = mask + candidates + mfea1 + transition + reward + scaling + observation (oracle_rollout_random)."""
    from oracle.mtfjsp_oracle import OracleEnv

    ins = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.instances")
    J, M, E = wl["J"], wl["M"], wl["E"]
    N = J * M
    probe_B = max(nthreads * 4, 64 if N <= 100 else 16)

    def run(B, episodes):
        d = ins.synthetic_instances(0, B, J, M, E, wl["seed"])
        w = ins.random_weights(0, B, wl["seed"])
        o = OracleEnv(B, J, M, E, left_shift=True, nthreads=nthreads)
        o.load(d["t"], d["p"], d["transT"], d["edge"])
        o.scaler_init()
        o.reset(w)
        o.rollout_random(N, seed=1)  # warm-up episode
        t0 = time.perf_counter()
        for ep in range(episodes):
            o.reset(w)
            o.scaler_reset()
            o.rollout_random(N, seed=2 + ep)
        return B * N * episodes / (time.perf_counter() - t0)

    rate = run(probe_B, 1)
    B = int(min(max(probe_B, rate * target_seconds / (N * 4)), 262144))
    B = max(nthreads, (B // nthreads) * nthreads)
    rate = run(B, 4)
    return rate, "%d envs x 4 episodes x %d steps (%s), %d thread(s)" % (B, N, wl["name"].split(",")[0], nthreads)


def run_reference(args, wl):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself is Python over
    networkx and cannot travel to the GPU box; this times its C restatement (oracle/, kind 'port') with all
    host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every host core this process may run on (torchrun exports OMP_NUM_THREADS=1; the oracle's parallel loops take an
    # explicit thread count, so that default does not starve the reference arm)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, cores)
    rates = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_rate(wl, cores, 1.0)
    t0 = time.perf_counter()
    sample = ""
    for _ in range(max(1, min(args.steps, 5))):
        r, sample = cpu_port_rate(wl, cores, 4.0)
        rates.append(r)
    rates.sort()
    val = rates[len(rates) // 2]
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["name"], "note": "CPU restatement of the reference env (oracle/, OpenMP over envs); "
                       "the Python reference itself measured ~157 env-steps/s/core at survey time (BASELINE.md)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def run_ours(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stdout_fd = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line only: NCCL prints its version banner with printf when NCCL_DEBUG=VERSION, so
        # everything the libraries write to fd 1 goes to stderr until the line is printed
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
    sh = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.sharding")
    J, M, E, B = wl["J"], wl["M"], wl["E"], args.batch or wl["B"]
    if args.scaling == "strong":  # fixed total batch: the workload's env count is split over the ranks
        B = max(1, B // world)
    N = J * M
    # weak scaling: B envs per GPU, env i of the B*world job lives on rank i // B (contiguous slices)
    first, count, d, w = sh.make_shard(B * world, rank, world, J, M, E, wl["seed"])
    assert count == B and first == rank * B
    w = torch.as_tensor(w).to(dev)
    env = envm.BatchedMTFJSPEnv(B, J, M, E, left_shift=True, obs_dtype=torch.float32, mask_mode=envm.MASK_ESA)
    env.load(d["t"], d["p"], d["transT"], d["edge"])
    env.scaler_init()
    env.reset(w)
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"s": 0}

    def one_step(seed=1234):
        if state["s"] == N:  # episode finished for every env: next episode (Run.py:615-665)
            env.reset(w)
            env.scaler_reset()
            state["s"] = 0
        env.random_step(seed=seed, env_offset=first)
        state["s"] += 1

    for _ in range(max(W, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count - l0
    ms = sh.max_over_ranks(ms, dev)
    value = B * world * K / (ms * 1e-3)

    # ---- dominant kernel alone: fused step+obs launches replaying recorded actions (device resident) ----
    env.reset(w); env.scaler_reset()
    rec_op = torch.empty((N, B), dtype=torch.int32, device=dev)
    rec_mc = torch.empty((N, B), dtype=torch.int32, device=dev)
    for s in range(N):
        env.random_step(seed=99, env_offset=first)
        rec_op[s].copy_(env.op); rec_mc[s].copy_(env.mach)
    assert int(env.done.sum().item()) == B and int(env.invalid.sum().item()) == 0
    kms, klaunch = 0.0, 0
    reps = max(1, min(4, K // N))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
    for rep in range(reps + 1):
        env.reset(w); env.scaler_reset()
        for s in range(N):
            evs[s][0].record()
            env.step_obs(rec_op[s], rec_mc[s])
            evs[s][1].record()
        torch.cuda.synchronize()
        if rep > 0:  # first replay is warm-up
            kms += sum(a.elapsed_time(b) for a, b in evs)
            klaunch += N
    k_us = kms * 1e3 / klaunch
    bytes_step = env.bytes_per_step()
    peak, peak_src = measured_peak()
    achieved = bytes_step * B / (k_us * 1e-6) / 1e9
    # the kernel the timed region above actually launches: the one-launch random-rollout step (policy + candidate-machine
    # features + step + obs), one CUDA-event pair per launch on the launching stream
    fused = None
    if env.random_step_is_fused:
        fms = 0.0
        for rep in range(reps + 1):
            env.reset(w); env.scaler_reset()
            for s in range(N):
                evs[s][0].record()
                env.random_step(seed=500 + rep, env_offset=first)
                evs[s][1].record()
            torch.cuda.synchronize()
            if rep > 0:
                fms += sum(a.elapsed_time(b) for a, b in evs)
        f_us = fms * 1e3 / klaunch
        fbytes = env.bytes_per_random_step()
        fused = {"kernel_us": f_us, "bytes_per_env_step": fbytes, "achieved": fbytes * B / (f_us * 1e-6) / 1e9}

    # ---- e2e: host-buffer C-ABI call per step (pinned H2D action pairs, D2H packed step records = step info + job
    # mask + candidates; mtfjsp_step_host_packed cuts the batch into chunks whose copies overlap the other chunks' kernels) ----
    h_act = torch.stack([rec_op.cpu(), rec_mc.cpu()], dim=2).contiguous().pin_memory()  # [N,B,2] (op, machine)
    _, h_rec = env.host_buffers()
    rec_view = h_rec.numpy().view(env.host_record_dtype())[:, 0]
    Ke = min(K, 4 * N)

    host_step = env.host_stepper(h_rec)                       # prepared call: buffers bound once
    act_ptr = [h_act[s].data_ptr() for s in range(N)]         # this step's pinned [B,2] action array

    def e2e_steps(n):
        s = 0
        env.reset(w); env.scaler_reset()
        for _ in range(n):
            if s == N:
                env.reset(w); env.scaler_reset(); s = 0
            host_step(act_ptr[s])
            s += 1

    e2e_steps(N + 3)  # one whole episode first: every action buffer's graph is instantiated outside the timed region
    barrier()
    t0 = time.perf_counter()
    e2e_steps(Ke)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = sh.max_over_ranks(e2e_s, dev)
    e2e_val = B * world * Ke / e2e_s
    assert float(rec_view["info6"][:, 1].sum()) in (0.0, float(B))
    h_op = rec_op.cpu(); h_mc = rec_mc.cpu()
    # ---- BASELINE.json configs[4] (forward half): the same env slice driven by the MAPPO actors on the device ----
    policy = None
    if not args.no_policy and (J, M) == (6, 6):
        enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
        rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
        job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M, precision="tf32")
        mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(128), 12), M, precision="tf32")
        ro = rom.Rollout(env, job, mch, greedy=False, use_cuda_graph=True, seed=1)
        ro.begin_episode(w)
        for _ in range(4):
            ro.step()  # eager first step, warm-up, capture
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nstep = N - 4
        p0.record()
        for _ in range(nstep):
            ro.step()
        p1.record()
        barrier()
        pms = sh.max_over_ranks(p0.elapsed_time(p1), dev)
        assert int(env.done.sum().item()) == B and int(env.invalid.sum().item()) == 0
        policy = {"value": B * world * nstep / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms / nstep, "steps": nstep,
                  "what": "job actor (GIN encoder, tcgen05 TF32 layers) + machine actor (GAT) + sampling + env step/obs, "
                          "CUDA-graph replay, random-init weights", "hidden": 128}
    stats = sh.reduce_episode_stats(env.costs(), device=dev)  # the rollout side's only other exchange (6 doubles)

    # ---- BASELINE.json configs[4]: env slice + encoder + PPO update end to end, gradient allreduce share ----
    train = None
    if not args.no_train and (J, M) == (6, 6):
        enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
        rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
        ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")
        Bt = min(B, 8192)  # 65,536 envs over 8 GPUs
        env_t = envm.BatchedMTFJSPEnv(Bt, J, M, E, left_shift=True, obs_dtype=torch.float32)
        env_t.load(d["t"][:Bt], d["p"][:Bt], d["transT"][:Bt], d["edge"][:Bt])
        env_t.scaler_init()
        tj = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(128), 11), J, M, trainable=True)
        tm = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(128), 12), M, trainable=True)
        tc = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(128), 13), J, M, trainable=True)
        tro = rom.Rollout(env_t, tj.inference_twin("tf32"), tm.inference_twin("tf32"), greedy=False, use_cuda_graph=True,
                          seed=2 + rank)
        wt = [w[:Bt]]
        runs = {}
        for variant, enc_tf32 in (("tcgen05_tf32", True), ("library_fp32", False)):
            up = ppo.MAPPOUpdate(tj, tm, tc, ppo.PPOConfig(k_epochs=1, encoder_tf32=enc_tf32))
            tms, cms, ams = [], [], []
            for it in range(3):  # first iteration is warm-up (library initialisation)
                barrier()
                t0e, t1e, t2e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                t0e.record()
                btc = ppo.collect(tro, wt)
                t1e.record()
                losses, _ = up.update(btc, N)
                tro.job.refresh(); tro.mch.refresh()
                t2e.record()
                barrier()
                a_ms = up.allreduce_ms()
                if it > 0:
                    cms.append(sh.max_over_ranks(t0e.elapsed_time(t1e), dev)); tms.append(sh.max_over_ranks(t0e.elapsed_time(t2e), dev))
                    ams.append(a_ms)
                del btc
            runs[variant] = (sum(tms) / len(tms), sum(cms) / len(cms), sum(ams) / len(ams), [float(x) for x in losses],
                             up.allreduce_bytes // 3)
        tot, col, arm, losses, arb = runs["tcgen05_tf32"]
        train = {"value": Bt * world * N / (tot * 1e-3), "unit": UNIT, "envs_per_gpu": Bt, "buffer_steps": N, "k_epochs": 1,
                 "mini_bs": N, "collect_ms": col, "update_ms": tot - col,
                 "allreduce_ms": arm, "allreduce_share": arm / tot,
                 "allreduce_bytes_per_update": arb, "losses": losses,
                 "update_ms_library_fp32": runs["library_fp32"][0] - runs["library_fp32"][1],
                 "value_library_fp32": Bt * world * N / (runs["library_fp32"][0] * 1e-3),
                 "what": "one buffer (1 episode) collected with the tcgen05 rollout twins (CUDA-graph replay) + one batched PPO "
                         "update; every [rows,128] x [128,<=128] product of the update (graph encoders, GAT projections, policy "
                         "heads) runs forward / input-gradient / weight-gradient on the hand-written tcgen05 TF32 kernels "
                         "(PPOConfig.encoder_tf32); aggregation, grouped BatchNorm and GAE kernels hand-written; NCCL gradient "
                         "allreduce when n_gpus > 1.  *_library_fp32 = the same update with every GEMM on the FP32 library path "
                         "(the reference's arithmetic)"}
        del env_t, up, tro
        torch.cuda.empty_cache()

    clocks = sampler.stop() if sampler else None
    traffic = traffic_step = None
    tpath = os.path.join(ROOT, "profiles", "r01_env_kernel_traffic.json")
    if (J, M, B) == (6, 6, 65536) and os.path.exists(tpath):  # ncu --set full captures of these kernels on this workload
        tj = json.load(open(tpath))
        traffic_step = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        if "fused_random_step" in tj:
            traffic = tj["fused_random_step"]["dram_bytes_read"] + tj["fused_random_step"]["dram_bytes_write"]
    if fused is None:
        traffic = traffic_step

    # ---- the strict drop-in call: Parallel_env.DGFJSPEnv_paral_step with the reference's argument / return types
    # (python list of action pairs in, numpy float64 dense adjacency [B,N,N] + features + python info list out) ----
    dropin = None
    if rank == 0 and world == 1 and not args.no_dropin:
        pem = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.parallel_env")
        Bd = min(B, 8192)
        pe = pem.Parallel_env({"n_job": J, "n_machine": M, "n_edge": E, "env_batch": Bd, "GAMMA": 0.99,
                               "reward_scaling": {"scaling_divisor": 1}, "weight_mk": 0.4, "weight_ec": 0.4, "weight_tt": 0.2})
        pe.get_batch({k2: torch.as_tensor(d[k1][:Bd]) for k1, k2 in (("t", "t"), ("p", "p"), ("transT", "transT"), ("edge", "edge"))})
        pe.init_RewardScaling_sameBATCH(shape=4)
        pe.init_DGFJSPEnv_state0(weights=w[:Bd].cpu().numpy())
        acts = [list(zip(h_op[s2][:Bd].tolist(), h_mc[s2][:Bd].tolist())) for s2 in range(6)]
        pe.DGFJSPEnv_paral_step(acts[0])
        t0 = time.perf_counter()
        for s2 in range(1, 6):
            pe.DGFJSPEnv_paral_step(acts[s2])
        dt = time.perf_counter() - t0
        dropin = {"value": Bd * 5 / dt, "unit": UNIT, "envs": Bd, "d2h_bytes_per_step": Bd * (N * N * 8 + N * 96 + M * 64 + 80),
                  "api": "Parallel_env.DGFJSPEnv_paral_step (reference types: python action list in, numpy f64 dense "
                         "adjacency + features + python info list out)"}
        del pe

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r1, sample = cpu_port_rate(wl, 1, 8.0)
        cpu = {"value": r1, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "envs_per_gpu": B, "jobs": J, "machines": M, "edges": E, "obs_dtype": "f32",
                       "mask_mode": "ESA", "left_shift": True, "parallelism": "env-slices x%d (no data-path collective)" % world,
                       "l2": "working set %.0f MB per GPU > 126 MB L2 (inputs larger than L2)" % (B * (bytes_step + 2000) / 1e6),
                       "launches_per_step": round(launches / K, 3),
                       "kernels": "env_kernel_s<STEP|OBS|POLICY> (random policy + candidate-machine features + step + reward + "
                                  "obs + mask in one launch; reset launches env_kernel<RESET> + scaler_kernel every %d steps)" % N},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": B * h_rec.shape[1],
                    "api": "mtfjsp_step_host_packed (C ABI, pinned host buffers: [B,2] i32 action pairs in, [B] packed records "
                           "(f64 info6, i16 candidates, u8 job mask) out; observation tensors stay on the device)",
                    "steps": Ke},
            "gpu_launches": launches,
            "roofline": ({"bound": "hbm", "achieved": fused["achieved"], "peak": peak, "unit": "GB/s",
                          "frac": fused["achieved"] / peak, "traffic": traffic,
                          "kernel": "env_kernel_s<STEP|OBS|POLICY,float> (random policy + candidate-machine features + step + "
                                    "reward + observation + job mask; the kernel of the timed region)",
                          "kernel_us": fused["kernel_us"], "bytes_per_env_step": fused["bytes_per_env_step"],
                          "peak_source": peak_src, "steps_per_s_kernel_only": B / (fused["kernel_us"] * 1e-6),
                          "bytes_accounting": "SURVEY.md 8(d): a full rewrite of the observation every step, as the reference "
                                              "does; with the incremental observation the kernel moves `traffic` bytes per "
                                              "launch (ncu), i.e. frac_of_peak_moved = traffic / kernel time / peak",
                          "frac_of_peak_moved": (traffic / (fused["kernel_us"] * 1e-6) / 1e9 / peak) if traffic else None,
                          "step_obs_kernel": {"kernel": "env_kernel_s<STEP|OBS,float> (actions given: the actor-driven and "
                                                        "host-step paths)", "kernel_us": k_us, "bytes_per_env_step": bytes_step,
                                              "achieved": achieved, "frac": achieved / peak, "traffic": traffic_step}}
                         if fused else
                         {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "traffic": traffic, "kernel": "env_kernel<STEP|OBS,float> (fused step + reward + observation + job mask)",
                          "kernel_us": k_us, "bytes_per_env_step": bytes_step, "peak_source": peak_src,
                          "steps_per_s_kernel_only": B / (k_us * 1e-6)}),
            "clocks": clocks,
            "episode_stats": {k: float(v) for k, v in stats.items()},
        }
        if policy:
            line["policy_rollout"] = policy
        if train:
            line["train_iteration"] = train
        if dropin:
            line["e2e"]["dropin_parallel_env"] = dropin
        if cpu:
            line["cpu_baseline"] = cpu
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        if stdout_fd is not None:
            os.dup2(2, 1)  # teardown chatter stays off stdout as well
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3600)
    ap.add_argument("--warmup", type=int, default=36)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, what the driver measures): the workload's env count per GPU; strong: that count in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-policy", action="store_true", help="skip the actor-driven rollout measurement")
    ap.add_argument("--no-dropin", action="store_true", help="skip the Parallel_env (reference-typed) call measurement")
    ap.add_argument("--no-train", action="store_true", help="skip the rollout + PPO update measurement")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
