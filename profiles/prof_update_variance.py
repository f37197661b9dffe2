"""Run-to-run variation of the PPO update: per-iteration collect / update times and the caching allocator's
segment counters.    python profiles/prof_update_variance.py [envs] [iterations]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
J, M, E, H = 6, 6, 2, 128
N = J * M
pkg = importlib.import_module("e2e-mappo-for-mt-fjsp_b200")
envm = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.env")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
rom = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rollout")
ppo = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.ppo")
d = pkg.instances.synthetic_instances(0, B, J, M, E, 1002)
env = envm.BatchedMTFJSPEnv(B, J, M, E, obs_dtype=torch.float32)
env.load(d["t"], d["p"], d["transT"], d["edge"])
env.scaler_init()
job = enc.JobActor(enc.seeded_state_dict(enc.job_actor_keys(H), 11), J, M, hidden=H, trainable=True)
mch = enc.MachineActor(enc.seeded_state_dict(enc.machine_actor_keys(H), 12), M, hidden=H, trainable=True)
crit = enc.GlobalCritic(enc.seeded_state_dict(enc.global_critic_keys(H), 13), J, M, hidden=H, trainable=True)
ro = rom.Rollout(env, job.inference_twin("tf32"), mch.inference_twin("tf32"), greedy=False, use_cuda_graph=True, seed=2)
up = ppo.MAPPOUpdate(job, mch, crit, ppo.PPOConfig(k_epochs=1, encoder_tf32=True))
w = [pkg.instances.random_weights(0, B, 100)]
for it in range(ITERS):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    bt = ppo.collect(ro, w)
    e[1].record()
    up.recompute_old_logp(bt)
    e[2].record()
    up.update(bt, N)
    e[3].record()
    torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    print("iter %d: collect %.1f ms, old log-probs %.1f ms, update %.1f ms | reserved %.1f GB, segments allocated so far %d, "
          "alloc retries %d" % (it, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]),
                                st["reserved_bytes.all.current"] / 2 ** 30, st["segment.all.allocated"], st["num_alloc_retries"]),
          flush=True)
    del bt
