#!/usr/bin/env python
"""Benchmark of the hot path: batched rodent env steps/s (BASELINE.json metric, configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

b200 arm: `rodent_0.xml`, 4096 envs per GPU, random actions U(-1,1) redrawn every control step (pre-generated,
resident in HBM), one "step" = Rodent.step for the whole batch = 10 physics substeps + reward / done / obs +
the fused Brax episode / auto-reset wrappers.  Timed with CUDA events around every step on the launching
stream, L2 flushed between steps (outside the events), max over ranks.  `e2e` is the same step through
rr_env_step_host with pinned HOST action / obs / reward / done buffers (copies inside the timed region).
reference arm: the CPU oracle (oracle/, a C restatement of the MJX step -- mujoco / mjx / brax are not
installable in this image) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
N_FRAMES = 10
ALGO_BYTES_PER_ENV_STEP = 7204  # state in/out + action + 1263-float obs + scalars
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4; the bench line reports the MEASURED peak (rr_measure_fp32_peak)
# dram__bytes_read.sum + dram__bytes_write.sum of one rr_step_kernel launch (profiles/r02_kpar_ncu_raw.csv, `ncu --set full`):
# 5.25 MB + 0.26 MB.  Below the algorithmic 29.5 MB because the 25 MB of state / observation written by a launch is still
# resident in the 126 MB L2 when the kernel ends.
NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH = 5_245_184 + 259_584
# FP32 operations per env-step counted by ncu on the same launch (smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma x 2}_pred_on,
# profiles/r02_kpar_ncu_raw.csv): 1.0912e10 per 4096-env launch.  What the kernel EXECUTES (tree-sparse, active rows only); the
# roofline numerator is the oracle's operation count of the reference's dense formulation with non-zero operands (below).
NCU_FLOPS_PER_ENV_STEP = 1.0912e10 / 4096


def synthetic_track(n=250):
    return np.stack([0.002 * np.arange(n), np.zeros(n), np.full(n, 0.055)], 1).astype(np.float32)


# ------------------------------------------------------------------------------------------------ CPU oracle arm
_W = {}


def _oracle_init(iterations, ls_iterations, model="rodent_0"):
    """Pool initializer: one oracle environment per worker process (kept alive across bench steps)."""
    from brax_rodent_run_b200 import mjcf, model_blob
    from oracle import oracle
    m = mjcf.FlatModel.load(os.path.join(ROOT, "brax_rodent_run_b200", "assets", f"{model}.npz"))
    blob = model_blob.pack(m)
    env = oracle.OracleRodentEnv(blob, (m.nq, m.nv, m.nu, m.nbody), synthetic_track(), iterations=iterations,
                                 ls_iterations=ls_iterations, precision="f32")
    rng = np.random.default_rng(os.getpid())
    _W.update(env=env, m=m, rng=rng, steps=0)
    _oracle_reset()


def _oracle_reset():
    m, rng = _W["m"], _W["rng"]
    q = m.qpos0.copy()
    sf = int(rng.integers(0, 100))
    q[:3] = synthetic_track()[sf]
    _W["env"].reset(sf, q + rng.uniform(-.01, .01, m.nq), rng.uniform(-.01, .01, m.nv))
    _W["steps"] = 0


def _oracle_steps(n_steps):
    env, m, rng = _W["env"], _W["m"], _W["rng"]
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, done, _ = env.step(rng.uniform(-1, 1, m.nu))
        _W["steps"] += 1
        if done or _W["steps"] >= 1000:  # terminate_when_unhealthy / episode_length, as the B200 arm's fused wrappers
            _oracle_reset()
    return time.perf_counter() - t0


def oracle_flops_per_env_step(iterations, ls_iterations, model="rodent_0", envs=2, steps=12, settle=8):
    """Operation count of one env step from the instrumented oracle (oracle/rr_oracle_count.cpp): (total, useful) FP operations of
    the reference's dense formulation, averaged over a bounded random-action sample (`envs` x `steps` after `settle` steps).
    'useful' drops operations on structural zeros (a product with a zero operand, a sum of two zeros): the sparsity-exact count
    SURVEY 8(d) asks for."""
    from brax_rodent_run_b200 import mjcf, model_blob
    from oracle import oracle
    oracle.build()
    m = mjcf.FlatModel.load(os.path.join(ROOT, "brax_rodent_run_b200", "assets", f"{model}.npz"))
    blob = model_blob.pack(m)
    tot = use = n = 0
    for e in range(envs):
        rng = np.random.default_rng(100 + e)
        env = oracle.OracleRodentEnv(blob, (m.nq, m.nv, m.nu, m.nbody), synthetic_track(), iterations=iterations,
                                     ls_iterations=ls_iterations, precision="cnt")
        q = m.qpos0.copy()
        sf = int(rng.integers(0, 100))
        q[:3] = synthetic_track()[sf]
        env.reset(sf, q + rng.uniform(-.01, .01, m.nq), rng.uniform(-.01, .01, m.nv))
        for t in range(settle + steps):
            if t == settle:
                env.o.ops(reset=True)
            env.step(rng.uniform(-1, 1, m.nu))
        a, b = env.o.ops(reset=True)
        tot, use, n = tot + a, use + b, n + steps
    return tot / n, use / n


class OraclePool:
    """One oracle environment per host core; `rate(n)` steps every environment n times and returns env-steps/s with the
    slowest worker bounding the time (process start-up and model load excluded)."""

    def __init__(self, iterations, ls_iterations, cores=None, model="rodent_0"):
        from oracle import oracle
        oracle.build()
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_oracle_init,
                                                  initargs=(iterations, ls_iterations, model))

    def rate(self, n_steps):
        times = self.pool.map(_oracle_steps, [n_steps] * self.cores, chunksize=1)
        return self.cores * n_steps / max(times), max(times)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 25  # env steps per core per bench "step": a bounded sample of the 4096-env workload (~0.15 s of CPU per core)
    pool = OraclePool(args.iterations, args.ls_iterations, model=args.model)
    for _ in range(args.warmup):
        pool.rate(n)
    total_t, t0 = 0.0, time.perf_counter()
    for _ in range(args.steps):
        _, t = pool.rate(n)
        total_t += t
    wall = time.perf_counter() - t0
    cores = pool.cores
    pool.close()
    value = cores * n * args.steps / total_t
    sample = (f"{cores} envs (one per host core) x {n} env steps per bench step, {args.model}.xml, oracle fp32 C restatement of the "
              "MJX step")
    print(json.dumps({
        "impl": "reference", "metric": "rodent env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
        "note": "mujoco / mujoco-mjx / brax / jax are absent from this image (no wheels, no network): the reference arm "
                "is the repo's C restatement of the MJX step (oracle/rr_oracle.c) on all host cores",
    }))


def workload_config(args):
    return {"workload": f"{args.model}.xml run task, Rodent.step (10 substeps of 2 ms, CG solver) with random actions, "
                        f"{args.envs} envs/GPU", "envs_per_gpu": args.envs, "n_frames": N_FRAMES, "solver": "cg",
            "iterations": args.iterations, "ls_iterations": args.ls_iterations, "episode_length": 1000,
            "terminate_when_unhealthy": True, "l2": "flushed between timed steps (256 MiB write outside the event pair)",
            "includes": "physics + reward/done/metrics + observation + fused Episode/AutoReset wrappers"}


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from brax_rodent_run_b200 import _lib
    from brax_rodent_run_b200.env import Rodent

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K, W = args.envs, args.steps, args.warmup
    env = Rodent(synthetic_track(), num_envs=B, device=dev, model=args.model, solver="cg", iterations=args.iterations,
                 ls_iterations=args.ls_iterations, terminate_when_unhealthy=True, kinematics_outputs=False,
                 balance=args.balance, _lib_path=args.lib)
    env.wrap_for_training(episode_length=1000)
    L = env._L
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    state = env.reset(gen)
    actions = torch.rand((K + W, B, env.action_size), generator=gen, device=dev) * 2 - 1  # resident in HBM
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        state = env.step(state, actions[i])
    barrier()

    # ---- device-resident timing: per-step event pairs, L2 flushed in between -------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = L.rr_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(float(i))
        ev[i][0].record()
        state = env.step(state, actions[W + i])
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = L.rr_launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)
    done_frac = float(state.done.mean().item())

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------------------------
    Ke = min(K, 50)
    h_act = torch.empty((B, env.action_size), pin_memory=True)
    h_obs = torch.empty((B, env.observation_size), pin_memory=True)
    h_rew = torch.empty((B,), pin_memory=True)
    h_done = torch.empty((B,), pin_memory=True)
    host_actions = (torch.rand((Ke + 2, B, env.action_size)) * 2 - 1)
    buf, tens = env._out_buffers()
    ps = state.pipeline_state
    for k_, v_ in (("qpos", ps.qpos), ("qvel", ps.qvel), ("act", ps.act), ("qacc_warmstart", ps.qacc_warmstart), ("time", ps.time),
                   ("cur_frame", state.info["cur_frame"]), ("done", state.done), ("steps", state.info["steps"])):
        tens[k_].copy_(v_)
    f = state.info["first_pipeline_state"]
    buf.first_qpos, buf.first_qvel, buf.first_act = f.qpos.data_ptr(), f.qvel.data_ptr(), f.act.data_ptr()
    buf.first_qacc_warmstart, buf.first_time, buf.first_obs = f.qacc_warmstart.data_ptr(), f.time.data_ptr(), state.info["first_obs"].data_ptr()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def host_step(i):
        h_act.copy_(host_actions[i])
        _lib.check(L, L.rr_env_step_host(env._env, ctypes.byref(buf), ctypes.c_void_p(h_act.data_ptr()), N_FRAMES,
                                         ctypes.c_void_p(h_obs.data_ptr()), ctypes.c_void_p(h_rew.data_ptr()),
                                         ctypes.c_void_p(h_done.data_ptr()), stream))

    host_step(0); host_step(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        host_step(2 + i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Ke / float(t.item())
    clocks = sampler.stop() if sampler else None  # sampled every 20 ms over the device-timed loop and the end-to-end loop
    fp32_peak = ctypes.c_double(0.0)
    _lib.check(L, L.rr_measure_fp32_peak(ctypes.byref(fp32_peak), stream))
    extra = {}
    if not args.no_extra:
        extra = measure_extra(args, dev, world, rank)
    h2d = B * env.action_size * 4
    d2h = B * (env.observation_size + 2) * 4

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        kernel_ms = total_ms / K  # one rr_step_kernel launch per step; the event pair brackets only that launch
        d_ = env.dims  # state in + action in, state out + observation + 8 scalars out
        algo_bytes = 4 * (2 * (d_.nq + 2 * d_.nv + d_.na) + env.action_size + d_.obs_dim + 8)
        assert args.model != "rodent_0" or algo_bytes == ALGO_BYTES_PER_ENV_STEP
        achieved_gbs = algo_bytes * B / (kernel_ms * 1e-3) / 1e9
        kernel_name = "rr_step_kernel<3>" if d_.nv <= 96 else "rr_step_kernel<5>"
        # FP32 roofline (the binding roof): operations per env-step from the instrumented oracle on a bounded
        # sample of the same workload, divided by the FMA throughput measured on this device just now
        flops_total, flops_useful = oracle_flops_per_env_step(args.iterations, args.ls_iterations, args.model)
        achieved_tflops = flops_useful * B / (kernel_ms * 1e-3) / 1e12
        peak_tf = float(fp32_peak.value)
        out = {
            "metric": "rodent env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tf if peak_tf > 0 else None,
                         "traffic": NCU_DRAM_TRAFFIC_BYTES_PER_LAUNCH if (args.model, B) == ("rodent_0", ENVS_PER_GPU) else None,
                         "kernel": kernel_name,
                         "peak_source": "measured in this run: rr_measure_fp32_peak (8 FMA chains per thread on every SM); nominal "
                                        f"{FP32_PEAK_TFLOPS_NOMINAL:.1f}",
                         "flops_per_env_step": flops_useful, "flops_per_env_step_dense": flops_total,
                         "flops_per_env_step_ncu": NCU_FLOPS_PER_ENV_STEP if args.model == "rodent_0" else None,
                         "flops_source": "oracle/rr_oracle_count.cpp: FP operations with non-zero operands of the reference's dense "
                                         "formulation, 2 envs x 12 random-action env steps; _dense counts every operation; _ncu is "
                                         "what the kernel executes (tree-sparse, active rows only; profiles/r02_kpar_ncu_raw.csv)",
                         "note": "irregular 73-wide tree arithmetic: bound by FP32 issue / dependent-chain latency, not by HBM "
                                 "(7.2 KB per env-step); see roofline_hbm"},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                             "algorithmic_bytes": algo_bytes * B, "peak_source": peak_src},
            "wall_s": wall, "done_frac_last_step": done_frac,
        }
        if args.lib:
            out["lib_override"] = args.lib
        if extra:
            out["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            pool = OraclePool(args.iterations, args.ls_iterations, model=args.model)
            pool.rate(5)
            rate, _ = pool.rate(150)
            cores = pool.cores
            pool.close()
            out["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"{cores} envs (one per host core) x 150 env steps, {args.model}.xml, oracle fp32 (C "
                                             "restatement of the MJX step; mujoco / mjx are not installable here)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ extras on the bench line
def measure_extra(args, dev, world, rank):
    """The other half of BASELINE.json's metric and configs[4], measured after the timed region of the contract line and
    appended to the same JSON line: PPO train SPS (2048 envs / GPU, README configuration, NCCL gradient all-reduce when
    world > 1) and rodent_pair env-steps/s (4096 envs / GPU).  Bounded: 1 + 3 training steps, 3 + 10 env steps."""
    import torch
    import torch.distributed as dist
    from brax_rodent_run_b200 import ppo
    from brax_rodent_run_b200.env import Rodent
    out = {}
    try:
        env = Rodent(synthetic_track(), num_envs=2048, device=dev, model="rodent_0", iterations=args.iterations,
                     ls_iterations=args.ls_iterations, terminate_when_unhealthy=False, kinematics_outputs=False)
        cfg = ppo.PPOConfig(num_envs=2048)
        agent = ppo.PPO(env.wrap_for_training(cfg.episode_length), cfg)
        state = env.reset(rank)
        state, _ = agent.training_step(state)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, t0 = agent.env_steps, time.perf_counter()
        for _ in range(3):
            state, _ = agent.training_step(state)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        update_us = None
        if getattr(agent, "_graph", None) is not None:   # device time of one captured minibatch update (replays of the graph)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(50):
                agent._graph.replay()
            ev1.record()
            torch.cuda.synchronize(dev)
            update_us = ev0.elapsed_time(ev1) / 50 * 1e3
        out["ppo_train_sps"] = {"value": (agent.env_steps - e0) / float(dt.item()), "unit": "env-steps/s", "envs_per_gpu": 2048,
                                "update_graph_us": update_us,
                                "n_gpus": world, "training_steps": 3,
                                "learner": ("tcgen05: %d grouped TMA + tcgen05.mma TF32 GEMM launches per minibatch update, own Adam / gather / loss "
                                            "kernels (DESIGN.md section 11)" % agent._tc.launches_per_update) if agent._tc is not None
                                else "autograd + cuBLAS TF32",
                                "config": "readme.md:17-31: unroll 10, batch 512 x 64 minibatches, 8 epochs, CG 8/8, normalised obs"}
        del agent, env, state
    except Exception as ex:  # the contract line must survive a failure of an extra
        out["ppo_train_sps"] = {"error": repr(ex)[:200]}
    try:
        B = ENVS_PER_GPU
        env = Rodent(synthetic_track(), num_envs=B, device=dev, model="rodent_pair", iterations=args.iterations,
                     ls_iterations=args.ls_iterations, kinematics_outputs=False).wrap_for_training(1000)
        gen = torch.Generator(device=dev)
        gen.manual_seed(99 + rank)
        state = env.reset(gen)
        acts = torch.rand((13, B, env.action_size), generator=gen, device=dev) * 2 - 1
        for i in range(3):
            state = env.step(state, acts[i])
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            state = env.step(state, acts[3 + i])
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["rodent_pair"] = {"value": world * B * 10 / (float(t.item()) * 1e-3), "unit": "env-steps/s", "envs_per_gpu": B,
                              "n_gpus": world, "ms_per_step": float(t.item()) / 10, "steps": 10}
    except Exception as ex:
        out["rodent_pair"] = {"error": repr(ex)[:200]}
    return out


# ------------------------------------------------------------------------------------------------ PPO train SPS
def run_ppo(args):
    """Second half of BASELINE.json's metric: PPO train SPS on the README configuration (configs[2] / [3]): 2048 envs per GPU,
    unroll 10, batch 512 x 64 minibatches, 8 epochs, CG 8/8, normalised observations; `--steps` = training steps timed."""
    import torch
    import torch.distributed as dist
    from brax_rodent_run_b200 import ppo
    from brax_rodent_run_b200.env import Rodent
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 2048 if args.envs == ENVS_PER_GPU else args.envs
    env = Rodent(synthetic_track(), num_envs=B, device=dev, model=args.model, iterations=args.iterations,
                 ls_iterations=args.ls_iterations, terminate_when_unhealthy=False, kinematics_outputs=False)
    cfg = ppo.PPOConfig(num_envs=B)
    agent = ppo.PPO(env.wrap_for_training(cfg.episode_length), cfg)
    state = env.reset(rank)
    for _ in range(max(1, min(args.warmup, 2))):
        state, _ = agent.training_step(state)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    L = env._L
    e0, l0, t0 = agent.env_steps, L.rr_launch_count(), time.perf_counter()
    K = max(1, min(args.steps, 5))
    for _ in range(K):
        state, m = agent.training_step(state)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "PPO train SPS", "value": (agent.env_steps - e0) / float(dt.item()), "unit": "env-steps/s",
                          "n_gpus": world, "steps": K, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * float(dt.item()) / K,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TF32 matmuls in the MLPs)",
                          "data": "synthetic",
                          "config": {"workload": f"PPO on the {args.model}.xml run task, {B} envs/GPU, unroll 10, batch 512 x 64 "
                                                 "minibatches, 8 epochs, normalised observations (readme.md:17-31)",
                                     "iterations": args.iterations, "ls_iterations": args.ls_iterations,
                                     "env_steps_per_training_step": (agent.env_steps - e0) // K},
                          "gpu_launches": int(L.rr_launch_count() - l0),
                          "losses": {k: float(v) for k, v in m.items()}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--iterations", type=int, default=8)      # brax_rodent_run_ppo.py:52
    ap.add_argument("--ls-iterations", type=int, default=8)   # brax_rodent_run_ppo.py:53
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="rodent_0", help="rodent_0 (BASELINE configs[1], default), rodent_pair (configs[4]), rodent_new, ...")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--workload", default="step", choices=["step", "ppo"],
                    help="step: env-steps/s of Rodent.step (the contract line, default); ppo: PPO train SPS")
    ap.add_argument("--no-extra", action="store_true", help="skip the PPO-train-SPS / rodent_pair extras appended to the bench line")
    ap.add_argument("--lib", default=None, help="developer knob: A/B a kernel build (variants/librr_<name>.so); recorded in the line")
    ap.add_argument("--balance", action="store_true",
                    help="cost-sorted env -> CTA assignment (off: measured slower than the even contiguous split, see DESIGN.md)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "ppo":
        run_ppo(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
