#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched KManip env step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--env KManipSoloArmQPos --envs 4096 --dtype float32]

A "step" is one env step (10 physics sub-steps + action decode + obs/reward) of every env of the batch: one kernel
launch.  Default workload = BASELINE.json configs[1]: KManipSoloArm scene, 4096 batched envs per GPU, joint-position
actions (KManipSoloArmQPos), random actions, autoreset every 64 steps.  Envs shard across GPUs with no per-step
collective (weak scaling: 4096 envs per GPU); one NCCL all-reduce of the episode statistics closes the rollout.

Prints ONE JSON line (rank 0).  `value` = device-timed, inputs resident in HBM; `e2e` = the same metric through the
host-buffer C-ABI call (km_step_host: pinned host actions in, obs/reward/truncated out, copies inside the timed
region).  `--impl reference` times the CPU oracle port (the reference's own MuJoCo path cannot run in this image,
see BASELINE.md) on all host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (whole box, device-timed)"
UNIT = "env-steps/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=512)
    p.add_argument("--warmup", type=int, default=8)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--env", default="KManipSoloArmQPos")
    p.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    p.add_argument("--dtype", default="float32")
    p.add_argument("--lanes", type=int, default=0)
    p.add_argument("--epb", type=int, default=0)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--ik-mode", type=int, default=0, help="0: fast fixed-iteration LM IK (default of the batched path); 1: exact-parity scipy-TRF restatement")
    p.add_argument("--no-extra", action="store_true", help="skip the other BASELINE.json configurations (extra_configs)")
    return p.parse_args()


def config_of(a):
    """`config` of the JSON line: identical for our arm and the reference arm (same workload, same batch)."""
    return {"workload": workload_name(a), "envs_per_gpu": a.envs, "sub_steps_per_env_step": 10}


# the other BASELINE.json configurations, measured on one GPU after the headline (env id, envs, dtype, BASELINE config)
EXTRA_CONFIGS = [
    ("KManipSoloArm", 8192, "float32", "configs[2]: SoloArm + IK + cube contacts, 65536 envs across 8 GPUs = 8192 per GPU"),
    ("KManipSoloArm", 65536, "float32", "configs[2] total batch on one GPU"),
    ("KManipDualArm", 32768, "float32", "configs[3]: DualArm + 2 IK solves, 32768 envs per GPU"),
    ("KManipTorso", 16384, "float64", "configs[4]: Torso, fp64 validation build, 16384 envs per GPU"),
    ("KManipSoloArmQPos", 65536, "float32", "throughput batch of the headline workload"),
]
NOMINAL_FMA_TFLOPS = {"float32": 74.4, "float64": 37.2}    # 148 SMs x 128 (64) lanes x 2 x 1.965 GHz


def workload_name(a):
    ik = ", exact-parity TRF IK" if getattr(a, "ik_mode", 0) == 1 else ""
    return f"{a.env}, {a.envs} envs/GPU, random actions, autoreset every 64 steps, {a.dtype}{ik}"


# ------------------------------------------------------------------------------------------------ CPU legs (oracle port)
def cpu_leg(env_id: str, n: int, budget_s: float, threads: int):
    """Oracle port timed on `threads` host threads on a bounded sample of the workload; returns (env-steps/s, sample)."""
    import numpy as np
    from oracle import oracle as om
    om.build()
    o = om.Oracle(env_id)
    st = om.batch_reset_state(o, n, seed=0)
    if o.nmocap == 0:
        st["mocap"] = np.zeros((n, 0))
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (16, n, o.task.act_dim)).astype(np.float32)
    for i in range(2):
        om.batch_step(o, st, acts[i], autoreset=True, seed=0, nthreads=threads)
    t0 = time.perf_counter()
    steps = 0
    while True:
        om.batch_step(o, st, acts[steps % 16], autoreset=True, seed=0, nthreads=threads)
        steps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or steps >= 1000:
            break
    return n * steps / dt, f"{n} envs x {steps} env-steps of the same workload, {threads} OpenMP threads, {dt:.1f} s"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_all = time.perf_counter()
    vals = []
    sample = ""
    per = max(1.0, min(a.cpu_seconds, 60.0) / max(a.steps + a.warmup, 1))
    # each "step" of the reference arm is a bounded sample; keep the whole run within a few minutes
    k = max(1, min(a.steps, 8))
    w = max(0, min(a.warmup, 2))
    for i in range(w + k):
        v, sample = cpu_leg(a.env, a.envs, max(1.0, a.cpu_seconds / (w + k)), cores)
        if i >= w:
            vals.append(v)
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": a.envs / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(a),
        "timing": "host wall clock (time.perf_counter) around the CPU loop: this arm has no device",
        "note": "CPU oracle port of the reference path on all host threads; real MuJoCo is not installable here (BASELINE.md)",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ our arm
def fma_peak(local, f32):
    """FMA throughput of the CUDA-core pipe measured on this GPU right now (km_measure_fma_peak), TFLOP/s, or None."""
    import ctypes as C
    from gym_kmanip_b200 import _lib
    pk = C.c_double(0)
    rc = _lib.load().km_measure_fma_peak(local, 32 if f32 else 64, C.byref(pk))
    return pk.value if rc == 0 and pk.value > 0 else None


def fp_roofline(env, dtype, n, ms_per_step, measured_peak, flops):
    """FP-pipe roofline of one configuration: counted algorithmic FLOPs / launch time against the measured FMA peak of
    this run AND the nominal peak (74.4 TFLOP/s fp32, 37.2 fp64 at the 1965 MHz maximum SM clock)."""
    fl = flops.get(env)
    if not fl:
        return None
    ach = fl * n / (ms_per_step * 1e-3) / 1e12
    nominal = NOMINAL_FMA_TFLOPS[dtype]
    return {"bound": "fp32" if dtype == "float32" else "fp64", "achieved": ach, "peak": measured_peak or nominal, "unit": "TFLOP/s",
            "frac": ach / (measured_peak or nominal), "frac_of_nominal": ach / nominal, "peak_nominal": nominal,
            "flop_per_env_step": fl,
            "peak_source": "measured (km_measure_fma_peak, dependent-FMA chains, this run)" if measured_peak else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
            "flop_source": "counted by the oracle's operation-counting build (profiles/flops_per_env_step.json)"}


def time_steps(sim, acts, steps, flush, torch):
    """CUDA-event time of `steps` launches (L2 flushed between them, outside the event pairs); returns total ms."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 255)
        evs[i][0].record()
        sim.step(acts[i % acts.shape[0]], autoreset=True, contacts=False)
        evs[i][1].record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs)


def run_extra_configs(local, flush, flops, peaks, torch):
    """The other BASELINE.json configurations on this GPU, a few steps each from step 24 of an episode (cube landed), so
    that the driver-run JSON line -- not only profiles/ -- carries them."""
    from gym_kmanip_b200.batch_sim import BatchSim
    out = []
    for env, n, dtype, what in EXTRA_CONFIGS:
        try:
            sim = BatchSim(env, n, device=local, dtype=dtype, seed=0)
            sim.reset()
            gen = torch.Generator(device=sim.device).manual_seed(99)
            acts = torch.rand(8, n, sim.act_dim, device=sim.device, generator=gen) * 2 - 1
            for i in range(24):
                sim.step(acts[i % 8], autoreset=True, contacts=False)
            torch.cuda.synchronize()
            k = 8
            ms = time_steps(sim, acts, k, flush, torch) / k
            out.append({"baseline_config": what, "env": env, "envs": n, "dtype": dtype, "steps": k, "ms_per_step": ms,
                        "value": n / (ms * 1e-3), "unit": UNIT, "launch": sim.launch_config(),
                        "roofline_fp": fp_roofline(env, dtype, n, ms, peaks.get(dtype), flops)})
            sim.close()
            del sim, acts
        except Exception as e:   # a configuration that cannot run must not take the headline line with it
            out.append({"baseline_config": what, "env": env, "envs": n, "dtype": dtype, "error": repr(e)})
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from gym_kmanip_b200.batch_sim import BatchSim

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = a.envs
    sim = BatchSim(a.env, n, device=local, dtype=a.dtype, seed=0, env0=rank * n, ik_mode=a.ik_mode)
    if a.lanes or a.epb:
        sim.configure(a.lanes, a.epb)
    cfg = sim.launch_config()
    sim.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    nact = 16
    acts = torch.rand(nact, n, sim.act_dim, device=dev, generator=gen) * 2 - 1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        sim.step(acts[i % nact], autoreset=True, contacts=False)
    sim.episode_stats(reset=True)          # the rollout totals are accumulated by the step kernel itself (km_episode_stats)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sim.launches
    t_wall = time.perf_counter()
    dev_ms = time_steps(sim, acts, a.steps, flush, torch)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = sim.launches - l0
    ret_sum = sim.episode_stats()          # [sum of rewards, env steps, truncations, success steps] of the timed steps
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(ret_sum, op=dist.ReduceOp.SUM)     # the single collective of the rollout
    dev_ms = float(tmax.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    value = world * n * a.steps / (dev_ms * 1e-3)

    # ---- e2e: the host-buffer C-ABI call, pinned host actions in, obs/reward/truncated out, copies timed
    h_act = (torch.rand(n, sim.act_dim) * 2 - 1).pin_memory()
    h_obs = torch.empty(n, sim.obs_dim, dtype=sim.tdtype).pin_memory()
    h_rew = torch.empty(n, dtype=sim.tdtype).pin_memory()
    h_tr = torch.empty(n, dtype=torch.uint8).pin_memory()
    ke = max(8, min(a.steps, 64))
    for _ in range(3):
        sim.step_host(h_act, h_obs, h_rew, h_tr, autoreset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(ke):
        sim.step_host(h_act, h_obs, h_rew, h_tr, autoreset=True)
    e1.record()
    barrier()
    e2e_s = max(time.perf_counter() - t0, e0.elapsed_time(e1) * 1e-3)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * n * ke / float(te.item())
    esize = 4 if sim.tdtype == torch.float32 else 8
    h2d = n * sim.act_dim * 4
    d2h = n * (sim.obs_dim * esize + esize + 1)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        # algorithmic HBM bytes per env-step: state record in + out, running return in + out, action in,
        # obs + reward + flag out
        alg_bytes = 2 * sim.state_dim * esize + 2 * 4 * 2 + 2 * esize + sim.act_dim * 4 + sim.obs_dim * esize + esize + 1
        ms_per_step = dev_ms / a.steps
        achieved = alg_bytes * n / (ms_per_step * 1e-3) / 1e9
        flops = {}
        try:
            flops = json.load(open(os.path.join(ROOT, "profiles", "flops_per_env_step.json")))
        except Exception:
            pass
        fma = {"float32": fma_peak(local, True), "float64": fma_peak(local, False)}
        traffic = flops.get("dram_traffic_bytes_per_launch", {}).get(f"{a.env}:{n}:{cfg['lanes_per_env']}")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if sim.tdtype == torch.float32 else "f64", "data": "synthetic",
            "config": config_of(a),
            "timing": "CUDA events on the launching stream around every launch; L2 flushed (256 MiB fill) between timed launches, outside the event pairs; max over ranks",
            "launch": cfg,
            "sub_steps_per_s": value * 10,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the same command, from profiles/ (not measured in this run)",
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": alg_bytes,
                         "note": "the path is FP-pipe/latency bound by construction (SURVEY.md 8d); see roofline_fp"},
            "roofline_fp": fp_roofline(a.env, a.dtype, n, ms_per_step, fma.get(a.dtype), flops),
            "episode_stats": {"mean_reward": float(ret_sum[0] / max(float(ret_sum[1]), 1.0)), "env_steps": float(ret_sum[1]),
                              "truncations": float(ret_sum[2]), "successes": float(ret_sum[3]),
                              "source": "accumulated by the step kernel (km_episode_stats)"},
            "wall_s_timed_region": t_wall,
        }
        sim.close()
        if world == 1 and not a.no_extra:
            line["extra_configs"] = run_extra_configs(local, flush, flops, fma, torch)
        if not a.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            v, sample = cpu_leg(a.env, n, a.cpu_seconds, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    else:
        sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
