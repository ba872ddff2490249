#!/usr/bin/env python3
"""Benchmark of the PGTG hot path (BASELINE.json: env-steps/sec, device-timed, vs CPU PGTGEnv; % HBM
roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one fused tick (traffic, agent move, reward/termination, same-step auto-reset with
on-device map regeneration, observation write) over every env of the rank's shard. Under torchrun
each rank owns `envs_per_gpu` envs (weak scaling, global env ids, no data-path collective); the only
collective is the episode-statistics all-reduce at the end of the timed region.

Prints ONE JSON line (rank 0): value = whole-job env-steps/s with inputs resident in HBM,
e2e = the same through `pgtg_step_host` with pinned HOST buffers (copies inside the timed region),
roofline = algorithmic bytes (SURVEY.md 8d formula) / measured kernel time vs MEASURED_PEAKS.json,
cpu_baseline = the CPU oracle port timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name -> (PGTGEnv kwargs, envs per GPU, cpu sample envs)
WORKLOADS = {
    # BASELINE config 5 per-GPU shard == the "default map settings" the north star quotes the target on
    "default-2M": (dict(), 2 * 1024 * 1024, 16384),
    # BASELINE config 3
    "traffic-64k": (dict(traffic_density=0.05, random_map_obstacle_probability=0.2), 65536, 4096),
    # BASELINE config 4
    "large-1M": (dict(random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8, traffic_density=0.2,
                      random_map_obstacle_probability=0.5), 1024 * 1024, 256),
    # small smoke-sized run
    "default-64k": (dict(), 65536, 8192),
}


def algorithmic_bytes(kw: dict) -> float:
    """bytes one env-step must move (int8 planes, scalars, agent state r/w, tile
    descriptors + used bits, car records r/w)."""
    C = len(kw.get("features_to_include_in_observation", [0] * 9))
    P = 9 if not kw.get("use_sliding_observation_window") else 1 + 2 * kw.get("sliding_observation_window_size", 4)
    T = kw.get("random_map_width", 4) * kw.get("random_map_height", 4)
    lane_sq = {4: 304, 64: 1150}.get(T, 19 * T)  # measured mean lane squares per map (SURVEY a8 / probe)
    n_cars = int(lane_sq * kw.get("traffic_density", 0.0))
    return C * P * P + 16 + 8 + 2 + 1 + 2 * 16 + T * 2 + 2 * ((T + 7) // 8) + 2 * 8 * n_cars


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (profiling guide recipe):
    one streaming `nvidia-smi -lms 50` process, rows kept only if sampled inside [begin, end]."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        time.sleep(0.12)
        self.proc.terminate()
        self.thread.join(timeout=3)
        rows = [r for t, r in self.rows if self.t_begin - 0.02 <= t <= self.t_end + 0.08 and len(r) >= 10]
        if not rows:
            rows = [r for _, r in self.rows[-2:] if len(r) >= 10]
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None  # noqa: E731
        sm = [num(r[2]) for r in rows if num(r[2]) is not None]
        mx = [num(r[3]) for r in rows if num(r[3]) is not None]
        pw = [num(r[4]) for r in rows if num(r[4]) is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    power_w_max=max(pw) if pw else None, samples=len(rows))


def cpu_baseline(kw: dict, n_envs: int, seconds: float, threads: int) -> dict:
    """The oracle (C port of the reference tick, oracle/pgtg_oracle.c) on the host cores."""
    from oracle.oracle import OracleVectorEnv

    env = OracleVectorEnv(num_envs=n_envs, threads=threads, seed=1, **kw)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 9, n_envs).astype(np.int32) for _ in range(4)]
    env.step(acts[0])
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        env.step(acts[steps % 4])
        steps += 1
    dt = time.perf_counter() - t0
    env.close()
    return dict(value=n_envs * steps / dt, unit="env-steps/s", cores=threads, kind="port",
                sample=f"{n_envs} envs x {steps} ticks ({dt:.1f} s) of the same workload, C oracle port, {threads} threads, incl. auto-reset")


def run_reference(args, kw, n_cpu):
    """`--impl reference`: the reference's CPU implementation of the path -- here the oracle port
    (the Python reference cannot travel to the GPU box) -- with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle.oracle import OracleVectorEnv

    env = OracleVectorEnv(num_envs=n_cpu, threads=threads, seed=1, **kw)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 9, n_cpu).astype(np.int32) for _ in range(4)]
    inner = 8  # ticks per "step" of this arm: a bounded sample of the workload
    for i in range(args.warmup):
        env.step(acts[i % 4])
    t0 = time.perf_counter()
    for s in range(args.steps):
        for i in range(inner):
            env.step(acts[(s + i) % 4])
    dt = time.perf_counter() - t0
    value = n_cpu * inner * args.steps / dt
    sample = f"{n_cpu} envs x {inner} ticks per step, C oracle port of the reference tick, {threads} threads, incl. auto-reset"
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32+f64", "data": "synthetic", "config": {"workload": args.workload, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="default-2M", choices=list(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (overrides the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--final-observation", action="store_true")
    args = ap.parse_args()
    kw, n_gpu_envs, n_cpu = WORKLOADS[args.workload]
    if args.envs:
        n_gpu_envs = args.envs
    if args.impl == "reference":
        run_reference(args, kw, n_cpu)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pgtg_b200 import PGTGVectorEnv

    N = n_gpu_envs
    env = PGTGVectorEnv(N, device=dev, seed=2026, env_id_base=rank * N, final_observation=args.final_observation, **kw)
    env.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    pool = [torch.randint(0, 9, (N,), device=dev, dtype=torch.int32, generator=g) for _ in range(8)]
    for i in range(max(args.warmup, 3)):
        env.step(pool[i % 8])
    env.episode_stats(reset=True)  # first call loads the reduction kernel; not part of the timed region
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: K fused launches, inputs resident in HBM -------------------------
    launches0 = env.launch_count()
    env.raw.enable_timing(args.steps)  # CUDA events around each kernel, on the launching stream
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    if sampler:
        sampler.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        kev[i][0].record()
        env.step(pool[i % 8])
        kev[i][1].record()
    stats = env.episode_stats(all_reduce=True)  # the only collective: 8 doubles, once
    ev1.record()
    barrier()
    if sampler:
        sampler.end()
    ms = ev0.elapsed_time(ev1)
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ktime = env.raw.timing()
    env.raw.enable_timing(0)
    kernel_ms = ktime["tick_ms"]  # the dominant kernel: the fused tick
    launches = env.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * N * args.steps / (ms * 1e-3)

    # ---- the dominant kernel timed ALONE (overlap off: kernels back to back on this stream) --------
    alone_steps = min(args.steps, 40)
    env.raw.set_overlap(False)
    for i in range(3):
        env.step(pool[i % 8])
    env.raw.enable_timing(alone_steps)
    for i in range(alone_steps):
        env.step(pool[i % 8])
    alone = env.raw.timing()
    env.raw.enable_timing(0)
    env.raw.set_overlap(True)
    env.episode_stats(reset=True)

    # ---- end to end through the host-buffer entry (pgtg_step_host) ------------------------------
    C, P = env.hc.pod.num_channels, env.hc.window
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    out = dict(obs_map=pin((N, C, P, P), torch.int8), obs_position=pin((N, 2), torch.int32), obs_velocity=pin((N, 2), torch.int32),
               reward=pin((N,), torch.float64), terminated=pin((N,), torch.uint8), truncated=pin((N,), torch.uint8))
    hact = [torch.randint(0, 9, (N,), dtype=torch.int32).pin_memory().numpy() for _ in range(2)]
    e2e_value = None
    if args.e2e_steps > 0:
        env.step_host(hact[0], out)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            env.step_host(hact[i % 2], out)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * N * args.e2e_steps / float(t.item())
    h2d = N * 4
    d2h = sum(v.nbytes for v in out.values())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        b_alg = algorithmic_bytes(kw)
        overlapped_kernel_ms = kernel_ms
        kernel_ms = alone["tick_ms"]
        achieved = b_alg * N / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload)
        cpu = cpu_baseline(kw, n_cpu, args.cpu_seconds, os.cpu_count() or 1) if args.cpu_seconds > 0 else None
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": args.workload, "envs_per_gpu": N, "kwargs": kw, "actions": "uniform random, 8 resident int32 tensors cycled",
                       "rng": "philox4x32-10 per env", "auto_reset": "same step, on-device map regeneration",
                       "final_observation": bool(args.final_observation),
                       "l2": f"working set {(b_alg * N) / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "path": "pgtg_step_host: pinned host actions in, observation/reward/flags out"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_env_step": b_alg, "kernel_ms": kernel_ms,
                         "kernel": "pgtg_tick_kernel (fused tick: step + auto-reset + observation write)",
                         "note": "kernel_ms/achieved/frac: the tick kernel timed alone (overlap off, CUDA events around the launch, "
                                 f"{alone_steps} launches); in the pipeline it shares the SMs with the map-generation kernel",
                         "tick_ms_in_pipeline": overlapped_kernel_ms, "mapgen_ms_in_pipeline": ktime["mapgen_ms"],
                         "mapgen_ms_alone": alone["mapgen_ms"], "step_ms": step_ms,
                         "whole_step_frac": b_alg * N / (step_ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "episode_stats": stats,
        }
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
