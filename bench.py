#!/usr/bin/env python
"""bench.py -- MC bead-updates/sec of the PIGS hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] -- liquid He-4, N = 256
atoms at rho = 0.365 sigma^-3, 2M = 30 beads, Chin action, bisection moves,
worm algorithm on (CWorm = 0.5, Nobdm = 10, swap), mixed + thermodynamic
energies, g(r), S(k) and OBDM estimators -- run as `chains` independent Markov
chains per GPU (weak scaling: the per-GPU chain count is fixed).  One "step" =
one Monte-Carlo block of `mc_steps_per_block` driver steps (vpi.f90:297-475) on
every chain, followed by the block-boundary reduction (all-reduce for N > 1).
One bead-update = one UpdateAction evaluation (vpi_mod.f90:2491), counted on
the device per slice class.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the
library's stream, inputs resident in HBM; `e2e` adds the host<->device copies
of the chains' state through the C ABI with host buffers.  The reference arm
times the CPU oracle (the C++ restatement of the Fortran reference -- no Fortran
compiler exists in this image) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "bead-updates/s"
CHAINS_PER_GPU = 2368          # C3 / C2: 148 SMs x 16 chain groups, the same on every GPU (weak scaling)
C5_CHAINS = 4096               # C5 = BASELINE.json configs[4]: 4096 N=64 chains split over the GPUs (strong scaling)


def metric_name(workload):
    return {"C3": "MC bead-updates/sec, liquid He-4 N=256", "C2": "MC bead-updates/sec, liquid He-4 N=64",
            "C5": "MC bead-updates/sec, 4096 independent He-4 N=64 chains"}[workload]


MC_STEPS_PER_BLOCK = 4
SEED = 20260101


def oracle_cfg(cfg):
    c = dict(cfg)
    for k in ("trap", "swapping", "wf_table", "v_table", "crystal"):
        if k in c:
            c[k] = int(bool(c[k]))
    c.pop("tables", None)
    return c


# ------------------------------------------------------------------ CPU legs (the only oracle/ users here)
def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_workers(workload, W, K, impl="ref", cores=None):
    """one independent reference process per host core, pinned, own seed, same run.
    Returns (bead-updates/s summed over the processes, seconds of the slowest, kind, n_processes)."""
    import subprocess
    cores = host_cores() if cores is None else cores
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.ref_worker", str(c), str(1982 + i), str(W), str(K), workload, impl],
                              cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for i, c in enumerate(cores)]
    upd, tmax, kind = 0, 0.0, "?"
    for pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("reference worker failed: " + err[-800:])
        d = json.loads(out.strip().splitlines()[-1])
        upd += d["updates"]
        tmax = max(tmax, d["seconds"])
        kind = d["kind"]
    return upd / tmax, tmax, kind, len(cores)


def cpu_baseline(workload, W=1, K=8, reps=1, extras=True):
    """the reference program (oracle/_ref: its Fortran machine-translated to C++, g++ -O2) on every host core;
    beside it the hand-written port at -O2 and at -O3 -march=x86-64-v3"""
    vals, secs = [], 0.0
    for _ in range(reps):
        v, t, kind, n = run_workers(workload, W, K, "ref")
        vals.append(v)
        secs += t
    vals.sort()
    out = dict(value=vals[len(vals) // 2], unit=UNIT, cores=n, seconds_per_rep=secs / reps,
               kind="reference" if kind.startswith("reference") else "port", implementation=kind, cpu=cpu_model(),
               sample=f"{n} pinned processes (one per host core, taskset-style affinity), each one independent {workload} chain "
                      f"running the reference program for {K} MC steps after {W} warm-up steps; {reps} repetition(s), median; "
                      f"{secs:.1f} s of CPU wall time",
               spread=[vals[0], vals[-1]])
    if extras:
        out["value_port_O2"] = run_workers(workload, W, K, "port")[0]
        out["value_port_native"] = run_workers(workload, W, K, "native")[0]
    return out


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args.workload, W=max(args.warmup, 1), K=args.steps, reps=3, extras=True)
    v = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.workload), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * cb["seconds_per_rep"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong" if args.workload == "C5" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "Np": cfg["Np"], "Nb": cfg["Nb"], "chains": cb["cores"], "mc_steps_per_step": 1},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if bit and (r & bit) and n not in ("nvmlClocksThrottleReasonAll", "nvmlClocksThrottleReasonNone", "nvmlClocksThrottleReasonGpuIdle"):
                        self.reasons.add(n.replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.1)
        except Exception as e:       # clocks are evidence, not a dependency
            self.reasons.add(f"sampler-error:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------ the GPU arm
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from pathintegralgroundstate_b200 import PigsCuda, measure_fp64_peak
    from pathintegralgroundstate_b200.multi_gpu import init_process_group, shard_chains, _CudaArray
    from pathintegralgroundstate_b200.workloads import synthetic_paths, flops_per_bead_update

    rank, local, world = init_process_group("nccl")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    strong = args.workload == "C5"
    if strong:                                   # total work fixed: the 4096 chains are split over the ranks
        first, n_chains = shard_chains(C5_CHAINS if args.chains == CHAINS_PER_GPU else args.chains, rank, world)
    else:
        n_chains = args.chains
        first, _ = shard_chains(n_chains * world, rank, world)

    sim = PigsCuda(cfg, n_chains=n_chains, rng=args.rng, seed=SEED, chain_offset=first, device=local, schedule=args.schedule)
    sim.fill_tables("hfdb")
    P, xe = synthetic_paths(cfg, n_chains, seed=SEED + 7919 * rank)
    # pinned host buffers of the chains' state (the e2e leg copies them every step)
    hP = torch.from_numpy(P).pin_memory()
    hX = torch.from_numpy(xe).pin_memory()
    hO = torch.zeros(n_chains, dtype=torch.int32).pin_memory()
    hW = torch.zeros(n_chains, dtype=torch.int32).pin_memory()
    del P, xe
    sim.set_state_all(hP.numpy(), hX.numpy(), hO.numpy(), hW.numpy())
    lib_stream = torch.cuda.ExternalStream(sim.stream(), device=local)
    ptr, nvec = sim.block_vector()
    vec_t = torch.as_tensor(_CudaArray(ptr, nvec), device=f"cuda:{local}")
    nstep = args.mc_steps

    def one_step():
        sim.run_block(nstep, sync=False)
        if world > 1:
            sim.sync()
            dist.all_reduce(vec_t, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        sim.sync()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()

    # ---- timed region: K steps, CUDA events on the library's stream, max over ranks
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sim.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, upd_cls = 0.0, np.zeros(3)
    barrier()
    t0 = time.perf_counter()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        one_step()
        if world > 1:
            torch.cuda.synchronize()
            b = sim.unpack_block_vector(vec_t.cpu().numpy())[0]
        else:
            sim.sync()
            b = sim.get_block()[0]
        upd_cls += np.array(b["bead_updates"], dtype=float)
        kernel_ms += sim.last_block_ms()
        if not (np.isfinite(b["sumE"]) and np.isfinite(b["sumEt"]) and b["idiag_block"] > 0):
            raise SystemExit(f"bench: the block result is not finite (sumE {b['sumE']}, sumEt {b['sumEt']}): refusing to time garbage")
    last_block = {"E_per_particle_mixed": float(b["sumE"]) / max(int(b["idiag_block"]), 1) / cfg["Np"],
                  "E_per_particle_thermodynamic": float(b["sumEt"]) / max(int(b["idiag_block"]), 1) / cfg["Np"],
                  "diagonal_samples": int(b["idiag_block"])}
    ev1.record(lib_stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = sim.launch_count() - l0
    clocks = sampler.stop()
    t_ms = torch.tensor([dev_ms, wall * 1e3, kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        # bead-updates counted in the all-reduced vector are already global
        upd_total = float(upd_cls.sum())
    else:
        upd_total = float(upd_cls.sum())
    dev_ms, wall_ms, kernel_ms = (float(x) for x in t_ms.cpu())
    value = upd_total / (dev_ms * 1e-3)

    # ---- e2e: the same K steps through the C ABI with HOST buffers (state up, block, results + state down)
    barrier()
    e2e_upd = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.set_state_all(hP.numpy(), hX.numpy(), hO.numpy(), hW.numpy())
        one_step()
        if world > 1:
            torch.cuda.synchronize()
            b = sim.unpack_block_vector(vec_t.cpu().numpy())[0]
        else:
            sim.sync()
            b = sim.get_block()[0]
        e2e_upd += float(sum(b["bead_updates"]))
        sim.get_state_all(out=(hP.numpy(), hX.numpy(), hO.numpy(), hW.numpy()))     # straight into the pinned buffers
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = e2e_upd / float(e2e_t.item())
    state_bytes = hP.numel() * 8 + hX.numel() * 8 + 2 * n_chains * 4

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (the persistent sweep kernel)
    fl = flops_per_bead_update(cfg["Np"])
    if world > 1:
        flops = sum(f * n for f, n in zip(fl, upd_cls)) / world      # this rank's share of the global count
    else:
        flops = sum(f * n for f, n in zip(fl, upd_cls))
    peak = measure_fp64_peak(local)
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "note": "FP64 vector pipe (DFMA); peak measured live with pigs_measure_fp64_peak (MEASURED_PEAKS.json has "
                        "no FP64 entry); achieved = algorithmic flops 2(N-1)c+40 per bead-update (c=28/46/37 by slice "
                        "class) / CUDA-event time of the sweep kernel launches"}
    out = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "schedule": sim.schedule_name(), "Np": cfg["Np"], "Nb": cfg["Nb"], "chains_per_gpu": n_chains,
                   "mc_steps_per_step": nstep, "rng": args.rng, "worm": "on (CWorm=0.5, Nobdm=10, swap)",
                   "estimators": "mixed+thermodynamic energy, g(r), S(k), OBDM",
                   "l2": f"inputs larger than L2: {state_bytes / 1e6:.0f} MB of paths per GPU stream from HBM",
                   "parallelism": f"{world} x independent chain shards, one NCCL all-reduce of the block vector per step"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": state_bytes,
                "d2h_bytes_per_step": state_bytes + nvec * 8},
        "gpu_launches": int(launches),
        "last_block": last_block,
        "roofline": roofline,
        "wall_ms_per_step": wall_ms / args.steps,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--mc-steps", type=int, default=MC_STEPS_PER_BLOCK, help="driver MC steps per bench step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C3", choices=["C3", "C2", "C5"],
                    help="C3: the metric's configuration (N=256, worm on), weak scaling; C5: 4096 N=64 chains split over the GPUs")
    ap.add_argument("--rng", default="philox", choices=["philox", "mt"],
                    help="philox: the production streams; mt: every chain replays its own MT19937 stream in the reference's order")
    ap.add_argument("--schedule", type=int, default=-1, help="-1 auto, 0 one warp per chain, 1 team (4 warps per chain)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                      # timing hygiene: at least 3 warm-up steps
    from pathintegralgroundstate_b200.workloads import config
    cfg = config("C2" if args.workload == "C5" else args.workload)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
