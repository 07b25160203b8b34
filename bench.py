#!/usr/bin/env python
"""bench.py — KMC events/s of the hot path (BASELINE.json `metric`) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline workload — the shape the north_star target is quoted on: a 128^3 lattice with 10^6 concurrent
trajectories per GPU on the parity path (the reference's own First Reaction Method, reseau.py:483-562: default
0.84/0.15/0.01 host/TADF/emitter composition, charges = 1: one electron and one hole in flight per trajectory,
field 0.1 eV/nm), the only path whose event sequences are bit-exact under replayed RNG.  One bench "step" = one
`hf_run` launch that advances every trajectory by EVENTS_PER_STEP KMC events (`Lattice.operations`,
reseau.py:580-595).  An "event" is one fired `_first_reaction_method` call.  Weak scaling: every rank carries
its own 10^6 trajectories (global Philox ids) on the one shared lattice realisation, no data-path collective; the
one NCCL all-reduce of the tally vector is inside every `e2e` step.  At N > 1 a second field, `strong`, splits
10^6 trajectories in total over the ranks.  BASELINE configs[1] (64^3, 10^5 trajectories) stays as the leg
`configs1`.

One JSON line is printed by rank 0; see DESIGN.md section 5 for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner with a
# plain printf whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run
# and the JSON line is written to a private duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIMS = (128, 128, 128)
PROPS = (0.84, 0.15, 0.01)
CHARGES = 1
FIELD = 0.1
SEED = 0x48464B4D43            # "HFKMC" 
TRAJ_PER_GPU = 1_000_000
EVENTS_PER_STEP = 1000         # KMC events per trajectory per bench step (one launch: 10^9 events, ~0.18 s)
C1_DIMS, C1_TRAJ_PER_GPU, C1_EVENTS = (64, 64, 64), 100_000, 1000     # BASELINE configs[1], the leg `configs1`
NEIGHBOURS = 26
# compulsory bytes per event, no caching credit
BYTES_PER_EVENT = NEIGHBOURS * (8 + 1) + 2 * CHARGES * 4 + 2 * CHARGES * 16 + 24 + 2
FALLBACK_HBM_GBS = 6650.0      # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
WORKLOAD = (f"north_star shape, parity path: {TRAJ_PER_GPU} concurrent trajectories/GPU x {EVENTS_PER_STEP} KMC events per "
            f"step on one {DIMS[0]}^3 lattice, charges={CHARGES}, props={PROPS}, field={FIELD}")
# the SAME dict in both arms (ours / --impl reference): what differs between them (the bounded CPU sample) is
# stated in `cpu_baseline.sample`, not here
CONFIG = {
    "workload": WORKLOAD, "lattice": list(DIMS), "charges": CHARGES, "props": list(PROPS), "field": FIELD,
    "trajectories_per_gpu": TRAJ_PER_GPU, "events_per_trajectory_per_step": EVENTS_PER_STEP,
    "events_fired": "move events only: with one carrier of each type the +0.1 eV per step towards the collecting electrode "
                    "(quirk Q8, reseau.py:287, 319) keeps both at their injection faces, so bound / decay / capture / "
                    "re-injection do not occur in the timed region; that is what the reference does on this device too",
    "l2": "flushed between timed steps (256 MiB memset); the 34 MB of HOMO / LUMO energies are L2-resident by nature",
    "timing": "CUDA events around each hf_run on its stream, summed over steps, max over ranks",
}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def _traffic_entry(key):
    """An ncu-derived entry of profiles/traffic.json, or None when it is absent or was captured from other kernel
    sources than the ones that are built now (each entry carries the sha256 of the files its kernel comes from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[key]
        from hyperfluorescence_b200.build import source_hashes
        stamp = e.get("sources")
        if not stamp or source_hashes(stamp.keys()) != stamp:
            return None
        return e
    except Exception:
        return None


def measured_traffic(key="hf_run_c1_kernel", events_per_launch=None):
    """dram bytes per launch of a step kernel from the committed ncu capture (scaled per event), if still valid."""
    e = _traffic_entry(key)
    if not e or events_per_launch is None:
        return None
    return float(e["dram_bytes_per_event"]) * events_per_launch


def issue_roofline(events_per_s_per_gpu, key, clocks=None):
    """What actually bounds these kernels (DESIGN.md sections 3, 9): instruction issue.  Warp instructions per
    event come from the committed ncu capture (profiles/traffic.json), the rate is measured live, the peak is
    4 schedulers x 148 SMs x the SM clock sampled during the run (else the nominal 1.965 GHz)."""
    e = _traffic_entry(key)
    if not e:
        return None
    wipe = float(e["warp_inst_per_event"])
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak = 148 * 4 * mhz * 1e6
    achieved = events_per_s_per_gpu * wipe
    return {"bound": "issue", "warp_inst_per_event": wipe, "achieved": achieved / 1e9, "peak": peak / 1e9,
            "unit": "Gwarp-inst/s", "frac": achieved / peak, "source": e.get("source")}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


_ORACLE_LATTICES = {}


def cpu_sample(cores, per_core_trajectories=500, dims=DIMS, events=EVENTS_PER_STEP, total=TRAJ_PER_GPU):
    """The CPU baseline: oracle port (oracle/frm_oracle.c, pinned bit-exactly to the unmodified
    reference) on `cores` host threads, a bounded sample of the same workload."""
    from oracle import frm_oracle as fo
    if dims not in _ORACLE_LATTICES:
        _ORACLE_LATTICES[dims] = fo.OracleLattice.generate(dims, PROPS, SEED)       # the device generates the same one
    lat = _ORACLE_LATTICES[dims]
    n_traj = min(total, cores * per_core_trajectories)
    t0 = time.perf_counter()
    res = lat.run_batch(CHARGES, FIELD, SEED, 0, n_traj, events, n_threads=cores)
    dt = time.perf_counter() - t0
    sample = (f"the first {n_traj} of the {total} trajectories x {events} events on the same {dims[0]}^3 lattice "
              f"(charge injection and initial events included), {cores} threads")
    return res["events"] / dt, sample, res


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference is pure
    Python and cannot travel to the GPU box, so this times its C restatement (oracle port) with all
    host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_sample(cores, 16)                                                           # builds the 128^3 lattice (untimed)
    for _ in range(args.warmup):
        cpu_sample(cores, 50)
    t0 = time.perf_counter()
    events = 0
    sample = ""
    for _ in range(args.steps):
        v, sample, res = cpu_sample(cores)
        events += res["events"]
    dt = time.perf_counter() - t0
    value = events / dt
    line = {
        "impl": "reference", "metric": "KMC exciton events/sec (box)", "value": value, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port",
                         "sample": "each step: " + sample,
                         "note": "C restatement of the reference (bit-pinned to it); the unmodified Python reference does "
                                 "~1e3 events/s per core on this device (SURVEY.md section 6) and cannot build a 128^3 "
                                 "lattice (~13 GB of Python objects)"},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def max_over_ranks(x, world):
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def path_p_leg(nat, hd, rank, local_rank, world, dims, B, S, K, W, first_trajectory, clocks_of=None, e2e_steps=2):
    """K timed hf_run launches of B trajectories x S events on one `dims` lattice (Path P, charges = 1), then the
    same work end to end through the C-ABI with host buffers.  Returns a dict of raw measurements."""
    import torch
    import torch.distributed as dist
    stream = torch.cuda.current_stream()
    sim = nat.Sim(dims, PROPS, FIELD, CHARGES, 1, n_trajectories=B, n_realisations=1, seed=SEED,
                  first_trajectory=first_trajectory, first_realisation=0, device=local_rank)
    sim.set_stream(stream.cuda_stream)            # timed with torch events: the launches must be on torch's stream
    sim.generate_lattice()
    sim.inject()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(W):
        sim.run(S)
    torch.cuda.synchronize()
    sim.run_timing()                                                            # reset the lib's own event timers
    sampler = ClockSampler(local_rank) if clocks_of is not None else None
    if sampler:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()                                                           # L2 flush between timed iterations
        starts[k].record(stream)
        sim.run(S)
        stops[k].record(stream)
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in zip(starts, stops)), world)
    lib_ms, lib_launches = sim.run_timing()                                     # same launches, timed inside the lib
    tallies = hd.all_reduce_tallies(sim)
    fired_expected = world * B * S * (K + W)
    if tallies["fired"] != fired_expected or tallies["stopped"]:
        raise SystemExit(f"bench: {tallies['fired']} events fired, {fired_expected} expected, {tallies['stopped']} "
                         "trajectories stopped: the timed region did not do the work it is credited with")

    # end to end through the C-ABI with HOST buffers, every step: lattice arrays from pinned host memory ->
    # hf_upload_lattice (H2D), hf_inject, hf_run, tally reduction on the device + NCCL all-reduce over the ranks
    # (the one collective of the path), device -> host read of the result.  Wall clock, max over ranks.
    host = sim.download_lattice(0)
    pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in host.items()}
    h2d = sum(v.nbytes for v in pinned.values())
    d2h = nat.HF_N_TALLIES * 8

    def e2e_step():
        sim.upload_lattice(0, pinned["species"], pinned["homo"], pinned["lumo"], pinned["s1"], pinned["t1"])
        sim.inject()
        sim.run(S)
        return hd.all_reduce_tallies(sim)
    e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    if got["fired"] != world * B * S:
        raise SystemExit(f"bench e2e: {got['fired']} events fired, {world * B * S} expected")
    sim.close()
    del flush
    return dict(total_ms=total_ms, lib_ms=lib_ms, lib_launches=lib_launches, wall_s=t_wall, clocks=clocks, tallies=tallies,
                fired_expected=fired_expected, e2e_value=world * B * S * e2e_steps / e2e_s, e2e_steps=e2e_steps, h2d=h2d, d2h=d2h,
                value=world * B * S * K / (total_ms * 1e-3))


E2E_PATH = ("hf_upload_lattice(pinned host) + hf_inject + hf_run + hf_reduce_tallies_device + NCCL all-reduce of the tallies "
            "(N > 1) + device-to-host read")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hyperfluorescence_b200 import _native as nat
    from hyperfluorescence_b200 import distributed as hd

    rank, local_rank, world = hd.init_process_group()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hyperfluorescence_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    B, S, K, W = TRAJ_PER_GPU, EVENTS_PER_STEP, args.steps, args.warmup

    # weak scaling: rank r owns global trajectories [r*B, (r+1)*B) of the one shared realisation 0
    m = path_p_leg(nat, hd, rank, local_rank, world, DIMS, B, S, K, W, rank * B, clocks_of=True, e2e_steps=max(2, min(K, 3)))
    value, clocks, tallies = m["value"], m["clocks"], m["tallies"]
    peak, peak_src = hbm_peak()
    launch_s = (m["lib_ms"] / max(m["lib_launches"], 1)) * 1e-3
    achieved = B * S * BYTES_PER_EVENT / launch_s / 1e9
    line = {
        "metric": "KMC exciton events/sec (box)", "value": value, "unit": "events/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": m["total_ms"] / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "wall_s_bracket": m["wall_s"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic("hf_run_c1_kernel", B * S), "peak_source": peak_src, "kernel": "hf_run_c1_kernel",
                     "bytes_per_event": BYTES_PER_EVENT, "events_per_launch": B * S,
                     "avg_launch_ms": m["lib_ms"] / max(m["lib_launches"], 1),
                     "note": "not HBM-bound: the 34 MB of site energies are L2-resident and trajectory state stays in registers "
                             "for the whole launch (ncu: dram bytes per launch in `traffic`); the binding unit is instruction "
                             "issue, see `issue`, DESIGN.md section 3 and profiles/"},
        "e2e": {"value": m["e2e_value"], "unit": "events/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "steps": m["e2e_steps"], "path": E2E_PATH},
        "issue": issue_roofline(value / world, "hf_run_c1_kernel", clocks),
        "gpu_launches": K,
        "clocks": clocks,
        "tallies": dict({k: tallies[k] for k in ("emission", "recombination", "injection", "fired", "stopped", "move_e", "move_h")},
                        fired_expected=m["fired_expected"]),
    }
    if world > 1:
        # strong scaling of the same workload: 10^6 trajectories IN TOTAL, split over the ranks
        Bs = TRAJ_PER_GPU // world
        ms = path_p_leg(nat, hd, rank, local_rank, world, DIMS, Bs, S, K, W, rank * Bs, e2e_steps=2)
        line["strong"] = {"value": ms["value"], "unit": "events/s", "scaling": "strong", "trajectories_total": Bs * world,
                          "trajectories_per_gpu": Bs, "ms_per_step": ms["total_ms"] / K,
                          "e2e": {"value": ms["e2e_value"], "unit": "events/s", "h2d_bytes_per_step": ms["h2d"],
                                  "d2h_bytes_per_step": ms["d2h"], "steps": ms["e2e_steps"]},
                          "note": "hf_run_c1_kernel holds 148 SMs x 768 threads = 113 664 trajectories per wave; a shard that is not "
                                  "a multiple of that pays for a partly filled last wave"}
    if not args.no_ga:
        mc = path_p_leg(nat, hd, rank, local_rank, world, C1_DIMS, C1_TRAJ_PER_GPU, C1_EVENTS, 3, 2, rank * C1_TRAJ_PER_GPU, e2e_steps=2)
        line["configs1"] = {"metric": "KMC exciton events/sec (box), BASELINE configs[1]", "value": mc["value"], "unit": "events/s",
                            "config": {"lattice": list(C1_DIMS), "trajectories_per_gpu": C1_TRAJ_PER_GPU,
                                       "events_per_trajectory_per_step": C1_EVENTS, "charges": CHARGES},
                            "e2e": {"value": mc["e2e_value"], "unit": "events/s", "h2d_bytes_per_step": mc["h2d"],
                                    "d2h_bytes_per_step": mc["d2h"], "steps": mc["e2e_steps"]}}
        line["ga"] = with_clocks(ga_fitness_rate, local_rank, rank, local_rank, world)
        line["device_eqe"] = with_clocks(device_eqe_rate, local_rank, rank, local_rank, world)
        line["exciton_path"] = with_clocks(exciton_rate, local_rank, rank, local_rank, world, cpu=not args.no_cpu)
        line["disorder_sweep"] = with_clocks(exciton_rate, local_rank, rank, local_rank, world, lattice=256, realisations_total=64,
                                             label="disorder_sweep")
        line["high_density"] = with_clocks(dense_rate, local_rank, rank, local_rank, world, cpu=not args.no_cpu)
    if world == 1 and rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, sample, _ = cpu_sample(cores, 1500)
        v1, _, _ = cpu_sample(1, 3000)
        line["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample,
                                "one_core_value": v1}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


GA_DIMS, GA_CHARGES, GA_POPULATION, GA_TRAJECTORIES, GA_STEPS = (10, 10, 5), 4, 256, 10_000, 500


def with_clocks(fn, local_rank, *a, **kw):
    """Run one bench leg with the SM clock / throttle sampler around it; the leg's dict gets `clocks`."""
    sampler = ClockSampler(local_rank)
    sampler.start()
    try:
        out = fn(*a, **kw)
    finally:
        clocks = sampler.stop()
    if isinstance(out, dict):
        out["clocks"] = clocks
    return out


GA_REPS = 3


def ga_fitness_rate(rank, local_rank, world):
    """BASELINE configs[2]: one fitness sweep of 256 composition candidates x 10^4 trajectories on the
    reference driver's device (OLED.py:6-10: 10x10x5, charges = 4), population sharded over the
    ranks, fitness = IQE (reseau.py:595).  Timed end to end: lattice generation, injection, stepping,
    per-candidate tallies, all-gather."""
    import numpy as np
    import torch
    from hyperfluorescence_b200.evolution import Compositions, DeviceFitness, GAConfig
    cfg = GAConfig(population=GA_POPULATION, trajectories_per_candidate=GA_TRAJECTORIES, steps=GA_STEPS)
    fit = DeviceFitness(GA_DIMS, charges=GA_CHARGES, cfg=cfg, seed=SEED, device=local_rank, rank=rank, world=world)
    pop = Compositions(GA_DIMS).random(np.random.default_rng(7), GA_POPULATION)
    fit(pop)                                                        # warm-up (allocations, first launch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(GA_REPS):                                         # the mean of GA_REPS sweeps: a single one varies by +- 20 %
        f = fit(pop)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / GA_REPS
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    fit.close()
    # the same sweep with the exciton path's EQE as the fitness (own model: DESIGN.md section 9)
    from hyperfluorescence_b200.evolution import ExcitonFitness
    xdims, xsteps = (32, 32, 16), 200
    xcfg = GAConfig(population=GA_POPULATION, trajectories_per_candidate=GA_TRAJECTORIES, steps=xsteps)
    xfit = ExcitonFitness(xdims, cfg=xcfg, seed=SEED, device=local_rank, rank=rank, world=world)
    xpop = Compositions(xdims).random(np.random.default_rng(7), GA_POPULATION)
    xfit(xpop)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(GA_REPS):
        xf = xfit(xpop)
    torch.cuda.synchronize()
    xdt = (time.perf_counter() - t0) / GA_REPS
    if world > 1:
        t = torch.tensor([xdt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        xdt = float(t.item())
    xfit.close()
    eqe = {"metric": "GA fitness evals/sec (EQE from the exciton path)", "value": GA_POPULATION / xdt, "unit": "evals/s",
           "events_per_s": GA_POPULATION * GA_TRAJECTORIES * xsteps / xdt,
           "config": {"lattice": list(xdims), "population": GA_POPULATION, "trajectories_per_candidate": GA_TRAJECTORIES,
                      "events_per_trajectory": xsteps, "cutoff": 4.0},
           "best_eqe_percent": float(xf.max()), "mean_eqe_percent": float(xf.mean())}
    return {"metric": "GA fitness evals/sec", "value": GA_POPULATION / dt, "unit": "evals/s", "eqe_sweep": eqe,
            "events_per_s": GA_POPULATION * GA_TRAJECTORIES * GA_STEPS / dt,
            "config": {"lattice": list(GA_DIMS), "charges": GA_CHARGES, "population": GA_POPULATION,
                       "trajectories_per_candidate": GA_TRAJECTORIES, "events_per_trajectory": GA_STEPS,
                       "sharding": f"{GA_POPULATION // world} candidates per GPU"},
            "best_iqe_percent": float(f.max()), "mean_iqe_percent": float(f.mean()),
            "timing": "wall clock around the public call DeviceFitness(population), mean of 3 sweeps, end to end: compositions host -> device "
                      "(256 x 3 doubles), lattice generation, injection, stepping, per-candidate tallies device -> host, all-gather",
            "e2e": {"value": GA_POPULATION / dt, "unit": "evals/s", "h2d_bytes_per_step": GA_POPULATION // world * 24,
                    "d2h_bytes_per_step": GA_POPULATION // world * 40}}


DE_TRAJ_PER_GPU, DE_STEPS, DE_CUTOFF = 200_000, 2000, 3.0


def device_eqe_rate(rank, local_rank, world):
    """The integrated device: the reference driver's own device (OLED.py:6-10: 10x10x5,
    charges = 4) with the excitons living in the SAME First Reaction Method as the carriers — formed where the
    carriers meet (reseau.py:529-533), then hopping (Forster / Dexter), crossing (ISC / RISC) and decaying with the
    exciton path's rate laws instead of decaying at tau = 0.  Reports events/s and the device EQE = outcoupling x IQE.
    The exciton part is this repository's own model (parity unpinned); with the excitons set to decay at once the
    same code path reproduces every golden trajectory of the reference bit for bit."""
    from hyperfluorescence_b200 import _native as nat
    from hyperfluorescence_b200 import distributed as hd
    B = DE_TRAJ_PER_GPU
    sim = nat.Sim(GA_DIMS, PROPS, FIELD, GA_CHARGES, 1, n_trajectories=B, n_realisations=1, seed=SEED,
                  first_trajectory=rank * B, device=local_rank)
    sim.generate_lattice()
    params = nat.x_params_from_dict(dict(cutoff=DE_CUTOFF))
    sim.x_configure(params)
    sim.p_excitons(2)
    sim.inject()
    sim.run(DE_STEPS // 4)                                     # warm-up: the first excitons form
    sim.sync()
    sim.run_timing()
    for _ in range(3):
        sim.run(DE_STEPS)
    ms, n = sim.run_timing()
    t = hd.all_reduce_tallies(sim)
    sim.close()
    ms = max_over_ranks(ms, world)
    fired_timed = world * B * DE_STEPS * n
    iqe = 200.0 * t["emission"] / max(t["injection"], 1)
    return {"metric": "KMC events/sec (box), integrated device (carriers + excitons in one event loop)",
            "value": fired_timed / (ms * 1e-3), "unit": "events/s",
            "config": {"lattice": list(GA_DIMS), "charges": GA_CHARGES, "trajectories_per_gpu": B, "events_per_trajectory_per_launch": DE_STEPS,
                       "cutoff": DE_CUTOFF, "kernel": "hf_run_small_kernel<true>",
                       "model": "own (parity unpinned) beyond the decay-at-once gate: DESIGN.md section 11"},
            "device_iqe_percent": iqe, "device_eqe_percent": params.outcoupling * iqe,
            "events": {k: t[k] for k in ("fired", "stopped", "move_e", "move_h", "bound", "decay", "x_forster", "x_dexter", "x_isc", "x_risc",
                                         "recombination", "emission", "injection")}}


X_L, X_TRAJ_PER_GPU, X_EVENTS, X_CUTOFF = 128, 1_000_000, 100, 4.0


def exciton_traffic(events_per_launch):
    """DRAM bytes of one exciton-kernel launch, scaled from the committed ncu capture (per event)."""
    return measured_traffic("hfx_run_kernel", events_per_launch)


def exciton_rate(rank, local_rank, world, lattice=X_L, realisations_total=1, label="north_star", cpu=False):
    """The exciton path (own model, parity unpinned: the reference has no exciton dynamics).
    north_star: the north_star's target configuration — 128^3 lattice, 10^6 concurrent single-exciton
    trajectories per GPU, cutoff 4 lattice constants (256 transfer channels + 3 on-site), rejection-free
    selection.  disorder_sweep: BASELINE configs[4] — 256^3, 64 disorder realisations sharded over the ranks,
    10^6 trajectories per GPU.  Events/s over all ranks and the algorithmic-bytes roofline fraction of
    SURVEY.md 8(d)."""
    import torch
    from hyperfluorescence_b200 import _native as nat
    n_real = max(1, realisations_total // world)
    sim = nat.Sim((lattice,) * 3, PROPS, FIELD, 1, 1, n_trajectories=X_TRAJ_PER_GPU, n_realisations=n_real, seed=SEED,
                  first_trajectory=rank * X_TRAJ_PER_GPU, first_realisation=rank * n_real, device=local_rank)
    sim.generate_lattice()
    sim.x_configure(nat.x_params_from_dict(dict(cutoff=X_CUTOFF)))
    sim.x_spawn()
    for _ in range(2):
        sim.x_run(X_EVENTS)
    sim.sync()
    sim.run_timing()
    for _ in range(3):
        sim.x_run(X_EVENTS)
    ms, n = sim.run_timing()
    launch = sim.x_launch_info()
    tab = sim.x_tables()
    m = int(tab["g6"].size)
    from hyperfluorescence_b200 import distributed as hd
    t = hd.all_reduce_named(sim.x_tallies())                                    # the one collective of the path
    cpu_line = None
    if cpu and world == 1 and rank == 0:
        cpu_line = exciton_cpu_sample(sim, lattice, tab)
    e2e = None
    if label == "north_star":
        # end to end through the C-ABI with HOST buffers every step: lattice arrays from pinned host memory ->
        # hf_upload_lattice (H2D), hf_x_configure (tables + Boltzmann weights), hf_x_spawn, hf_x_run, hf_x_read_tallies (D2H)
        host = sim.download_lattice(0)
        pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in host.items()}
        xp = nat.x_params_from_dict(dict(cutoff=X_CUTOFF))
        def e2e_step():
            sim.upload_lattice(0, pinned["species"], pinned["homo"], pinned["lumo"], pinned["s1"], pinned["t1"])
            sim.x_configure(xp)
            sim.x_spawn()
            sim.x_run(X_EVENTS)
            return sim.x_tallies()
        e2e_step()
        sim.sync()
        t0 = time.perf_counter()
        for _ in range(2):
            e2e_step()
        sim.sync()
        dt = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist
            v = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            dt = float(v.item())
        e2e = {"value": world * X_TRAJ_PER_GPU * X_EVENTS * 2 / dt, "unit": "events/s",
               "h2d_bytes_per_step": int(sum(v.nbytes for v in pinned.values())), "d2h_bytes_per_step": 16 * 8, "steps": 2,
               "path": "hf_upload_lattice(pinned host) + hf_x_configure + hf_x_spawn + hf_x_run + hf_x_read_tallies"}
    sim.close()
    if world > 1:
        import torch.distributed as dist
        v = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms = float(v.item())
    events = world * X_TRAJ_PER_GPU * X_EVENTS * n
    bytes_per_event = m * (1 + 8) + (m + 7) // 8 + 2 * 21 + 16                  # SURVEY.md 8(d), Path X
    peak, _ = hbm_peak()
    rate = events / (ms * 1e-3)
    decayed = max(1, t["generated"] - world * X_TRAJ_PER_GPU)
    l2_resident = n_real * lattice ** 3 * 8 < 60e6
    out = {"metric": "KMC exciton events/sec (box), exciton path", "value": rate, "unit": "events/s",
           "config": {"workload": label, "lattice": [lattice] * 3, "realisations_per_gpu": n_real,
                      "trajectories_per_gpu": X_TRAJ_PER_GPU, "cutoff": X_CUTOFF,
                      "channels": m + 3, "events_per_trajectory_per_launch": X_EVENTS,
                      "weights_layout": "4x4x1 bricks" if launch["brick"] else "row-major", "trajectories_per_warp": launch["block"],
                      "model": "own (parity unpinned): DESIGN.md section 9, oracle/exciton_oracle.c"},
           "roofline": {"bound": "hbm", "bytes_per_event": bytes_per_event,
                        "achieved": rate * bytes_per_event / world / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": rate * bytes_per_event / world / 1e9 / peak, "kernel": "hfx_run_kernel",
                        "traffic": exciton_traffic(X_TRAJ_PER_GPU * X_EVENTS) if l2_resident else None,
                        "note": ("the 2 x 17 MB of Boltzmann weights of a 128^3 lattice are L2-resident (ncu: L2 hit 99.9 %, "
                                 "DRAM ~1.6 B/event), so the algorithmic-bytes fraction is not an HBM utilisation; the kernel is "
                                 "bound by instruction issue and the L1 data pipe (`issue`; profiles/r02h_*, r01f_*)") if l2_resident else
                                ("weights exceed the L2: row-major with 32 trajectories per warp the kernel moved 7.9 KB/event from DRAM "
                                 "at 4.2 TB/s (profiles/r01g_*); the adaptive warp batch trades that for L2 hits (profiles/r01h_*) and "
                                 "the 4 x 4 x 1 brick layout cuts the lines a sphere touches by 36 % (profiles/r02g_brick_product.txt)")},
           "issue": issue_roofline(rate / world, "hfx_run_kernel") if l2_resident else None,
           "photons_per_exciton": (t["rad_tadf"] + t["rad_fluo"] + t["rad_host"]) / decayed}
    if cpu_line:
        out["cpu_baseline"] = cpu_line
    if e2e:
        out["e2e"] = e2e
    return out


def exciton_cpu_sample(sim, lattice, tab):
    """The exciton specification (oracle/exciton_oracle.c, POSIX threads over trajectories) on this box's host
    cores, on the device's own lattice: a bounded sample of the same workload."""
    from hyperfluorescence_b200 import _native as nat
    from oracle import exciton_oracle as xo
    lat = sim.download_lattice(0)
    ws1, wt1 = sim.x_weights(0)
    params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=X_CUTOFF)
    o = xo.ExcitonOracle((lattice,) * 3, lat["species"], ws1, wt1, params, tab)
    cores = os.cpu_count() or 1
    n_traj = 1500 * cores
    o.run_batch(SEED, 0, max(cores, 64), 10, n_threads=cores)                  # warm the page cache / threads
    t0 = time.perf_counter()
    o.run_batch(SEED, 0, n_traj, X_EVENTS, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n_traj * X_EVENTS / dt, "unit": "events/s", "cores": cores, "kind": "port",
            "sample": f"{n_traj} of the {X_TRAJ_PER_GPU} trajectories x {X_EVENTS} events on the same {lattice}^3 lattice, {cores} threads"}


XD_L, XD_DEVICES_PER_GPU, XD_EXCITONS, XD_EVENTS = 128, 4096, 256, 200


def dense_cpu_sample(sim, tab):
    """The high-density specification (oracle/exciton_dense_oracle.c, POSIX threads over devices) on this box's
    host cores, on the device's own lattice: a bounded sample of the same workload (fill included)."""
    from hyperfluorescence_b200 import _native as nat
    from oracle import exciton_oracle as xo
    lat = sim.download_lattice(0)
    ws1, wt1 = sim.x_weights(0)
    params = dict(nat.x_params_to_dict(nat.x_default_params()), cutoff=X_CUTOFF)
    o = xo.DenseExcitonOracle((XD_L,) * 3, lat["species"], ws1, wt1, params, tab)
    cores = os.cpu_count() or 1
    n_dev = XD_DEVICES_PER_GPU // 4
    t0 = time.perf_counter()
    o.run_devices(SEED, 0, n_dev, XD_EXCITONS, XD_EVENTS, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n_dev * XD_EVENTS / dt, "unit": "events/s", "cores": cores, "kind": "port",
            "sample": f"{n_dev} of the {XD_DEVICES_PER_GPU} devices x {XD_EXCITONS} excitons x {XD_EVENTS} events "
                      f"(initial fill and totals included) on the same {XD_L}^3 lattice, {cores} threads"}


def dense_rate(rank, local_rank, world, cpu=False):
    """BASELINE configs[3]: high-density excitons on a 128^3 lattice with occupancy blocking and
    exciton-exciton annihilation (own model, parity unpinned: DESIGN.md section 10).  4096 devices per GPU x
    256 interacting excitons each (10^6 excitons per GPU); one event = one channel of one exciton."""
    import torch
    from hyperfluorescence_b200 import _native as nat
    sim = nat.Sim((XD_L,) * 3, PROPS, FIELD, 1, 1, n_trajectories=XD_DEVICES_PER_GPU, seed=SEED,
                  first_trajectory=rank * XD_DEVICES_PER_GPU, device=local_rank)
    sim.generate_lattice()
    sim.x_configure(nat.x_params_from_dict(dict(cutoff=X_CUTOFF)))
    sim.xd_configure(XD_EXCITONS)
    sim.xd_spawn()
    sim.xd_run(XD_EVENTS)
    sim.sync()
    sim.run_timing()
    for _ in range(3):
        sim.xd_run(XD_EVENTS)
    ms, n = sim.run_timing()
    from hyperfluorescence_b200 import distributed as hd
    t = hd.all_reduce_named(sim.xd_tallies())                                   # the one collective of the path
    cpu_line = dense_cpu_sample(sim, sim.x_tables()) if (cpu and world == 1 and rank == 0) else None
    # end to end through the C-ABI with HOST buffers every step: lattice from pinned host memory -> hf_upload_lattice,
    # hf_x_configure (tables + Boltzmann weights), hf_xd_configure, hf_xd_spawn (fill), hf_xd_run, tallies + all-reduce
    host = sim.download_lattice(0)
    pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in host.items()}
    xp = nat.x_params_from_dict(dict(cutoff=X_CUTOFF))
    def e2e_step():
        sim.upload_lattice(0, pinned["species"], pinned["homo"], pinned["lumo"], pinned["s1"], pinned["t1"])
        sim.x_configure(xp)
        sim.xd_configure(XD_EXCITONS)
        sim.xd_spawn()
        sim.xd_run(XD_EVENTS)
        return hd.all_reduce_named(sim.xd_tallies())
    e2e_step()
    sim.sync()
    t0 = time.perf_counter()
    for _ in range(2):
        e2e_step()
    sim.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    e2e = {"value": world * XD_DEVICES_PER_GPU * XD_EVENTS * 2 / e2e_s, "unit": "events/s",
           "h2d_bytes_per_step": int(sum(v.nbytes for v in pinned.values())), "d2h_bytes_per_step": 24 * 8, "steps": 2,
           "path": "hf_upload_lattice(pinned host) + hf_x_configure + hf_xd_configure + hf_xd_spawn (fill) + hf_xd_run + tallies"}
    sim.close()
    if world > 1:
        import torch.distributed as dist
        v = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms = float(v.item())
    lost = max(1, sum(v for k, v in t.items() if k.startswith(("rad_", "nonrad_", "ann_"))))
    return {"metric": "KMC exciton events/sec (box), high-density devices", "value": world * XD_DEVICES_PER_GPU * XD_EVENTS * n / (ms * 1e-3),
            "unit": "events/s",
            "config": {"lattice": [XD_L] * 3, "devices_per_gpu": XD_DEVICES_PER_GPU, "excitons_per_device": XD_EXCITONS,
                       "cutoff": X_CUTOFF, "events_per_device_per_launch": XD_EVENTS, "kernel": "hfxd_run_kernel",
                       "model": "own (parity unpinned): DESIGN.md section 10, oracle/exciton_dense_oracle.c"},
            "annihilation_fraction": (t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"]) / lost, "e2e": e2e,
            **({"cpu_baseline": cpu_line} if cpu_line else {})}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ga", action="store_true", help="skip the GA fitness-sweep, exciton-path, disorder-sweep and high-density legs")
    args = ap.parse_args()
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.warmup < 3 and args.impl == "ours" and not args.no_cpu:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
