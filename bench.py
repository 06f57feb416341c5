#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full SPH step (Sedov 3-D), the metric of BASELINE.json.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation of the same
                                                           step loop (oracle/_ref, -Ofast build) on the host cores

A "step" is one non-initial iteration of the hydro driver loop (app/drivers/hydro/main_driver.cc:175-277): kicks,
drift, range, keys, sort, tree, neighbour search, density / acceleration / de/dt sweeps, h update, dt.  Workload at
N=1: Sedov blast wave, FCC lattice in the unit sphere scaled to ~64 M particles (BASELINE.json configs[1]); inputs
(~100 B/particle of live state, 6+ GB) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
`value` is timed with CUDA events inside the library on its own stream (fsph_phase_times "total"), inputs resident in
HBM; `e2e` is the same step driven through the C ABI from pinned HOST buffers with the state uploaded before and the
results downloaded after every step inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_UPDATE = 1386.0   # hydro thermokinetic path, fp64 SoA, every column read/written once per phase
# per-phase algorithmic bytes per particle; the neighbour search is the traversal part of the
# density sweep in the reference (x,h 32 B read + tree summaries ~16 B read + neighbour count 4 B written)
PHASE_BYTES = {"density": 120.0, "acceleration": 112.0, "dedt": 120.0, "dudt": 120.0, "search": 52.0, "sort": 416.0,
               "tree": 90.0, "range_keys": 64.0, "elementwise": 404.0, "timestep": 88.0}
PAIR_FLOPS = {"density": 43.0, "acceleration": 77.0, "dedt": 90.0, "dudt": 82.0}   # SURVEY.md 8d, per accepted pair
FP64_PEAK_TFLOPS = 37.2   # nominal B200 vector fp64 (148 SM x 64 DFMA/clk x 2 x 1.965 GHz); not in MEASURED_PEAKS.json


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        # the first samples can predate the load; use the upper half as "under load"
        sm_sorted = sorted(sm)
        med = sm_sorted[len(sm_sorted) // 2] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def sedov_fields(target_particles):
    from flecsph_b200 import initial_data as idata
    nx = idata.nx_for_particles(target_particles)
    f, par = idata.sedov_3d(nx=nx)
    return f, par, nx


def make_cuda_system(f, par, device):
    from flecsph_b200 import capi
    p = capi.default_params(3, **par)
    s = capi.BodySystem(p, backend="cuda", device=device)
    s.load_bodies(f)
    s.set_time(p.initial_dt, 0.0, 0)
    return s


# state a non-initial step reads / results it produces (column, components)
def e2e_columns(capi):
    up = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1),
          (capi.COL_E, 1), (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1)]
    down = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1),
            (capi.COL_E, 1), (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1), (capi.COL_RHO, 1), (capi.COL_P, 1),
            (capi.COL_U, 1), (capi.COL_VSIG, 1), (capi.COL_DT, 1)]
    return up, down


def run_cpu_reference(nx, steps, warmup):
    """The reference's own step loop (unmodified headers, -Ofast build of oracle/_ref) on one host core."""
    from flecsph_b200 import capi, initial_data as idata
    path = capi.library_path("ref_fast")
    kind = "reference"
    if not os.path.exists(path):
        return None
    f, par = idata.sedov_3d(nx=nx)
    p = capi.default_params(3, **par)
    s = capi.BodySystem(p, backend="ref_fast")
    s.load_bodies(f)
    s.set_time(p.initial_dt, 0.0, 0)
    n = s.nbodies()
    s.step(initial=True)
    for _ in range(warmup):
        s.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    t = time.perf_counter() - t0
    s.close()
    return {"value": n * steps / t, "unit": "particle-updates/s", "cores": 1, "kind": kind,
            "sample": "Sedov 3-D FCC sphere lattice_nx=%d (%d particles), %d non-initial steps after the initial one + %d warm-up, "
                      "1 rank (the reference hot path is single-threaded per MPI rank; no MPI in this image), -Ofast" % (nx, n, steps, warmup),
            "ms_per_step": 1e3 * t / steps, "n": n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", type=float, default=64e6, help="target particle count per GPU")
    ap.add_argument("--cpu-nx", type=int, default=64, help="lattice_nx of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3) if a.impl == "cuda" else a.warmup
    config = {"workload": "Sedov blast wave 3-D (hydro_3d, data/sedov_nx20.par semantics), FCC lattice in the unit sphere, "
                          "Wendland C4, eta=1.2, variable h, adaptive dt, thermokinetic de/dt path",
              "l2": "inputs (>= 100 B/particle live state) exceed the 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = run_cpu_reference(a.cpu_nx, a.steps, a.warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libflecsph_ref_3d_fast.so was not built/shipped"}))
            return 0
        config["particles"] = r["n"]
        line = {"impl": "reference", "metric": "particle-updates/sec per SPH step (Sedov 3D)", "value": r["value"],
                "unit": "particle-updates/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from flecsph_b200 import capi
    if os.environ.get("FSPH_BENCH_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["FSPH_BENCH_TRACE"]), repeat=True, file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path in the product)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        f, par, nx = sedov_fields(a.particles)
        n = n_total = len(f["mass"])
        s = make_cuda_system(f, par, local_rank)
        config.update({"particles_per_gpu": n, "lattice_nx": nx, "parallelism": "single GPU"})
    else:
        # weak scaling: ONE Sedov sphere of world x 64 M bodies, Morton domain decomposition over the GPUs of the box
        # (sample-sort all-to-all, locally-essential-tree halo, ghost refresh: NCCL over NVLink; fsph_dist.cu).
        # Every rank generates a slab of lattice layers; the first update_iteration re-partitions by key.
        from flecsph_b200 import initial_data as idata

        def allsum(x):
            t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            return int(t[0])
        nx = idata.nx_for_particles(a.particles * world)
        f, par = idata.sedov_3d(nx=nx, slab=(rank, world), allsum=allsum)
        n_total = allsum(len(f["mass"]))
        n = n_total // world
        s = make_cuda_system(f, par, local_rank)
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init_nccl(uid[0], rank, world)
        config.update({"particles_total": n_total, "particles_per_gpu": n, "lattice_nx": nx,
                       "parallelism": "Morton-key domain decomposition over %d GPUs (one rank per GPU): sample-sort all-to-all, "
                                      "locally-essential-tree halo, ghost refresh before every sweep, NCCL" % world})
    del f
    s.step(initial=True)
    for _ in range(warmup):
        s.step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = s.launch_count()
    t_wall0 = time.perf_counter()
    total_ms = 0.0
    phase_acc = {}
    for _ in range(a.steps):
        s.step()
        ph = s.phase_times()
        total_ms += ph.get("total", 0.0)
        for k, v in ph.items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = s.launch_count() - l0
    clocks = sampler.stop()
    nghosts = s.nghosts()
    nn_sum = float(s.get(capi.COL_NNEIGH).astype(np.float64).sum())
    t_dev = total_ms / 1e3
    tt = torch.tensor([t_dev, t_wall], dtype=torch.float64, device="cuda")
    ts = torch.tensor([nn_sum, float(nghosts)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {k: v / a.steps for k, v in phase_acc.items()})
        config["phases_ms_per_rank"] = [{k: round(v, 2) for k, v in d.items() if not k.startswith("dd:")} for d in per_rank]
        for k in list(phase_acc):
            phase_acc[k] = max(d.get(k, 0.0) for d in per_rank) * a.steps
        config["ghosts_per_gpu_mean"] = float(ts[1]) / world
    t_dev, t_wall = float(tt[0]), float(tt[1])
    nn_mean = float(ts[0]) / n_total
    value = n_total * a.steps / t_dev

    # ---- e2e: host buffers, H2D of the state + D2H of the results inside the timed region ------------------
    e2e = None
    if not a.no_e2e:
        up, down = e2e_columns(capi)
        bufs = {}
        cap = n if world == 1 else int(1.25 * n) + 4096   # the owned count of a rank drifts with the re-partitioning
        for col, nc in down:
            dt_ = torch.int64 if col == capi.COL_ID else torch.float64
            bufs[col] = []
            for _ in range(nc):
                t = torch.empty(cap, dtype=dt_)
                try:
                    t = t.pin_memory()
                except RuntimeError:      # the box refused to pin more host memory: staged copies, still host buffers
                    config["e2e_host_buffers"] = "pageable (pinning failed)"
                bufs[col].append(t)

        def download():
            if s.nbodies() > cap:
                raise SystemExit("e2e staging buffers too small for this rank's share")
            for col, nc in down:
                for c in range(nc):
                    s.get_component(col, c, bufs[col][c].data_ptr())

        def upload():
            for col, nc in up:
                for c in range(nc):
                    s.set_component(col, c, bufs[col][c].data_ptr())
        download()
        h2d = sum(nc for _, nc in up) * 8 * n_total
        d2h = sum(nc for _, nc in down) * 8 * n_total
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            upload()
            s.step()
            download()
        barrier()
        te = time.perf_counter() - t0
        tte = torch.tensor([te], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tte, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * a.steps / float(tte[0]), "unit": "particle-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tte[0]) / a.steps}
        del bufs

    # tear down on every rank in the same order (ncclCommDestroy of the library's communicator, then torch's)
    s.close()
    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, how = measured_peaks()
    phases = {k: v / a.steps for k, v in phase_acc.items() if k != "total"}
    dom = max((k for k in phases if k in ("search", "density", "acceleration", "dedt", "dudt", "sort", "tree")), key=lambda k: phases[k])
    dom_ms = phases[dom]
    ach = n * PHASE_BYTES[dom] / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant phase from the committed ncu capture of this exact workload (profiles/), else null
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_search_traffic_64M.json")
    if dom == "search" and world == 1 and nx == 444 and os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj["traffic"], "profiles/r1_search_traffic_64M.json (%s)" % tj["how"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": how, "algorithmic_bytes_per_particle": PHASE_BYTES[dom], "ms_per_launch": dom_ms,
                "note": "not HBM-bound: the neighbour walk is latency-bound, the pair sweeps run with the shared-memory pipe at 73-77 % and the fp64 pipe at 47-55 % of peak (profiles/r1_ncu_*.md); see fp64 and roofline_step"}
    step_ach = n * ALGO_BYTES_PER_UPDATE / (t_dev / a.steps) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                     "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE}
    fp64 = {}
    for k in ("density", "acceleration", "dedt", "dudt"):
        if k in phases and phases[k] > 0:
            tf = n * nn_mean * PAIR_FLOPS[k] / (phases[k] * 1e-3) / 1e12
            fp64[k] = {"tflops": tf, "frac_of_nominal_37.2": tf / FP64_PEAK_TFLOPS}
    cpu = None
    if not a.no_cpu:
        r = run_cpu_reference(a.cpu_nx, 2, 0)
        if r is not None:
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": "particle-updates/sec per SPH step (Sedov 3D)", "value": value, "unit": "particle-updates/s", "n_gpus": world,
            "steps": a.steps, "warmup": warmup, "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "roofline_step": roofline_step, "fp64": fp64, "cpu_baseline": cpu,
            "phases_ms": phases, "mean_neighbors": nn_mean, "wall_ms_per_step": 1e3 * t_wall / a.steps}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
