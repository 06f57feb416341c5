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
HBM; `e2e` is the same step driven through the C ABI from pinned HOST buffers (fsph_step_pipelined): every step its input
state is copied host -> device and its results device -> host, all inside the timed region; the results leave on a side
stream as soon as a column is final, and the uploads of the next step -- its inputs are the buffers this step filled, each
upload ordered on the device after the download of the same buffer -- overlap the tail of this step's downloads.
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
               "tree": 90.0, "range_keys": 64.0, "elementwise": 404.0, "timestep": 88.0,
               # the search's 52 B split over its two big kernels: the walk reads the tree summaries (16 B per
               # particle), the pair kernel reads x, h of every body (32 B) and writes its neighbour count (4 B)
               "search:walk": 16.0, "search:pairs": 36.0}
KERNEL_OF = {"search:pairs": "k_block_pairs", "search:walk": "k_block_walk", "density": "k_sweep<density>", "acceleration": "k_sweep<acceleration + signal speed>",
             "dedt": "k_sweep<de/dt>", "dudt": "k_sweep<du/dt>"}
PAIR_FLOPS = {"density": 43.0, "acceleration": 77.0, "dedt": 90.0, "dudt": 82.0}   # SURVEY.md 8d, per accepted pair
FP64_NOMINAL_TFLOPS = 37.2   # nominal B200 vector fp64 (148 SM x 64 DFMA/clk x 2 x 1.965 GHz); the measured DFMA rate of
                             # the GPU the run uses (fsph_measure_fp64_tflops) is the denominator of the fp64 fractions


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


def make_sphere_system(kind, nx, device, slab=None, allsum=None, **par_over):
    """A Sedov / Noh / collapse FCC sphere generated on the device (fsph_generate_lattice + fsph_id_fill: the same bodies
    as flecsph_b200/initial_data.py, bit for bit).  Returns (system, n_local)."""
    from flecsph_b200 import capi, initial_data as idata
    par = {"sedov": idata.sedov_3d, "noh": idata.noh_3d, "collapse": idata.collapse_3d}[kind](nx=4)[1]
    par.update(par_over)
    p = capi.default_params(3, **par)
    s = capi.BodySystem(p, backend="cuda", device=device)
    n_local, _ = idata.sphere_on_device(s, kind, nx, slab=slab, allsum=allsum)
    s.set_time(p.initial_dt, 0.0, 0)
    return s, n_local


def time_steps(s, steps, warmup, barrier=None):
    """initial iteration + warm-up + `steps` timed iterations; returns (device ms per step, phases of the timed steps)."""
    s.step(initial=True)
    for _ in range(warmup):
        s.step()
    if barrier:
        barrier()
    tot, acc = 0.0, {}
    for _ in range(steps):
        s.step()
        ph = s.phase_times()
        tot += ph.get("total", 0.0)
        for k, v in ph.items():
            acc[k] = acc.get(k, 0.0) + v
    if barrier:
        barrier()
    return tot / steps, {k: round(v / steps, 2) for k, v in acc.items() if k != "total" and not k.startswith("dd:")}


# state a non-initial step reads / results it produces (column, components)
def e2e_columns(capi):
    up = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1),
          (capi.COL_E, 1), (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1)]
    down = [(capi.COL_POS, 3), (capi.COL_VEL, 3), (capi.COL_ACC, 3), (capi.COL_GACC, 3), (capi.COL_MASS, 1), (capi.COL_H, 1),
            (capi.COL_E, 1), (capi.COL_DEDT, 1), (capi.COL_CS, 1), (capi.COL_ID, 1), (capi.COL_RHO, 1), (capi.COL_P, 1),
            (capi.COL_U, 1), (capi.COL_VSIG, 1), (capi.COL_DT, 1)]
    return up, down


def _cpu_worker(nx, steps, warmup, barrier, q):
    """One replica of the reference's own step loop (unmodified headers, -Ofast build of oracle/_ref) on one host core."""
    from flecsph_b200 import capi, initial_data as idata
    f, par = idata.sedov_3d(nx=nx)
    p = capi.default_params(3, **par)
    s = capi.BodySystem(p, backend="ref_fast")
    s.load_bodies(f)
    s.set_time(p.initial_dt, 0.0, 0)
    n = s.nbodies()
    s.step(initial=True)
    for _ in range(warmup):
        s.step()
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    t = time.perf_counter() - t0
    s.close()
    q.put((n, t))


def run_cpu_reference(nx, steps, warmup, nproc=None):
    """The reference hot path is single-threaded per MPI rank and there is no MPI in this image, so "all the host
    cores" means one independent replica of the same Sedov sample per core, started together (ranks without
    communication: an upper bound for what an MPI run of the reference could reach on this host)."""
    import multiprocessing as mp
    from flecsph_b200 import capi
    if not os.path.exists(capi.library_path("ref_fast")):
        return None
    nproc = nproc or max(1, min(os.cpu_count() or 1, 32))
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(nproc), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(nx, steps, warmup, barrier, q)) for _ in range(nproc)]
    for pr in procs:
        pr.start()
    res = [q.get() for _ in procs]
    for pr in procs:
        pr.join()
    n, t = res[0][0], max(r[1] for r in res)
    return {"value": nproc * n * steps / t, "unit": "particle-updates/s", "cores": nproc, "kind": "reference",
            "sample": "Sedov 3-D FCC sphere lattice_nx=%d (%d particles) per replica, %d independent replicas (one per host core; the "
                      "reference hot path is single-threaded per MPI rank and there is no MPI in this image), %d non-initial steps "
                      "after the initial one + %d warm-up, -Ofast; %.3g updates/s per core" % (nx, n, nproc, steps, warmup, n * steps / t),
            "ms_per_step": 1e3 * t / steps, "n": n, "per_core": n * steps / t}


def reference_nx(steps, warmup, budget_s=270.0, rate_per_core=6.0e4):
    """Largest lattice_nx (<= 112, i.e. ~1.0 M bodies per replica) whose --steps/--warmup run fits the time budget."""
    from flecsph_b200 import initial_data as idata
    nmax = budget_s * rate_per_core / (steps + warmup + 1.5)
    nx = 112
    while nx > 40 and 0.74 * nx ** 3 > nmax:
        nx -= 2
    return nx


def parity_record(capi, dist, torch, rank, world, local_rank, nx=143):
    """Decomposed-vs-one-rank check over NCCL on a Sedov sphere of ~2.1 M bodies, outside the timed region: every rank
    loads a random share, two driver iterations run decomposed, and every rank compares the bodies it ends up owning
    with its own single-GPU run of the same input (matched by id).  Returns {bodies, max_rel, nneigh_mismatch}."""
    from flecsph_b200 import initial_data as idata
    f, par = idata.sedov_3d(nx=nx)
    n = len(f["mass"])
    owner = np.random.default_rng(11).integers(0, world, n)
    mine = {k: np.ascontiguousarray(v[owner == rank]) for k, v in f.items()}
    p = capi.default_params(3, **par)
    s = capi.BodySystem(p, backend="cuda", device=local_rank)
    s.load_bodies(mine)
    s.set_time(p.initial_dt, 0.0, 0)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init_nccl(uid[0], rank, world)
    one = capi.BodySystem(p, backend="cuda", device=local_rank)
    one.load_bodies(f)
    one.set_time(p.initial_dt, 0.0, 0)
    for it in range(2):
        s.step(initial=(it == 0))
        one.step(initial=(it == 0))
    names = ["pos", "vel", "rho", "P", "u", "cs", "acc", "dedt", "h", "vsig", "dt"]
    ids, rid = s.get(capi.COL_ID), one.get(capi.COL_ID)
    order = np.argsort(rid)
    at = order[np.searchsorted(rid[order], ids)]
    worst = 0.0
    for nm in names:
        col = capi.COL_NAMES.index(nm)
        ref = one.get(col)
        got = s.get(col)
        scale = np.abs(ref).max()
        worst = max(worst, float(np.abs(got - ref[at]).max() / max(scale, 1e-300)))
    mism = int((s.get(capi.COL_NNEIGH) != one.get(capi.COL_NNEIGH)[at]).sum())
    tt = torch.tensor([worst, float(mism), float(len(ids))], dtype=torch.float64, device="cuda")
    mx = tt.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(tt)
    s.close()
    one.close()
    return {"bodies": n, "owned_sum": int(tt[2]), "max_rel": float(mx[0]), "nneigh_mismatch": int(tt[1]), "ranks": world,
            "what": "Sedov nx=%d, random initial distribution, 2 driver iterations over NCCL vs the same input on one GPU; "
                    "max over %s of max|a-b|/max|b|, and the number of bodies whose neighbour count differs" % (nx, ",".join(names))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles", type=float, default=64e6, help="target particle count per GPU")
    ap.add_argument("--cpu-nx", type=int, default=0, help="lattice_nx of the CPU sample (0: chosen from the time budget)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e with the uploads issued one by one before fsph_step_streamed (round-2 path) instead of fsph_step_pipelined")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations / scaling sub-records")
    ap.add_argument("--host-ids", action="store_true", help="generate the initial data on the host (numpy) instead of on the device")
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
        nx = a.cpu_nx or reference_nx(a.steps, a.warmup)
        r = run_cpu_reference(nx, a.steps, a.warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libflecsph_ref_3d_fast.so was not built/shipped"}))
            return 0
        config["particles"] = r["n"] * r["cores"]
        config["particles_per_replica"] = r["n"]
        config["lattice_nx"] = nx
        line = {"impl": "reference", "metric": "particle-updates/sec per SPH step (Sedov 3D)", "value": r["value"],
                "unit": "particle-updates/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        sys.stdout.write("\n")   # the reference's own logging may have left an unterminated line
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from flecsph_b200 import capi, initial_data as idata
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

    def allsum(x):
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t[0])

    def join(s):
        if world > 1:
            uid = [capi.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            s.comm_init_nccl(uid[0], rank, world)

    # measured denominators of this GPU: DFMA rate and copy bandwidth
    lib = capi.load(capi.PRODUCT_LIB)
    fp64_peak = float(lib.fsph_measure_fp64_tflops(local_rank, 3))
    copy_gbs = float(lib.fsph_measure_copy_gbs(local_rank, 2 << 30, 3))

    t_setup0 = time.perf_counter()
    nx = idata.nx_for_particles(a.particles * world)
    slab = (rank, world) if world > 1 else None
    if a.host_ids:
        f, par = idata.sedov_3d(nx=nx, slab=slab, allsum=allsum if world > 1 else None)
        s = make_cuda_system(f, par, local_rank)
        n_local = len(f["mass"])
        del f
    else:
        s, n_local = make_sphere_system("sedov", nx, local_rank, slab=slab, allsum=allsum if world > 1 else None)
    n_total = allsum(n_local)
    n = n_total // world
    join(s)
    if world == 1:
        config.update({"particles_per_gpu": n, "lattice_nx": nx, "parallelism": "single GPU"})
    else:
        # weak scaling: ONE Sedov sphere of world x 64 M bodies, Morton domain decomposition over the GPUs of the box
        # (sample-sort all-to-all, locally-essential-tree halo, ghost refresh: NCCL over NVLink; fsph_dist.cu).
        # Every rank generates a slab of lattice layers; the first update_iteration re-partitions by key.
        config.update({"particles_total": n_total, "particles_per_gpu": n, "lattice_nx": nx,
                       "parallelism": "Morton-key domain decomposition over %d GPUs (one rank per GPU): sample-sort all-to-all, "
                                      "locally-essential-tree halo, ghost refresh before every sweep, NCCL" % world})
    config["initial_data"] = "host numpy" if a.host_ids else "device (fsph_generate_lattice + fsph_id_fill), %.1f s" % (time.perf_counter() - t_setup0)
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
        config["dd_phases_ms_per_rank"] = [{k: round(v, 2) for k, v in d.items() if k.startswith("dd:")} for d in per_rank]
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
        # results leave through fsph_step_streamed: every column goes to its pinned buffer on a side stream as soon as
        # the iteration has written it for the last time (ids / masses / positions after update_iteration, ...), the call
        # returns when the last byte has arrived
        plan = [(col, c, bufs[col][c].data_ptr()) for col, nc in down for c in range(nc)]
        plan_in = [(col, c, bufs[col][c].data_ptr()) for col, nc in up for c in range(nc)]
        if a.e2e_serial:
            for _ in range(a.steps):
                upload()
                if s.nbodies() > cap:
                    raise SystemExit("e2e staging buffers too small for this rank's share")
                s.step_streamed(plan)
        else:
            # fsph_step_pipelined: the buffers step k filled are the inputs of step k+1; every byte goes host -> device ->
            # host every step, the uploads of step k+1 (each ordered on the device after step k's copy of the same
            # buffer) run while the tail of step k's results is still on its way down
            for _ in range(a.steps):
                if s.nbodies() > cap:
                    raise SystemExit("e2e staging buffers too small for this rank's share")
                s.step_pipelined(plan_in, plan)
            s.stream_wait()
        config["e2e_path"] = "fsph_step_streamed after serial uploads" if a.e2e_serial else "fsph_step_pipelined"
        barrier()
        te = time.perf_counter() - t0
        tte = torch.tensor([te], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tte, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * a.steps / float(tte[0]), "unit": "particle-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tte[0]) / a.steps}
        del bufs
    s.close()

    # ---- sub-records outside the timed region: the other BASELINE.json configurations, strong scaling, N-rank parity ----
    configs = {}
    if not a.no_configs:
        def sub(kind, target, label, **par_over):
            nxs = idata.nx_for_particles(target)
            ss, nl = make_sphere_system(kind, nxs, local_rank, slab=slab, allsum=allsum if world > 1 else None, **par_over)
            ntot = allsum(nl)
            join(ss)
            ms, ph = time_steps(ss, 3, 2, barrier)
            tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ss.close()
            configs[label] = {"bodies": ntot, "lattice_nx": nxs, "ms_per_step": float(tm[0]), "updates_per_s": ntot / (float(tm[0]) * 1e-3),
                              "phases_ms_rank0": ph, "n_gpus": world}
        if world == 1:
            # BASELINE.json configs[3] and [4] at full size on one GPU (KH: periodic box, eta 1.5; collapse: quintic, FMM)
            # a sub-record that fails (these are not the headline) is reported as such instead of taking the line with it
            def guarded_sub(label, fn):
                try:
                    fn()
                except Exception as e:      # noqa: BLE001
                    configs[label] = {"error": str(e)[:300]}

            def kh():
                # app/id_generators/KH/main.cc with data/KH_3d.par semantics (HCP blocks, density ratio 2), lattice_nx for ~32 M bodies
                f, par = idata.kh_3d_reference(nx=328)
                ss = make_cuda_system(f, par, local_rank)
                nb = len(f["mass"])
                del f
                ms, ph = time_steps(ss, 3, 2)
                configs["kelvin_helmholtz_3d"] = {"bodies": nb, "bodies_with_periodic_copies": ss.nbodies(), "lattice_nx": 328, "ms_per_step": ms,
                                                  "updates_per_s": nb / (ms * 1e-3), "phases_ms_rank0": ph, "n_gpus": 1}
                ss.close()
            guarded_sub("kelvin_helmholtz_3d", kh)
            guarded_sub("collapse_newtonian_3d", lambda: sub("collapse", 32e6, "collapse_newtonian_3d"))
            guarded_sub("sedov_3d_16M", lambda: sub("sedov", 16e6, "sedov_3d_16M"))
        else:
            sub("sedov", 64e6, "sedov_3d_64M_strong_scaling")     # the metric as BASELINE.json quotes it: 64 M bodies in total
            sub("noh", 128e6, "noh_3d_128M")
            sub("collapse", 32e6, "collapse_newtonian_3d")
            configs["parity"] = parity_record(capi, dist, torch, rank, world, local_rank)

    # tear down on every rank in the same order (ncclCommDestroy of the library's communicators, then torch's)
    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, how = measured_peaks()
    phases = {k: v / a.steps for k, v in phase_acc.items() if k != "total"}
    # the dominant KERNEL: each of these phases is one launch per step, timed with its own pair of CUDA events
    dom = max((k for k in phases if k in ("search:pairs", "search:walk", "density", "acceleration", "dedt", "dudt")), key=lambda k: phases[k])
    dom_ms = phases[dom]
    ach = n * PHASE_BYTES[dom] / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant phase from the committed ncu capture of this exact workload (profiles/), else null
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_traffic_64M.json")
    if world == 1 and nx == 444 and os.path.exists(tp):
        tj = json.load(open(tp))
        if dom in tj.get("phases", {}):
            traffic, traffic_src = tj["phases"][dom], "profiles/r2_traffic_64M.json (%s)" % tj["how"]
    roofline = {"bound": "hbm", "kernel": KERNEL_OF.get(dom, dom), "phase": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": how, "copy_gbs_this_gpu": copy_gbs,
                "algorithmic_bytes_per_particle": PHASE_BYTES[dom], "ms_per_launch": dom_ms,
                "note": "the dominant phase is fp64-issue / latency bound, not HBM bound (profiles/r2_ncu_*.md): see fp64 (fractions of the "
                        "measured DFMA rate) and roofline_step"}
    step_ach = n * ALGO_BYTES_PER_UPDATE / (t_dev / a.steps) / 1e9
    roofline_step = {"bound": "hbm", "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                     "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE}
    fp64 = {"peak_tflops_measured": fp64_peak, "peak_tflops_nominal": FP64_NOMINAL_TFLOPS,
            "how": "fsph_measure_fp64_tflops: 8 independent DFMA chains per thread, 8 CTAs of 256 threads per SM, best of 3"}
    for k in ("density", "acceleration", "dedt", "dudt"):
        if k in phases and phases[k] > 0:
            tf = n * nn_mean * PAIR_FLOPS[k] / (phases[k] * 1e-3) / 1e12
            fp64[k] = {"tflops": tf, "frac_of_measured": tf / fp64_peak if fp64_peak > 0 else None, "frac_of_nominal_37.2": tf / FP64_NOMINAL_TFLOPS}
    cpu = None
    if not a.no_cpu and world == 1:
        r = run_cpu_reference(a.cpu_nx or 90, 2, 0)
        if r is not None:
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": "particle-updates/sec per SPH step (Sedov 3D)", "value": value, "unit": "particle-updates/s", "n_gpus": world,
            "steps": a.steps, "warmup": warmup, "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "roofline_step": roofline_step, "fp64": fp64, "cpu_baseline": cpu,
            "phases_ms": phases, "mean_neighbors": nn_mean, "wall_ms_per_step": 1e3 * t_wall / a.steps, "configs": configs}
    if "parity" in configs:
        line["parity"] = configs.pop("parity")
    sys.stdout.write("\n")   # the reference's own logging (cpu_baseline leg) may have left an unterminated line
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
