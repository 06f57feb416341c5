"""NCCL transport of the domain decomposition, run under torchrun on >= 2 GPUs (tests/test_gpu_dist.py::test_nccl_ranks
launches it when the box has them):  every rank loads a random share of a Sedov sphere, three driver iterations run
decomposed over NCCL, and rank 0 checks the gathered result against its own single-GPU run of the same input
(neighbour counts equal, fp64 fields within 1e-10)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from flecsph_b200 import capi
from common import make_system, case, rel_err

FIELDS = ["pos", "vel", "rho", "P", "u", "e", "cs", "acc", "dedt", "dudt", "gacc", "gpot", "h", "vsig", "dt"]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # FSPH_CASE=collapse: newtonian configuration (hydro + FMM gravity over the gathered global tree)
    which = os.environ.get("FSPH_CASE", "sedov")
    f, par, dim = case(which, nx=int(os.environ.get("FSPH_NX", "40")), fmm=1)
    n = len(f["mass"])
    owner = np.random.default_rng(11).integers(0, world, n)
    mine = {k: np.ascontiguousarray(v[owner == rank]) for k, v in f.items()}
    p = capi.default_params(dim, **par)
    s = capi.BodySystem(p, backend="cuda", device=local)
    s.load_bodies(mine)
    s.set_time(p.initial_dt, 0.0, 0)
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init_nccl(uid[0], rank, world)
    for it in range(3):
        s.step(initial=(it == 0))
    ghosts = s.nghosts()
    snap = s.snapshot(FIELDS + ["id", "nneigh"])
    gathered = [None] * world
    dist.gather_object((snap, ghosts, s.get_time()), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        S = make_system("cuda", f, par, dim)
        for it in range(3):
            S.step(initial=(it == 0))
        ref = S.snapshot(FIELDS + ["id", "nneigh"])
        o = np.argsort(ref["id"])
        ids = np.concatenate([g[0]["id"] for g in gathered])
        o2 = np.argsort(ids)
        assert np.array_equal(ref["id"][o], ids[o2]), "bodies lost or duplicated"
        assert sum(g[1] for g in gathered) > 0, "no ghosts were exchanged"
        assert all(g[2] == S.get_time() for g in gathered), "time / dt differ"
        for nm in FIELDS + ["nneigh"]:
            got = np.concatenate([g[0][nm] for g in gathered])[o2]
            e = rel_err(got, ref[nm][o])
            tol = 0 if nm == "nneigh" else 1e-10
            print("%-7s max rel err %.3g" % (nm, e))
            if e > tol:
                ok = False
        print("owned per rank:", [len(g[0]["id"]) for g in gathered], "ghosts per rank:", [g[1] for g in gathered])
        print("NCCL DECOMPOSITION OK" if ok else "NCCL DECOMPOSITION MISMATCH")
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
