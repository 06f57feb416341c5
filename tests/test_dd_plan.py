"""Host-side planning of the domain decomposition (fsph_dd_choose_splitters / fsph_dd_align_splitter: pure functions of
the C-ABI library, no device), exercised the way fsph_update_iteration uses them: world_size-2 `gloo` ranks all-gather
their key samples and boundary windows, every rank must compute the same splitters, the partition must be balanced,
and a splitter must never cut a work cell (a maximal key-prefix group of <= 128 bodies, tree_topology.h:221-235)."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flecsph_b200 import capi  # noqa: E402

NS, WIN, D, B = 256, 128, 3, 21


def make_keys(n, seed):
    rng = np.random.default_rng(seed)
    # clustered 3-D Morton keys with the sentinel bit: a few dense clumps + background, unique
    pts = np.concatenate([rng.random((n // 2, 3)), 0.5 + 0.02 * rng.standard_normal((n - n // 2, 3))]).clip(0, 1 - 1e-9)
    c = (pts * (1 << B)).astype(np.uint64)
    key = np.full(len(c), np.uint64(1) << np.uint64(63), dtype=np.uint64)
    for b in range(B):
        for d in range(3):
            key |= ((c[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * 3 + d)
    return np.unique(key)


def plan(lib, nranks, local_sorted, gather):
    """What dist_update_iteration does on the host, given an all-gather callable."""
    n = len(local_sorted)
    idx = np.minimum(((np.arange(NS) + 0.5) * n / NS).astype(np.int64), n - 1)
    samples = np.ascontiguousarray(np.concatenate(gather(local_sorted[idx])), dtype=np.uint64)
    counts = np.ascontiguousarray(np.concatenate(gather(np.array([n], np.int64))), dtype=np.int64)
    split = np.zeros(nranks - 1, np.uint64)
    lib.fsph_dd_choose_splitters(nranks, NS, samples.ctypes.data, counts.ctypes.data, split.ctypes.data)
    out = []
    for s in split:
        pos = int(np.searchsorted(local_sorted, s, side="left"))
        below = gather(local_sorted[max(0, pos - WIN):pos])
        above = gather(local_sorted[pos:pos + WIN])
        b = np.ascontiguousarray(np.concatenate(below), dtype=np.uint64)
        a = np.ascontiguousarray(np.concatenate(above), dtype=np.uint64)
        out.append(lib.fsph_dd_align_splitter(D, b.ctypes.data, len(b), a.ctypes.data, len(a), int(s)))
    return np.array(out, dtype=np.uint64), counts


def check_plan(all_keys, split, nranks):
    all_keys = np.sort(all_keys)
    cuts = np.searchsorted(all_keys, split, side="left")
    sizes = np.diff(np.concatenate([[0], cuts, [len(all_keys)]]))
    assert sizes.min() > 0 and sizes.max() <= 1.1 * len(all_keys) / nranks + WIN
    for cpos in cuts:
        lo, hi = all_keys[cpos - 1], all_keys[cpos]
        x = int(lo) ^ int(hi)
        common = B - 1 - (x.bit_length() - 1) // D
        for d in range(common + 1):   # every node that straddles the cut holds more than 128 bodies
            sh = D * (B - d)
            p = int(hi) >> sh
            cnt = int(np.count_nonzero((all_keys >> np.uint64(sh)) == np.uint64(p))) if sh < 64 else len(all_keys)
            assert cnt > 128, "splitter cuts a work cell at depth %d (%d bodies)" % (d, cnt)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = capi.load(capi.PRODUCT_LIB)
    keys = make_keys(60000, 1)
    owner = np.random.default_rng(2).integers(0, world, len(keys))
    local = np.sort(keys[owner == rank])

    def gather(arr):
        objs = [None] * world
        dist.all_gather_object(objs, np.asarray(arr))
        return objs
    split, counts = plan(lib, world, local, gather)
    q.put((rank, split, int(counts.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]), "ranks disagree on the splitters"
    keys = make_keys(60000, 1)
    assert res[0][2] == len(keys)
    check_plan(keys, res[0][1], 2)


@pytest.mark.parametrize("nranks", [2, 3, 5, 8])
def test_plan_many_ranks_in_process(nranks):
    lib = capi.load(capi.PRODUCT_LIB)
    keys = make_keys(100000, 3)
    owner = np.random.default_rng(4).integers(0, nranks, len(keys))
    locals_ = [np.sort(keys[owner == r]) for r in range(nranks)]
    # the same all-gather semantics without processes: every "rank" contributes in lock step
    plans = []
    for r in range(nranks):
        calls = {"i": 0}

        def gather(arr, r=r, calls=calls):
            i = calls["i"]
            calls["i"] += 1
            return [np.asarray(contrib(q, i, split_so_far)) for q in range(nranks)]
        # contributions depend only on the call index and the (identical) splitters computed so far
        split_so_far = {}

        def contrib(q, i, ssf):
            loc = locals_[q]
            if i == 0:
                n = len(loc)
                idx = np.minimum(((np.arange(NS) + 0.5) * n / NS).astype(np.int64), n - 1)
                return loc[idx]
            if i == 1:
                return np.array([len(loc)], np.int64)
            k, which = divmod(i - 2, 2)
            s = ssf["split"][k]
            pos = int(np.searchsorted(loc, s, side="left"))
            return loc[max(0, pos - WIN):pos] if which == 0 else loc[pos:pos + WIN]
        # first pass: raw splitters (needed by contrib for the window calls)
        n0 = [len(x) for x in locals_]
        samples = np.ascontiguousarray(np.concatenate([contrib(q, 0, None) for q in range(nranks)]), dtype=np.uint64)
        counts = np.array(n0, np.int64)
        raw = np.zeros(nranks - 1, np.uint64)
        lib.fsph_dd_choose_splitters(nranks, NS, samples.ctypes.data, counts.ctypes.data, raw.ctypes.data)
        split_so_far["split"] = raw
        split, _ = plan(lib, nranks, locals_[r], gather)
        plans.append(split)
    for r in range(1, nranks):
        assert np.array_equal(plans[0], plans[r])
    check_plan(keys, plans[0], nranks)
