"""world_size-2 gloo test (CPU) of the one-process-per-GPU read sharding: each rank places its slice with
its own DB replica, rank 0 gathers, and the merged rows equal a single-process run of the whole batch.
The per-rank engine here is the CPU oracle (test infrastructure); on a GPU box bench.py runs the same
split with rappas_b200.Database per rank."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from rappas_b200 import _abi, shard, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = synth.make_db(0, 8, 299, n_keys=20000, mean_postings=12, seed=43)       # replica on every rank
        rb = synth.make_reads(db, 1001, (20, 220), seed=1043, n_rate=0.003)          # odd count: ragged split
        cfg = _abi.place_cfg()
        odb = O.OracleDB(db)
        local, (lo, hi) = shard.place_local_shard(lambda r, c: odb.place(r, c), rb, cfg, rank, world)
        assert local["n_rows"].shape[0] == hi - lo
        merged = shard.gather_results(local, dst=0)
        dist.barrier()
        if rank == 0:
            whole = odb.place(rb, cfg)
            for k in shard.RESULT_KEYS:
                assert np.array_equal(merged[k], whole[k], equal_nan=True), k
            open(os.path.join(tmp, "ok"), "w").write("%d" % merged["n_rows"].shape[0])
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_and_match_c_split():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            b = shard.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_placement_equals_single_process(tmp_path):
    O.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1001"


def _owner_restatement(alphabet, k, keys, n_parts):
    """numpy restatement of rp_common.h planar_from_code -> hash_key -> owner_of (uint32 arithmetic)."""
    bits = 2 if alphabet == 0 else 5
    keys = keys.astype(np.uint64)
    planar = np.zeros_like(keys)
    for i in range(k):
        st = (keys >> np.uint64(bits * i)) & np.uint64((1 << bits) - 1)
        for p in range(bits):
            planar |= ((st >> np.uint64(p)) & np.uint64(1)) << np.uint64(p * k + i)
    M = np.uint64(0xFFFFFFFF)
    u = np.uint64
    lo, hi = planar & M, planar >> u(32)
    a = lo ^ ((hi * u(0x9E3779B1)) & M)                                   # hash_key().a : lowbias32
    a ^= a >> u(16); a = (a * u(0x7FEB352D)) & M
    a ^= a >> u(15); a = (a * u(0x846CA68B)) & M
    a ^= a >> u(16)
    b = ((hi ^ ((lo * u(0x85EBCA6B)) & M)) + u(0x7F4A7C15)) & M           # hash_key().b : triple32
    b ^= b >> u(17); b = (b * u(0xED5AD4BB)) & M
    b ^= b >> u(11); b = (b * u(0xAC4C1B51)) & M
    b ^= b >> u(15); b = (b * u(0x31848BAB)) & M
    b ^= b >> u(14)
    m = ((a ^ (((b << u(16)) | (b >> u(16))) & M)) * u(0x85EBCA6B)) & M   # owner_of
    return (((m >> u(16)) * u(n_parts)) >> u(16)).astype(np.int32)


@pytest.mark.parametrize("alphabet,k", [(0, 10), (0, 15), (0, 20), (1, 6)])
def test_partition_of_keys_matches_restatement_and_balances(alphabet, k):
    import rappas_b200 as R
    rng = np.random.default_rng(k)
    bits = 2 if alphabet == 0 else 5
    if alphabet == 0:
        keys = rng.integers(0, 1 << (bits * k), size=50000, dtype=np.uint64)
    else:  # 5-bit states 0..19
        st = rng.integers(0, 20, size=(50000, k), dtype=np.uint64)
        keys = (st << (np.arange(k, dtype=np.uint64) * np.uint64(5))).sum(axis=1).astype(np.uint64)
    for w in (1, 2, 3, 8):
        own = R.partition_of_keys(alphabet, k, keys, w)
        assert np.array_equal(own, _owner_restatement(alphabet, k, keys, w))
        cnt = np.bincount(own, minlength=w)
        assert cnt.min() > 0.9 * len(keys) / w and cnt.max() < 1.1 * len(keys) / w


# ------------------------------------------------------------------------------------------------------------
# Exchange form (rp_xchg.cu): the host-side plan -- who sends how many keys / posting-block units to whom, and
# where they land -- computed by every rank from the all-gathered count matrices.  World-size-2 and -3 gloo
# groups on the CPU: every rank derives ITS plan (rp_xchg_plan, pure host code of the product library) and the
# ranks' plans must agree pairwise: what l sends p is what p expects from l, and the receive ranges tile.
def _plan_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from rappas_b200 import _abi
    from rappas_b200._lib import check, load
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        J = 5
        rng = np.random.default_rng(100 + rank)
        mine = rng.integers(0, 1000, size=(J, world)).astype(np.uint64)      # probes of my sub-batch j for owner o
        mine[rng.integers(0, J)] = 0                                         # an empty sub-batch
        g = [torch.empty(J * world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(g, torch.from_numpy(mine.reshape(-1).astype(np.int64)))
        probes = np.stack([t.numpy().astype(np.uint64).reshape(J, world) for t in g])         # [w][j][o]
        # units I (as owner) send to home p for sub-batch j: any function of the probes I receive
        units_mine = (probes[:, :, rank] * np.uint64(3) // np.uint64(2)).astype(np.uint64)    # [p][j]
        g = [torch.empty(world * J, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(g, torch.from_numpy(units_mine.reshape(-1).astype(np.int64)))
        units = np.stack([t.numpy().astype(np.uint64).reshape(world, J) for t in g])          # [o][p][j]
        fn = load()
        plans = {}
        for direct in (0, 1):
            o = {k: np.zeros(n, np.uint64) for k, n in (("kso", world + 1), ("kro", world + 1), ("seg", world * J + 1),
                                                        ("home", world * J), ("pso", J * world), ("psc", J * world),
                                                        ("pro", J * world), ("prc", J * world), ("caps", 3))}
            check(fn["xchg_plan"](world, J, rank, direct, _abi.ptr(np.ascontiguousarray(probes)), _abi.ptr(np.ascontiguousarray(units)),
                                  *[_abi.ptr(o[k]) for k in ("kso", "kro", "seg", "home", "pso", "psc", "pro", "prc", "caps")]))
            plans[direct] = o
            # my own view is self-consistent
            assert o["kso"][-1] == probes[rank].sum() and o["kro"][-1] == probes[:, :, rank].sum()
            assert np.array_equal(np.diff(o["seg"]), probes[:, :, rank].reshape(-1))          # (source p, sub-batch j) order
            for j in range(J):
                rc, ro = o["prc"][j * world:(j + 1) * world], o["pro"][j * world:(j + 1) * world]
                assert np.array_equal(ro, np.concatenate([[0], np.cumsum(rc)[:-1]]))          # receive ranges tile
                assert rc.sum() <= o["caps"][1]
                if direct:
                    assert rc[rank] == 0 and o["psc"][j * world + rank] == 0
        # pairwise agreement: gather everybody's send / receive counts
        for direct in (0, 1):
            o = plans[direct]
            g = [None] * world
            dist.all_gather_object(g, {"kso": o["kso"], "kro": o["kro"], "psc": o["psc"], "prc": o["prc"]})
            for l in range(world):
                for p in range(world):
                    assert g[l]["kso"][p + 1] - g[l]["kso"][p] == g[p]["kro"][l + 1] - g[p]["kro"][l]
                    for j in range(J):
                        assert g[l]["psc"][j * world + p] == g[p]["prc"][j * world + l]
        open(os.path.join(tmp, "plan%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_exchange_plan_agrees_across_ranks(tmp_path, world):
    mp.spawn(_plan_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("plan%d" % r)).exists() for r in range(world))


@pytest.mark.parametrize("world,J", [(8, 7), (5, 1), (1, 3)])
def test_exchange_plan_of_every_rank_in_one_process(world, J):
    """The plan of all `world` ranks computed side by side (no process group needed: the plan is a pure function of the
    gathered count matrices) at the sizes the 8-GPU runs use.  What the PUSH transport relies on: owner l writes its
    blocks of sub-batch j for home p at p's receive offset [j][l], so for every (p, j) those ranges must tile p's
    buffer without overlap and stay under its capacity; and what l sends p is what p expects from l."""
    from rappas_b200 import _abi
    from rappas_b200._lib import check, load
    fn = load()
    rng = np.random.default_rng(7 * world + J)
    probes = rng.integers(0, 5000, size=(world, J, world)).astype(np.uint64)   # [home w][sub-batch j][owner o]
    probes[rng.integers(0, world), rng.integers(0, J)] = 0
    units = np.zeros((world, world, J), np.uint64)                               # [owner o][home p][j]
    for o in range(world):
        units[o] = (probes[:, :, o] * np.uint64(5) // np.uint64(3)).astype(np.uint64)
    for direct in (0, 1):
        plans = []
        for rank in range(world):
            out = {k: np.zeros(n, np.uint64) for k, n in (("kso", world + 1), ("kro", world + 1), ("seg", world * J + 1),
                                                          ("home", world * J), ("pso", J * world), ("psc", J * world),
                                                          ("pro", J * world), ("prc", J * world), ("caps", 3))}
            check(fn["xchg_plan"](world, J, rank, direct, _abi.ptr(np.ascontiguousarray(probes)), _abi.ptr(np.ascontiguousarray(units)),
                                  *[_abi.ptr(out[k]) for k in ("kso", "kro", "seg", "home", "pso", "psc", "pro", "prc", "caps")]))
            plans.append(out)
        for l in range(world):
            for p in range(world):
                assert plans[l]["kso"][p + 1] - plans[l]["kso"][p] == plans[p]["kro"][l + 1] - plans[p]["kro"][l] == probes[l, :, p].sum()
                for j in range(J):
                    expect = 0 if (direct and l == p) else units[l, p, j]
                    assert plans[l]["psc"][j * world + p] == plans[p]["prc"][j * world + l] == expect
        for p in range(world):
            for j in range(J):
                ro, rc = plans[p]["pro"][j * world:(j + 1) * world], plans[p]["prc"][j * world:(j + 1) * world]
                order = np.argsort(ro, kind="stable")
                ends = (ro + rc)[order]
                assert np.all(ro[order][1:] >= ends[:-1])          # the owners' ranges do not overlap
                assert ends.max(initial=0) <= plans[p]["caps"][1]  # and fit the receive buffer
