"""-m gpu, needs >= 2 GPUs: hash-partitioned DB with one PROCESS per GPU (CUDA IPC handles exchanged through
torch.distributed), each rank placing its own reads against all partitions; rows must equal the oracle's."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp, replicate_table):
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    import parity
    import rappas_b200 as R
    from rappas_b200 import _abi, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = synth.make_db(0, 10, 1999, n_keys=150000, mean_postings=24, seed=7)
        rb = synth.make_reads(db, 4000, (50, 300), seed=100 + rank, n_rate=0.002)
        g = R.Database.from_synth_partitioned_dist(db, device=rank, replicate_table=replicate_table)
        out = g.place(rb)
        oo = O.OracleDB(db).place(rb)
        parity.assert_placements_equal(out, oo, 7, oo["counts"][:, _abi.CNT_AMBIG] > 0)
        dist.barrier()  # nobody frees its partition while a peer still reads it
        g.close()
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("replicate_table", [False, True], ids=["table+postings", "postings-only"])
def test_partitioned_db_one_process_per_gpu(tmp_path, replicate_table):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), replicate_table), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def _xchg_worker(rank, world, port, tmp):
    """One process per GPU, NCCL between them: the exchange form (rp_xchg.cu) on a host-built DB and on
    device-generated partitions of the hash-defined DB."""
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    import parity
    import rappas_b200 as R
    from rappas_b200 import _abi, exchange, synth, synth_hash
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = [exchange.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, 0)
        os.environ["RP_XCHG_PROBES"] = "300000"  # several sub-batches: the pipeline and its buffer hand-over
        if os.environ.get("RP_TEST_XCHG_PUSH") is not None:
            os.environ["RP_XCHG_PUSH"] = os.environ["RP_TEST_XCHG_PUSH"]
        # (a) host-built DB, every rank uploads only its partition
        db = synth.make_db(0, 10, 1999, n_keys=150000, mean_postings=24, seed=7)
        rb = synth.make_reads(db, 3000 + 500 * rank, (50, 400), seed=100 + rank, iupac_rate=0.004, n_rate=0.002)
        part = R.Database.partition_of_synth(db, rank, rank, world)
        x = exchange.Exchange.nccl(part, rank, world, ids[0])
        o = O.OracleDB(db)
        for kw in (dict(), dict(amb_with_max=True)):
            cfg = _abi.place_cfg(**kw)
            out = x.place([rb], cfg)[0]
            oo = o.place(rb, cfg)
            So, _ = o.node_scores(rb, cfg, hitcount=False)
            parity.assert_placements_equal(out, oo, 7, None if kw else oo["counts"][:, _abi.CNT_AMBIG] > 0, So=So)
        assert x.stats()["payload_bytes"] > 0
        dist.barrier()
        x.close()
        part.close()
        # (b) the config-5 construction: partitions generated on the device, big tree, long reads
        ids = [exchange.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, 0)
        hdb = synth_hash.HashDB(k=11, n_nodes=9999, seed=99, occupancy=0.75, mean_postings=48)
        part = R.Database.from_hash_db(hdb, rank, rank, world)
        x = exchange.Exchange.nccl(part, rank, world, ids[0])
        proxy = synth.SynthDB(0, 11, 9999, hdb.thr_lin, hdb.thr_log10, np.zeros(0, np.uint64), np.zeros(1, np.uint64),
                              np.zeros(0, np.uint16), np.zeros(0, np.float32))
        rb = synth.make_reads(proxy, 400, (50, 1500), seed=40 + rank, mode="uniform", iupac_rate=0.005, n_rate=0.002)
        o = O.OracleDB(hdb.sub_db(synth_hash.probed_codes(rb, 11)))
        cfg = _abi.place_cfg()
        out = x.place([rb], cfg)[0]
        oo = o.place(rb, cfg)
        So, _ = o.node_scores(rb, cfg, hitcount=False)
        parity.assert_placements_equal(out, oo, 7, oo["counts"][:, _abi.CNT_AMBIG] > 0, So=So)
        dist.barrier()
        x.close()
        part.close()
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("push", ["1", "0"], ids=["push over peer memory", "nccl all-to-all of the blocks"])
def test_exchange_form_one_process_per_gpu_over_nccl(tmp_path, push, monkeypatch):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    monkeypatch.setenv("RP_TEST_XCHG_PUSH", push)
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_xchg_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
