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
