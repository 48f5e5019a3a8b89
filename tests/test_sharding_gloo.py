"""N>1 host logic on CPU: world_size-2 gloo run of the event sharding used by bench.py
--gpus N (events are independent, no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from traccc_b200.sharding import events_of_rank, stream_of_event


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_events, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from traccc_b200 import toy_detector
    mine = events_of_rank(n_events, rank, world)
    seeds = 0
    for i in mine:                                   # each rank seeds only its own events
        ev = toy_detector.generate_event(60, 500 + i)
        seeds += len(oracle.run(ev.xyz, dump=False).seeds["bottom"])
    t = torch.tensor([len(mine), seeds, 1000 + rank], dtype=torch.int64)
    tot = t.clone()
    dist.all_reduce(tot[:2], op=dist.ReduceOp.SUM)   # bookkeeping only
    mx = t[2:].clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)        # "max over ranks" timing reduction
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([tot[0], tot[1], mx[0]]))
    dist.barrier()
    dist.destroy_process_group()


def test_event_sharding_world2(tmp_path):
    n_events, world = 7, 2
    assert events_of_rank(7, 0, 2) == [0, 2, 4, 6] and events_of_rank(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((events_of_rank(n_events, r, 3) for r in range(3)), [])) == list(range(n_events))
    assert [stream_of_event(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    mp.spawn(_worker, args=(world, _free_port(), n_events, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle
    from traccc_b200 import toy_detector
    expect = sum(len(oracle.run(toy_detector.generate_event(60, 500 + i).xyz, dump=False).seeds["bottom"])
                 for i in range(n_events))
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got[0] == n_events and got[1] == expect and got[2] == 1000 + world - 1
