"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes agree on the row partition and on
the bootstrap of the communicator id (no GPU compute)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n in (0, 1, 63, 64, 1000, 20301, 10039316):
        cuts = pkg.shard_bounds(n, world)
        mine = [cuts[rank], cuts[rank + 1]]
        allc = [None] * world
        dist.all_gather_object(allc, mine)
        ok &= allc[0][0] == 0 and allc[-1][1] == n
        ok &= all(allc[r][1] == allc[r + 1][0] for r in range(world - 1))
        ok &= all(c % 64 == 0 for c in cuts[1:-1]) or n < 64 * world
    box = [b"x" * 128 if rank == 0 else None]       # the id bootstrap of Comm.from_torch
    dist.broadcast_object_list(box, src=0)
    ok &= box[0] == b"x" * 128
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_partition_and_bootstrap_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, 29533, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    for P in (1, 2, 3, 4, 8):
        for n in (0, 5, 64, 65, 1000405, 10039316):
            cuts = pkg.shard_bounds(n, P)
            assert cuts[0] == 0 and cuts[-1] == n and all(np.diff(cuts) >= 0)
