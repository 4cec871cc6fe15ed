"""world_size-2 checks of the N>1 plumbing on CPU (gloo, 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swarmkit_b200.dist import reduce_step


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, c = reduce_step([10.0 + rank, 5.0 - rank], [100.0 * (rank + 1)])
    out.put((rank, t, c))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_step_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, t, c in res:
        assert t == [11.0, 5.0] and c == [300.0]      # MAX of times, SUM of work


def _id_worker(rank, world, port, out):
    """The rendezvous bench.py / the shim performs for a node-sharded engine group: rank 0's ncclUniqueId reaches
    every rank; without a GPU the engine then refuses to start (no CPU fallback, also for world_size > 1)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swarmkit_b200 import PlacementEngine
    from swarmkit_b200.abi import EngineError, PE_ERR_NO_DEVICE
    from swarmkit_b200.engine import nccl_unique_id
    try:
        box = [nccl_unique_id() if rank == 0 else None]
    except RuntimeError as e:       # no libnccl on this machine: nothing to rendezvous with
        box = [None]
        out.put((rank, "no-nccl", str(e)))
        dist.barrier()
        dist.destroy_process_group()
        return
    dist.broadcast_object_list(box, src=0)
    verdict = "unexpected"
    if not torch.cuda.is_available():
        try:
            PlacementEngine(node_capacity=16, rank=rank, world_size=world, nccl_id=box[0])
        except EngineError as e:
            verdict = "refused" if e.code == PE_ERR_NO_DEVICE else f"code {e.code}"
    else:
        verdict = "gpu-present"
    out.put((rank, verdict, bytes(box[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_rendezvous_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    if res[0][1] == "no-nccl":
        pytest.skip("libnccl.so.2 not available")
    assert res[0][2] == res[1][2] and len(res[0][2]) == 128      # both ranks hold rank 0's id
    assert all(v in ("refused", "gpu-present") for _, v, _ in res)
