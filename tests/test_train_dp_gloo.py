"""CPU, world size 2 over gloo: the host-side logic of data-parallel training — gradient bucket all-reduce (mean over ranks), the
all-reduce(max) of the close-example table, rank-strided epoch lists."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl4vc_b200.train_dp import GradientAllReducer, shard_epoch_indices, sync_close_table


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv2hidden = torch.nn.Linear(6, 4)
        self.conv1D = torch.nn.Linear(3, 2)
        self.fcHidden2VT = torch.nn.Linear(4, 3)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _Tiny()
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    GradientAllReducer(model)()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(model.parameters()))
    table = torch.zeros(10, dtype=torch.uint8)
    table[rank * 3] = 1
    sync_close_table(table)
    ok = ok and table.tolist() == [1, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    mine = shard_epoch_indices(list(range(11)), rank, world)
    ok = ok and mine == list(range(10))[rank::2]
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_and_close_table_sync_world2():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out[0] and out[1]
