"""CPU, world_size 2, gloo: the data-parallel gradient exchange of the training step (one all-reduce of the flat gradient
buffer, 1/world folded into the optimizer) equals single-process training on the concatenated batch for a linear model --
the host-side logic of accelerate_shim.allreduce_gradients / FusedAdamW.grad_scale without any GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from prismer_b200 import accelerate_shim
    torch.manual_seed(0)
    w = torch.randn(16, 8)
    x = torch.randn(4 * world, 8)[rank * 4:(rank + 1) * 4]          # this rank's shard of the global batch
    y = x @ w.t()
    grad_local = (2 * y).t() @ x / x.shape[0]                        # d/dw mean_b |y|^2 over the local shard
    store = SimpleNamespace(grad_t=grad_local.reshape(-1).clone())
    model = torch.nn.Linear(8, 16)
    model._prismer_store = store                                     # what engine._store(model) returns
    accelerate_shim.allreduce_gradients(model)                       # ONE all-reduce(sum) of the flat buffer
    out[rank] = store.grad_t.view(16, 8) * (1.0 / world)             # FusedAdamW.grad_scale = 1/world
    dist.destroy_process_group()


def test_two_rank_gradient_equals_global_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    w = torch.randn(16, 8)
    x = torch.randn(4 * world, 8)
    y = x @ w.t()
    ref = (2 * y).t() @ x / x.shape[0]
    for r in range(world):
        assert torch.allclose(out[r], ref, atol=1e-5), r
