"""CPU, world_size 2, gloo: the data-loader side of ``Accelerator.prepare`` (accelerate's DataLoaderShard / BatchSamplerShard
semantics the reference trainers rely on, train_caption.py:115-117,126,140-147): every rank sees a disjoint shard in rank-
interleaved batch order, the same number of equally sized steps, and ``gather_for_metrics`` returns each sample exactly once in
dataset order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data import DataLoader, Dataset


class _Toy(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {"rgb": torch.full((3, 2, 2), float(i)), "obj_detection": {"label": torch.full((1,), float(i)), "instance": torch.tensor([i])}}, i


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, bs, out, shuffle=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from prismer_b200.accelerate_shim import Accelerator
    acc = Accelerator()
    torch.manual_seed(100 + rank)                                   # ranks have DIFFERENT global RNG states
    loader = acc.prepare(DataLoader(_Toy(n), batch_size=bs, shuffle=shuffle))
    assert loader.dataset.n == n
    seen, gathered, steps = [], [], 0
    for experts, ids in loader:
        steps += 1
        assert experts["rgb"].shape[0] == bs and torch.equal(experts["obj_detection"]["instance"][:, 0], ids)
        seen += ids.tolist()
        g_ids, g_rgb = acc.gather_for_metrics((ids, experts["rgb"][:, 0, 0, 0]))
        assert torch.equal(g_ids.float(), g_rgb)
        gathered += g_ids.tolist()
    assert steps == len(loader)
    out[rank] = (seen, gathered, steps)
    acc.wait_for_everyone()
    dist.destroy_process_group()


def _run(n, bs, world=2, shuffle=False):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n, bs, out, shuffle), nprocs=world, join=True)
    return out


def test_shuffled_epoch_uses_one_permutation_on_all_ranks():
    out = _run(n=16, bs=4, shuffle=True)
    assert sorted(out[0][0] + out[1][0]) == list(range(16))          # disjoint shards covering the epoch
    assert out[0][1] == out[1][1] and sorted(out[0][1]) == list(range(16)) and out[0][1] != list(range(16))


def test_even_epoch_is_partitioned_in_interleaved_batch_order():
    out = _run(n=16, bs=4)
    assert out[0][0] == [0, 1, 2, 3, 8, 9, 10, 11] and out[1][0] == [4, 5, 6, 7, 12, 13, 14, 15]
    assert out[0][1] == list(range(16)) == out[1][1] and out[0][2] == out[1][2] == 2


def test_uneven_epoch_is_completed_and_deduplicated():
    out = _run(n=13, bs=4)                     # 4 batches (last one short) -> 2 rounds of 2 ranks
    assert out[0][2] == out[1][2] == 2         # same number of steps on every rank, every batch full
    assert len(out[0][0]) == len(out[1][0]) == 8
    assert out[0][1] == list(range(13)) == out[1][1]      # each sample exactly once, dataset order
    out = _run(n=9, bs=4)                      # 3 batches: the last round has one (short) batch for two ranks
    assert out[0][2] == out[1][2] == 2 and out[0][1] == list(range(9)) == out[1][1]


def test_single_process_prepare_moves_nothing_and_gathers_identity():
    from prismer_b200.accelerate_shim import Accelerator, _to_device
    os.environ.pop("WORLD_SIZE", None)
    acc = Accelerator()
    loader = acc.prepare(DataLoader(_Toy(5), batch_size=2))
    assert len(loader) == 3 and sum(len(i) for _, i in loader) == 5
    t = torch.arange(3)
    assert acc.gather_for_metrics((t, t))[0] is t
    moved = _to_device({"a": [t, "text"], "b": (t, 3)}, torch.device("cpu"))
    assert moved["a"][1] == "text" and moved["b"][1] == 3 and torch.equal(moved["a"][0], t)
