"""The subset of ``accelerate.Accelerator`` the reference trainers use (train_caption.py:93,117,132,144-147,158-159,173-176),
backed by one NCCL communicator: data parallelism shards only the global batch; the single collective of a training step
is ONE all-reduce of the flat fp32 gradient buffer over NVLink / NVSwitch, issued right after the
engine's backward; the 1/world average is folded into the fused AdamW update."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import engine


# ------------------------------------------------------------------------------------------------ data-loader side of prepare()
def _to_device(obj, device, non_blocking=True):
    """What accelerate's prepared loaders do with ``device_placement=True``: tensors (and anything with a tensor-like ``.to``, e.g.
    ``data.CompactMap``) inside nested dict / list / tuple batches go to the device; strings and numbers are left alone."""
    if torch.is_tensor(obj) or (hasattr(obj, "to") and hasattr(obj, "u8")):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, dict):
        return type(obj)((k, _to_device(v, device, non_blocking)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)) and not isinstance(obj, str):
        return type(obj)(_to_device(v, device, non_blocking) for v in obj)
    return obj


class BatchShard:
    """accelerate's ``BatchSamplerShard`` (split_batches=False, even_batches=True): of every ``world`` consecutive batches of the
    base batch sampler, process ``rank`` takes the ``rank``-th.  A short last batch / an incomplete last round is completed with
    indices from the start of the epoch so that every process runs the same number of equally sized steps (the duplicated
    samples are dropped again by ``Accelerator.gather_for_metrics``)."""

    def __init__(self, batch_sampler, world: int, rank: int):
        self.batch_sampler, self.world, self.rank = batch_sampler, world, rank
        self.batch_size = getattr(batch_sampler, "batch_size", None)
        self.drop_last = getattr(batch_sampler, "drop_last", False)

    def __len__(self):
        n = len(self.batch_sampler)
        return n // self.world if self.drop_last else -(-n // self.world)

    def __iter__(self):
        first, group = [], []
        for batch in self.batch_sampler:
            batch = list(batch)
            if len(first) < (self.batch_size or len(batch)) * self.world:
                first += batch                                   # indices to recycle when the epoch does not divide evenly
            group.append(batch)
            if len(group) == self.world and len(batch) == (self.batch_size or len(batch)):
                yield group[self.rank]
                group = []
        if group and not self.drop_last:
            bs = self.batch_size or len(group[0])
            flat = [i for b in group for i in b]
            need = bs * self.world - len(flat)
            pool = first
            while need > 0 and pool:
                take = pool[:need]
                flat += take
                need -= len(take)
            yield flat[self.rank * bs:(self.rank + 1) * bs]


class ShardedLoader:
    """Prepared data loader: this process' shard of the batches, already on the device.  Keeps ``.dataset`` / ``len()`` (the
    reference reads ``test_loader.dataset.data_list`` and ``len(train_loader)``, train_caption.py:127,150)."""

    def __init__(self, loader, accelerator):
        from torch.utils.data import DataLoader
        self.base, self.acc = loader, accelerator
        self.dataset = loader.dataset
        self.batch_size = loader.batch_size
        world, rank = accelerator.num_processes, accelerator.process_index
        if world > 1:
            shard = BatchShard(loader.batch_sampler, world, rank)
            kw = dict(num_workers=loader.num_workers, collate_fn=loader.collate_fn, pin_memory=loader.pin_memory,
                      timeout=loader.timeout, worker_init_fn=loader.worker_init_fn)
            if loader.num_workers > 0:
                kw.update(prefetch_factor=loader.prefetch_factor, persistent_workers=loader.persistent_workers)
            self.loader = DataLoader(loader.dataset, batch_sampler=shard, **kw)
        else:
            self.loader = loader
        # shuffling must draw the SAME permutation on every rank (accelerate synchronises the sampler's generator): one seed from
        # rank 0, advanced per epoch
        self.epoch = 0
        seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)
        if world > 1:
            seed = seed.to(accelerator.device) if dist.get_backend() == "nccl" else seed
            dist.broadcast(seed, 0)
        self.seed = int(seed)
        total, per_round = len(loader.dataset), (loader.batch_size or 1) * world
        self.remainder = total % per_round if not getattr(loader, "drop_last", False) else 0

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        self.acc._end_of_loader, self.acc._remainder = False, 0
        sampler = getattr(self.base, "sampler", None)
        if isinstance(sampler, torch.utils.data.RandomSampler) and self.acc.num_processes > 1:
            sampler.generator = torch.Generator().manual_seed(self.seed + self.epoch)
        self.epoch += 1
        it = iter(self.loader)
        try:
            cur = next(it)
        except StopIteration:
            return
        while True:
            try:
                nxt = next(it)
            except StopIteration:
                self.acc._end_of_loader, self.acc._remainder = True, self.remainder
                yield _to_device(cur, self.acc.device)
                return
            yield _to_device(cur, self.acc.device)
            cur = nxt


class Accelerator:
    def __init__(self, mixed_precision: str = "bf16", **_):
        self.mixed_precision = mixed_precision
        self.use_distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
        if self.use_distributed and not dist.is_initialized():
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            dist.init_process_group(backend)
        self.process_index = dist.get_rank() if self.use_distributed else 0
        self.num_processes = dist.get_world_size() if self.use_distributed else 1
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_process_index)
            self.device = torch.device("cuda", self.local_process_index)
        else:
            self.device = torch.device("cpu")
        self._models = []
        self._optimizers = []
        self._fused_optimizer = False
        self._end_of_loader, self._remainder = False, 0

    @property
    def is_main_process(self):
        return self.process_index == 0

    def print(self, *a, **k):
        if self.is_main_process:
            print(*a, **k)

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o.to(self.device)
                st = engine.prepare(o, self.device)
                if self.use_distributed:           # replicas start identical (DDP broadcasts rank 0's parameters and buffers)
                    dist.broadcast(st.master_t, 0)
                    dist.broadcast(st.master_f, 0)
                    for b in o.buffers():          # BatchNorm running statistics / num_batches_tracked
                        dist.broadcast(b.data, 0)
                    st.refresh(force=True)
                self._models.append(o)
            elif hasattr(o, "grad_scale"):
                o.grad_scale = 1.0 / self.num_processes
                self._fused_optimizer = True
                self._optimizers.append(o)
            elif isinstance(o, torch.optim.Optimizer):
                self._optimizers.append(o)
            elif isinstance(o, torch.utils.data.DataLoader):
                o = ShardedLoader(o, self)         # this rank's batches, moved to the device (train_caption.py:115-117,126)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def backward(self, loss):
        loss.backward()
        if self.use_distributed:
            for m in self._models:
                # with the fused optimizer the 1/world average is folded into its kernel; a stock torch optimizer needs it here
                allreduce_gradients(m, average=not self._fused_optimizer)

    def wait_for_everyone(self):
        if self.use_distributed:
            dist.barrier()

    def gather(self, obj):
        """Concatenate every process' tensors along dim 0 in rank order (nested tuple / list / dict of tensors)."""
        if not self.use_distributed:
            return obj
        if torch.is_tensor(obj):
            t = obj.contiguous()
            if t.dim() == 0:
                t = t[None]
            out = [torch.empty_like(t) for _ in range(self.num_processes)]
            dist.all_gather(out, t)
            return torch.cat(out, 0)
        if isinstance(obj, dict):
            return type(obj)((k, self.gather(v)) for k, v in obj.items())
        if isinstance(obj, (list, tuple)) and all(torch.is_tensor(v) or isinstance(v, (list, tuple, dict)) for v in obj):
            return type(obj)(self.gather(v) for v in obj)
        raise TypeError(f"gather: unsupported object of type {type(obj).__name__} (tensors in tuples / lists / dicts only)")

    def gather_for_metrics(self, obj):
        """``gather`` + drop the samples that were duplicated to even out the last round of a prepared loader
        (train_caption.py:147,190: ``data_ids, captions = accelerator.gather_for_metrics((data_ids, captions))``)."""
        out = self.gather(obj)
        if self.use_distributed and self._end_of_loader and self._remainder > 0:
            n = self._remainder
            cut = lambda o: (o[:n] if torch.is_tensor(o) else (type(o)((k, cut(v)) for k, v in o.items()) if isinstance(o, dict)
                                                                else type(o)(cut(v) for v in o)))
            out = cut(out)
        return out

    def save(self, obj, path):
        if self.is_main_process:
            torch.save(obj, path)

    @staticmethod
    def _suffix(i):
        return "" if i == 0 else f"_{i}"

    def save_state(self, output_dir):
        """``accelerator.save_state`` (train_caption.py:173-176) with accelerate's file layout: ``pytorch_model[_i].bin`` (reference-layout
        state_dict), ``optimizer[_i].bin`` (moments, step count, param_groups) on the main process and ``random_states_<rank>.pkl``
        (python / numpy / torch generators) on every process -- everything ``load_state`` needs to resume a run."""
        import pickle
        import random
        import numpy as np
        os.makedirs(output_dir, exist_ok=True)
        if self.is_main_process:
            for i, m in enumerate(self._models):
                torch.save(m.state_dict(), os.path.join(output_dir, f"pytorch_model{self._suffix(i)}.bin"))
            for i, o in enumerate(self._optimizers):
                torch.save(o.state_dict(), os.path.join(output_dir, f"optimizer{self._suffix(i)}.bin"))
        rng = {"python": random.getstate(), "numpy": np.random.get_state(), "torch": torch.get_rng_state()}
        if torch.cuda.is_available():
            rng["cuda"] = torch.cuda.get_rng_state_all()
        for i, m in enumerate(self._models):       # device-side Philox key of the dropout masks (engine.ParamStore.seed)
            rng[f"dropout_seed{self._suffix(i)}"] = int(engine._store(m).seed.item()) if hasattr(m, "_prismer_store") else 0
        with open(os.path.join(output_dir, f"random_states_{self.process_index}.pkl"), "wb") as f:
            pickle.dump(rng, f)
        self.wait_for_everyone()

    def load_state(self, input_dir):
        """Inverse of ``save_state``: weights into the flat fp32 masters (bf16 compute copies re-derived), optimizer state, RNG streams."""
        import pickle
        import random
        import numpy as np
        for i, m in enumerate(self._models):
            sd = torch.load(os.path.join(input_dir, f"pytorch_model{self._suffix(i)}.bin"), map_location="cpu")
            m.load_state_dict(sd)
            engine._store(m).refresh(force=True)
        for i, o in enumerate(self._optimizers):
            pth = os.path.join(input_dir, f"optimizer{self._suffix(i)}.bin")
            if os.path.exists(pth):
                o.load_state_dict(torch.load(pth, map_location=self.device))
        pth = os.path.join(input_dir, f"random_states_{self.process_index}.pkl")
        if os.path.exists(pth):
            with open(pth, "rb") as f:
                rng = pickle.load(f)
            random.setstate(rng["python"]); np.random.set_state(rng["numpy"]); torch.set_rng_state(rng["torch"])
            if "cuda" in rng and torch.cuda.is_available():
                torch.cuda.set_rng_state_all(rng["cuda"])
            for i, m in enumerate(self._models):
                if hasattr(m, "_prismer_store"):
                    engine._store(m).seed.fill_(rng.get(f"dropout_seed{self._suffix(i)}", 0))
        self.wait_for_everyone()


def allreduce_gradients(model, group=None, average: bool = False):
    """The one collective of the step: all-reduce(sum) of the flat fp32 gradient buffer (242 M elements for BASE
    freeze_vision).  The average's 1/world is applied inside the fused AdamW kernel (``FusedAdamW.grad_scale``); when a stock
    torch optimizer is used instead, pass ``average=True`` to divide here."""
    st = engine._store(model)
    dist.all_reduce(st.grad_t, op=dist.ReduceOp.SUM, group=group)
    if average:
        st.grad_t.div_(dist.get_world_size(group))
    return st.grad_t
