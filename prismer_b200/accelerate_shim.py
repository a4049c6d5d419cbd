"""The subset of ``accelerate.Accelerator`` the reference trainers use (train_caption.py:93,117,132,144-147,158-159,173-176),
backed by one NCCL communicator: data parallelism shards only the global batch; the single collective of a training step
is ONE all-reduce of the flat fp32 gradient buffer over NVLink / NVSwitch, issued right after the
engine's backward; the 1/world average is folded into the fused AdamW update."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import engine


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
        self._fused_optimizer = False

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
                if self.use_distributed:           # replicas start identical (DDP broadcasts rank 0's parameters)
                    dist.broadcast(st.master_t, 0)
                    dist.broadcast(st.master_f, 0)
                    st.refresh(force=True)
                self._models.append(o)
            elif hasattr(o, "grad_scale"):
                o.grad_scale = 1.0 / self.num_processes
                self._fused_optimizer = True
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

    def gather_for_metrics(self, t):
        if not self.use_distributed:
            return t
        out = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(out, t.contiguous())
        return torch.cat(out, 0)

    def save(self, obj, path):
        if self.is_main_process:
            torch.save(obj, path)

    def save_state(self, output_dir):
        if self.is_main_process:
            os.makedirs(output_dir, exist_ok=True)
            for m in self._models:
                torch.save(m.state_dict(), os.path.join(output_dir, "pytorch_model.bin"))


def allreduce_gradients(model, group=None, average: bool = False):
    """The one collective of the step: all-reduce(sum) of the flat fp32 gradient buffer (242 M elements for BASE
    freeze_vision).  The average's 1/world is applied inside the fused AdamW kernel (``FusedAdamW.grad_scale``); when a stock
    torch optimizer is used instead, pass ``average=True`` to divide here."""
    st = engine._store(model)
    dist.all_reduce(st.grad_t, op=dist.ReduceOp.SUM, group=group)
    if average:
        st.grad_t.div_(dist.get_world_size(group))
    return st.grad_t
