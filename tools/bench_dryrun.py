"""Host-logic dry run of bench.py on a machine without a GPU: the CUDA library is replaced by the recorder of
tests/test_engine_dryrun_cpu.py (arity-checks every C call, computes nothing) and torch.cuda by inert stand-ins, so the whole
`run_ours` flow -- model build, graph capture, timed loop, e2e prefetch loop, roofline / per-entry-point passes, JSON line -- executes.
Numbers are meaningless; a crash here is a crash on the GPU box.

    python tools/bench_dryrun.py [train|caption] [reference] [nosecondary] [<config name>]
"""
import contextlib, sys, json, io, types
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_engine_dryrun_cpu import _Recorder, _FakeStream, _FakeGraph
from prismer_b200 import _C, ops, engine
import bench

rec = _Recorder()
_C.lib = lambda: rec
ops._stream = lambda: 0
ops._req_cuda = lambda *t: None
engine._experts_check = lambda e: None

class FakeEvent:
    def __init__(self, *a, **k): pass
    def record(self, *a): pass
    def elapsed_time(self, other): return 1.0
    def synchronize(self): pass
torch.cuda.Stream = _FakeStream
torch.cuda.current_stream = lambda *a, **k: _FakeStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.CUDAGraph = _FakeGraph
torch.cuda.graph = lambda g, pool=None, **kw: contextlib.nullcontext()
torch.cuda.Event = FakeEvent
torch.cuda.set_device = lambda d: None
torch.cuda._sleep = lambda n: None
torch.Tensor.pin_memory = lambda self: self
_FakeStream.record_event = lambda self, e=None: None
FakeEvent.record = lambda self, s=None: None
# everything "on the device" stays on the CPU
real_device = torch.device
orig_to = torch.Tensor.to
def to(self, *a, **k):
    a = tuple(real_device('cpu') if isinstance(x, real_device) and x.type == 'cuda' else x for x in a)
    if 'device' in k and isinstance(k['device'], real_device) and k['device'].type == 'cuda': k['device'] = real_device('cpu')
    return orig_to(self, *a, **k)
torch.Tensor.to = to
bench_device = real_device('cpu')
orig_torch_device = torch.device
class DevShim:
    def __call__(self, *a, **k):
        return real_device('cpu')
torch.device = lambda *a, **k: real_device('cpu')
import torch.nn as nn
orig_mod_to = nn.Module.to
nn.Module.to = lambda self, *a, **k: self
orig_full = torch.tensor
torch.tensor = lambda *a, **k: orig_full(*a, **{kk: vv for kk, vv in k.items() if kk != 'device'})
orig_zeros = torch.zeros
mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl='ours', mode=mode, batch=2, no_cpu_baseline=True, eager=False,
                             reference_inputs='reference' in sys.argv, no_secondary='nosecondary' in sys.argv,
                             config=next((a for a in sys.argv[2:] if a in bench.CONFIGS), 'base_caption224'))
bench.run_ours(args)
print('calls:', sum(rec.calls.values()))
