"""Execute `-m gpu` test functions on a machine WITHOUT a GPU: the CUDA library is replaced by the arity-checking recorder of
tests/test_engine_dryrun_cpu.py, `.cuda()` / device="cuda" are mapped to the CPU.  Values are garbage, so AssertionError is expected;
any OTHER exception is a bug in the test or in the host code it drives -- found here instead of on the GPU box.

    python tools/gpu_tests_dryrun.py tests.test_zzz_beam_gpu tests.test_zzzz_compact_gpu ...
(tests that take a fixture, stack several parametrize marks or build CUDA graphs are skipped / reported by this simple driver)"""
import contextlib, importlib, inspect, sys, traceback, types, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pytest
from tests.test_engine_dryrun_cpu import _Recorder, _FakeStream, _FakeGraph, _FakeEvent
from prismer_b200 import _C, ops, engine
rec = _Recorder()
_C.lib = lambda: rec
ops._stream = lambda: 0
ops._req_cuda = lambda *t: None
engine._experts_check = lambda e: None
engine.SIDE_STREAM = False
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.current_stream = lambda *a, **k: _FakeStream()
torch.cuda.Event = _FakeEvent
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.current_device = lambda: 0
real_device = torch.device
orig_to = torch.Tensor.to
def to(self, *a, **k):
    a = tuple('cpu' if (isinstance(x, str) and x.startswith('cuda')) or (isinstance(x, real_device) and x.type == 'cuda') else x for x in a)
    if 'device' in k and str(k['device']).startswith('cuda'): k['device'] = 'cpu'
    return orig_to(self, *a, **k)
torch.Tensor.to = to
torch.Tensor.cuda = lambda self, *a, **k: self
import torch.nn as nn
nn.Module.cuda = lambda self, *a, **k: self
orig_module_to = nn.Module.to
nn.Module.to = lambda self, *a, **k: self
for fname in ['randn', 'zeros', 'empty', 'ones', 'full', 'tensor', 'arange', 'randperm', 'zeros_like', 'empty_like']:
    orig = getattr(torch, fname)
    def mk(orig):
        def f(*a, **k):
            if 'device' in k and str(k['device']).startswith('cuda'): k['device'] = 'cpu'
            if 'generator' in k and k['generator'] is not None and getattr(k['generator'], 'device', None) is not None and str(k['generator'].device).startswith('cuda'):
                k['generator'] = None
            return orig(*a, **k)
        return f
    setattr(torch, fname, mk(orig))
origG = torch.Generator
class G:
    def __new__(cls, device='cpu'):
        return origG('cpu')
torch.Generator = G
orig_prepare = engine.prepare
engine.prepare = lambda root, device=None: orig_prepare(root, real_device('cpu'))
import prismer_b200.modeling
bad = 0
for modname in (sys.argv[1:] or ['tests.test_zzz_beam_gpu', 'tests.test_zzz_surface_golden_gpu', 'tests.test_zzzz_compact_gpu']):
    mod = importlib.import_module(modname)
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith('test_'): continue
        params = [()]
        for mark in getattr(fn, 'pytestmark', []):
            if mark.name == 'parametrize':
                vals = mark.args[1]
                params = [v if isinstance(v, tuple) else (v,) for v in vals][:2]
        sig = inspect.signature(fn)
        for p in params:
            kwargs = {}
            try:
                if 'base' in sig.parameters:
                    base_fn = mod.base.__wrapped__ if hasattr(mod.base, '__wrapped__') else None
                    if base_fn is None:
                        continue
                    # shrink the BASE fixture: tiny model instead (host code identical)
                    kwargs['base'] = None
                    continue
                fn(*p)
                print('ran  ', modname.split('.')[-1], name, p if p else '')
            except AssertionError:
                print('assert', modname.split('.')[-1], name, '(expected: values are garbage)')
            except Exception as e:
                bad += 1
                print('ERROR', modname.split('.')[-1], name, type(e).__name__, str(e)[:200])
                traceback.print_exc(limit=4)
print('non-assertion errors:', bad)

