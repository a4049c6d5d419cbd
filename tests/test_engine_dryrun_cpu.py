"""CPU dry run of the ENGINE's host code: every C-ABI call is replaced by a recorder that checks the call's arity / scalar types against
the declared signature and returns success without computing anything (outputs stay uninitialised).  Forward + backward of the tiny
model, compact inputs, eval-mode BatchNorm, the experimental attention paths, greedy / beam generation and rank inference all run
through the real ``engine`` / ``ops`` / ``generation`` code, so NameError / AttributeError / wrong-argument-count / shape-plumbing bugs
in GPU-only branches surface here instead of on the GPU box.  (Values are meaningless by construction -- this is not a parity test.)"""
import ctypes
import random

import pytest
import torch

from prismer_b200 import _C, _C_decl, engine, ops, synthetic
from tests.helpers import TINY_DEC, build_model

EXPERTS = synthetic.DEFAULT_EXPERTS
STRUCT_CALLS = {"prismer_gemm_bf16": 2, "prismer_attention_fwd": 2,
                "prismer_attention_bwd": 2}


class _Recorder:
    def __init__(self):
        self.calls = {}

    def __getattr__(self, name):
        if not name.startswith("prismer_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls[name] = self.calls.get(name, 0) + 1
            if name in STRUCT_CALLS:
                assert len(args) == STRUCT_CALLS[name], (name, len(args))
                return 0
            sig = _C_decl.SIGNATURES[name]                       # KeyError = undeclared entry point
            assert len(args) == len(sig), f"{name}: {len(args)} arguments, {len(sig)} declared"
            for a, t in zip(args, sig):
                if t in (_C_decl.I, _C_decl.L, _C_decl.U32, _C_decl.U64):
                    assert isinstance(a, int) and not isinstance(a, bool) or isinstance(a, bool), (name, a, t)
                elif t is _C_decl.F:
                    assert isinstance(a, (int, float)), (name, a)
                else:
                    assert a is None or isinstance(a, int) or isinstance(a, ctypes._SimpleCData) or hasattr(a, "_obj"), (name, type(a))
            return 0
        return fn


@pytest.fixture()
def dry(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(_C, "lib", lambda: rec)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req_cuda", lambda *t: None)
    monkeypatch.setattr(engine, "_experts_check", lambda e: None)
    monkeypatch.setattr(engine, "SIDE_STREAM", False)
    return rec


def _tiny(train=True, dec=True):
    m, _ = build_model(256, 2, 16, 64, EXPERTS, TINY_DEC if dec else None, seed=3, device="cpu")
    engine.prepare(m, torch.device("cpu"))
    m.train(train)
    return m


def _batch(compact=False):
    ex = (synthetic.synth_compact_experts if compact else synthetic.synth_experts)(2, 64, EXPERTS, 64, 5)
    ids, mask = synthetic.synth_tokens(2, 8, TINY_DEC["vocab_size"], 5, ragged=True)
    labels = ids.masked_fill(ids == 1, -100)
    labels[:, :3] = -100
    return ex, ids, mask, labels


@pytest.mark.parametrize("compact,train", [(False, True), (True, True), (False, False)])
def test_train_step_host_code(dry, compact, train):
    m = _tiny(train)
    ex, ids, mask, labels = _batch(compact)
    random.seed(0)
    loss = engine.train_loss(m, ex, ids, mask, labels)
    loss.backward()
    assert m.text_decoder.lm_head.dense.weight.grad is not None
    assert dry.calls["prismer_gemm_bf16"] > 100 and dry.calls["prismer_attention_bwd"] > 5
    assert ("prismer_label_resample" in dry.calls) == compact and ("prismer_expand_labels" in dry.calls) == compact
    assert ("prismer_bn_relu_bwd_eval" in dry.calls) == (not train) and ("prismer_bn_relu_bwd" in dry.calls) == train


def test_generation_and_rank_host_code(dry):
    from prismer_b200.prismer_caption import rank
    m = _tiny(False)
    m.tokenizer = None
    ex, ids, mask, _ = _batch()
    with torch.no_grad():
        enc = m.expert_encoder(ex).transpose(0, 1)
        prefix = ids[:, :4].contiguous()
        g = m.text_decoder.generate(input_ids=prefix, encoder_hidden_states=enc, attention_mask=torch.ones_like(prefix), num_beams=1,
                                    max_length=8, min_length=6)
        b = m.text_decoder.generate(input_ids=prefix, encoder_hidden_states=enc, attention_mask=mask[:, :4], num_beams=3, max_length=8,
                                    min_length=6, length_penalty=-1)
    assert g.shape[0] == 2 and 4 < g.shape[1] <= 8 and b.shape[0] == 2 and 4 < b.shape[1] <= 8
    from types import SimpleNamespace
    m.tokenizer = SimpleNamespace(pad_token_id=1)
    ans, amask = synthetic.synth_tokens(5, 3, TINY_DEC["vocab_size"], 9, ragged=True)
    with torch.no_grad():
        r = rank(m, ex, ids[:, :4].contiguous(), mask[:, :4].contiguous(), ans, amask, 3)
    assert r.shape == (2,) and r.dtype == torch.int64


def test_optimizer_host_code(dry):
    from prismer_b200.optim import FusedAdamW
    m = _tiny(True)
    st = engine._store(m)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    opt.step()
    assert opt.t == 1 and dry.calls["prismer_adamw_step"] == 1
    opt.step_range(0, st.n_train_dec, True, False)           # decoder slice first (its gradients are final early) ...
    opt.step_range(st.n_train_dec, st.n_train, False, True)  # ... then the rest: one logical step
    assert opt.t == 2 and dry.calls["prismer_adamw_step"] == 3 and 0 < st.n_train_dec < st.n_train
    assert dry.calls.get("prismer_conv_weight_pack", 0) > 0   # trainable conv weights were re-packed for the next forward


@pytest.mark.parametrize("compact,stock_adamw", [(False, False), (True, True)])
def test_reference_training_loop_host_code(dry, monkeypatch, compact, stock_adamw):
    """The loop body of train_caption.py:126-133 as examples/train_caption_synthetic.py runs it: dataset -> DataLoader ->
    accelerator.prepare -> model(experts, caption, prefix=...) -> accelerator.backward -> optimizer.step, then beam-3 generate."""
    import os
    from torch.utils.data import DataLoader
    from prismer_b200.accelerate_shim import Accelerator
    from prismer_b200.optim import FusedAdamW
    from prismer_b200.prismer_caption import PrismerCaption
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    tiny = {"roberta_model": dict(TINY_DEC, model_name="roberta-tiny"), "vit_model": "tiny", "vit_dims": [16, 256, 2]}
    config = {"experts": EXPERTS, "prismer_model": "tiny", "image_resolution": 64, "freeze": "freeze_vision", "prismer_config": tiny,
              "prefix": "A picture of"}
    accelerator = Accelerator(mixed_precision="bf16")
    assert accelerator.device.type == "cpu"                          # no GPU here: the recorder stands in for the CUDA library
    model = PrismerCaption(config)
    if stock_adamw:                                                  # train_caption.py:111-117 order: optimizer first, then prepare
        optimizer = torch.optim.AdamW(params=filter(lambda p: p.requires_grad, model.parameters()), lr=5e-5, weight_decay=0.05)
        model = accelerator.prepare(model)
    else:
        model = accelerator.prepare(model)
        optimizer = accelerator.prepare(FusedAdamW(model, lr=5e-5, weight_decay=0.05))
    ds = synthetic.SyntheticCaptionDataset(4, EXPERTS, 64, 64, compact=compact, prefix=config["prefix"])
    loader = accelerator.prepare(DataLoader(ds, batch_size=2, collate_fn=ds.collate, shuffle=True, drop_last=True))
    model.train()
    steps = 0
    for experts, caption in loader:
        loss = model(experts, caption, prefix=config["prefix"])
        optimizer.zero_grad()
        accelerator.backward(loss)
        if stock_adamw:
            for p in model.parameters():                             # uninitialised "gradients" may hold NaN: keep the stock update finite
                if p.grad is not None:
                    p.grad.zero_()
        optimizer.step()
        steps += 1
    assert steps == 2 and dry.calls["prismer_gemm_bf16"] > 200
    frozen = [p for n, p in model.named_parameters() if "transformer.resblocks" in n and "adaptor" not in n]
    assert frozen and all(not p.requires_grad and p.grad is None for p in frozen)        # freeze_vision (prismer.py:45-49)
    model.eval()
    with torch.no_grad():
        captions = model(experts, train=False, prefix=config["prefix"])
    assert isinstance(captions, list) and len(captions) == 2 and all(isinstance(c, str) for c in captions)


class _FakeStream:
    def __init__(self, *a, **k): pass
    def wait_stream(self, other): pass
    def wait_event(self, e): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


class _FakeEvent:
    def __init__(self, *a, **k): pass
    def record(self, stream=None): pass
    def synchronize(self): pass
    def elapsed_time(self, other): return 1.0


class _FakeGraph:
    replays = 0
    def pool(self): return None
    def replay(self): _FakeGraph.replays += 1


@pytest.fixture()
def fake_cuda(monkeypatch):
    """Just enough of torch.cuda for GraphedTrainStep's control flow (streams, graph capture / replay) on a CPU-only host."""
    import contextlib
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _FakeGraph)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "graph", lambda g, pool=None, **kw: contextlib.nullcontext())
    _FakeGraph.replays = 0


@pytest.mark.parametrize("overlap,compact", [(False, False), (True, False), (True, True)])
def test_graphed_train_step_host_code(dry, fake_cuda, overlap, compact):
    from prismer_b200.optim import FusedAdamW
    m = _tiny(True)
    st = engine._store(m)
    opt = FusedAdamW(m, lr=1e-3)
    ex, ids, mask, labels = _batch(compact)
    graphed = engine.GraphedTrainStep(m, ex, ids, mask, labels, overlap=overlap)
    ex2, ids2, mask2, labels2 = _batch(compact)
    graphed.load_inputs(ex2, ids2, mask2, labels2)                   # new batch into the static buffers (host or device tensors)
    handles = []
    comm = lambda t: (handles.append(t.numel()), type("H", (), {"wait": lambda self: None})())[1]
    loss = graphed(comm)
    assert loss is graphed.loss and _FakeGraph.replays == (3 if overlap else 1)     # fwd + decoder bwd | encoder bwd | expert stems
    assert 0 < st.n_train_dec < st.n_train_late < st.n_train
    late = [n for n, p in m.named_parameters() if p.requires_grad and engine._store(m)._offset[id(p)][1] >= st.n_train_late]
    assert late and all("conv1." in n and "conv1.rgb" not in n or "instance_embedding" in n for n in late), late
    assert handles == ([st.n_train_dec, st.n_train_late - st.n_train_dec, st.grad_t.numel() - st.n_train_late] if overlap
                       else [st.grad_t.numel()])
    if overlap:                                                      # optimizer update of the decoder slice during the encoder backward
        n0 = dry.calls.get("prismer_adamw_step", 0)
        graphed(comm, on_decoder_grads=lambda: opt.step_range(0, st.n_train_dec, True, False))
        opt.step_range(st.n_train_dec, st.n_train, False, True)
        assert dry.calls["prismer_adamw_step"] == n0 + 2 and opt.t == 1
    assert all(p.grad is not None for p in m.parameters() if p.requires_grad)       # published views of the flat buffer


@pytest.mark.parametrize("patch,res,experts", [(14, 56, EXPERTS), (16, 64, []), (14, 56, ["normal", "edge", "ocr_detection"])])
def test_other_encoder_configurations_host_code(dry, patch, res, experts):
    """Patch-14 models (experts resampled 64 -> 73 / 64 -> 18, positional embedding resized for the expert tokens) and PrismerZ
    (rgb only: no stems, no resampler)."""
    m, _ = build_model(256, 2, patch, res, experts, TINY_DEC, seed=3, device="cpu")
    engine.prepare(m, torch.device("cpu"))
    m.train()
    ex = synthetic.synth_experts(2, res, experts, 64, 5)
    ids, mask = synthetic.synth_tokens(2, 8, TINY_DEC["vocab_size"], 5, ragged=True)
    labels = ids.masked_fill(ids == 1, -100)
    random.seed(0)
    engine.train_loss(m, ex, ids, mask, labels).backward()
    assert ("prismer_resample_bilinear" in dry.calls) == (len(experts) > 0)
    with torch.no_grad():
        enc = m.expert_encoder(ex)
    n_lat = 64 if experts else 0
    assert enc.shape == ((res // patch) ** 2 + n_lat, 2, 256)


def test_vqa_training_and_inference_host_code(dry):
    """train_vqa.py:125 / :161 call shapes through the real engine host code: weighted loss, rank over a candidate list, beam generate."""
    from prismer_b200.prismer_vqa import PrismerVQA
    tiny = {"roberta_model": dict(TINY_DEC, model_name="roberta-tiny"), "vit_model": "tiny", "vit_dims": [16, 256, 2]}
    m = PrismerVQA({"experts": EXPERTS, "prismer_model": "tiny", "image_resolution": 64, "freeze": "freeze_lang_vision", "prismer_config": tiny})
    engine.prepare(m, torch.device("cpu"))
    ex = synthetic.synth_experts(2, 64, EXPERTS, 64, 5)
    qs, ans = ["what is on the table", "how many dogs are there in the picture"], ["a cup", "two"]
    m.train()
    random.seed(0)
    loss = m(ex, qs, ans, weights=torch.tensor([1.0, 0.5]))
    loss.backward()
    named = dict(m.named_parameters())
    assert named["text_decoder.roberta.encoder.layer.0.1.self.query.weight"].grad is not None        # cross-attention stays trainable
    assert named["text_decoder.roberta.encoder.layer.0.0.attention.self.query.weight"].grad is None  # freeze_lang_vision (prismer.py:50-56)
    m.eval()
    with torch.no_grad():
        r = m(ex, qs, ["a cup", "two", "pizza", "a red ball", "none"], train=False, inference="rank", k_test=3)
        g = m(ex, qs, train=False, inference="generate")
    assert r.shape == (2,) and len(g) == 2 and all(isinstance(s, str) for s in g)


def test_refresh_sees_updates_made_through_the_parameters(dry):
    """a stock torch optimizer / load_state_dict write the masters through the nn.Parameter views, whose version
    counters are not the flat buffers' -- refresh() must re-cast the bf16 compute copies after either, and must not when idle."""
    m = _tiny(True)
    st = m._prismer_store
    st.refresh()
    n0 = dry.calls.get("prismer_cast_f32_bf16", 0)
    st.refresh()
    assert dry.calls.get("prismer_cast_f32_bf16", 0) == n0                      # nothing changed -> no cast
    p = m.text_decoder.lm_head.dense.weight
    p.grad = torch.ones_like(p)
    torch.optim.AdamW([p], lr=1e-2).step()
    st.refresh()
    n1 = dry.calls.get("prismer_cast_f32_bf16", 0)
    assert n1 > n0, "stock optimizer step not detected"
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    st.refresh()
    assert dry.calls.get("prismer_cast_f32_bf16", 0) > n1, "load_state_dict on a prepared model not detected"
    n2 = dry.calls.get("prismer_cast_f32_bf16", 0)
    with torch.no_grad():
        st.master_t.mul_(1.0)                                                      # flat-buffer writes (broadcast, fused optimizer) still count
    st.refresh()
    assert dry.calls.get("prismer_cast_f32_bf16", 0) > n2


def test_accelerator_save_state_load_state_roundtrip(dry, tmp_path):
    """Checkpoint resume (train_caption.py:96-109,173-176; round-1 VERDICT missing #8): weights, fused-optimizer moments / step / lr, RNG
    streams and the device-side dropout key survive save_state -> load_state into freshly built objects."""
    import random
    from prismer_b200.accelerate_shim import Accelerator
    from prismer_b200.optim import FusedAdamW
    acc = Accelerator()
    m = _tiny(True)
    opt = FusedAdamW(m, lr=5e-5, weight_decay=0.05)
    m, opt = acc.prepare(m, opt)
    opt.m.normal_(); opt.v.uniform_(); opt.t = 7
    opt.param_groups[0]["lr"] = 1.25e-5
    engine._store(m).seed.fill_(4321)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.01)
    random.seed(99); torch.manual_seed(98)
    acc.save_state(str(tmp_path))
    want_r, want_t = random.random(), torch.rand(3)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    acc2 = Accelerator()
    m2 = _tiny(True)
    opt2 = FusedAdamW(m2, lr=1.0)
    m2, opt2 = acc2.prepare(m2, opt2)
    acc2.load_state(str(tmp_path))
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v) and opt2.t == 7 and opt2.param_groups[0]["lr"] == 1.25e-5
    assert int(engine._store(m2).seed.item()) == 4321
    assert random.random() == want_r and torch.equal(torch.rand(3), want_t)


def test_optimizer_state_survives_a_layout_change(dry):
    """The flat buffers follow the backward's completion order (decoder | ViT + resampler | stems); a checkpoint written under another
    order is re-mapped parameter by parameter through the names stored with it, and a foreign one fails loudly."""
    from prismer_b200.optim import FusedAdamW
    m = _tiny(True)
    opt = FusedAdamW(m, lr=1e-3)
    opt.m.normal_(); opt.v.uniform_(); opt.t = 3
    sd = opt.state_dict()
    lay = sd["layout"]
    assert [n for n, _, _ in lay] == [n for n, p in sorted(((n, p) for n, p in m.named_parameters() if p.requires_grad),
                                                           key=lambda np_: engine._store(m)._offset[id(np_[1])][1])]
    # the same moments laid out in reverse parameter order
    rev, off = [], 0
    m_rev, v_rev = torch.zeros_like(opt.m), torch.zeros_like(opt.v)
    for n, o, k in reversed(lay):
        m_rev[off:off + k] = opt.m[o:o + k]; v_rev[off:off + k] = opt.v[o:o + k]
        rev.append((n, off, k)); off += (k + 7) // 8 * 8
    opt2 = FusedAdamW(m, lr=1e-3)
    opt2.load_state_dict({"m": m_rev, "v": v_rev, "t": 3, "layout": rev, "param_groups": sd["param_groups"]})
    for n, o, k in lay:
        assert torch.equal(opt2.m[o:o + k], opt.m[o:o + k]) and torch.equal(opt2.v[o:o + k], opt.v[o:o + k]), n
    assert opt2.t == 3
    with pytest.raises(KeyError):
        opt2.load_state_dict({"m": m_rev, "v": v_rev, "t": 3, "layout": [("not.a.parameter", 0, 8)], "param_groups": []})
    with pytest.raises(ValueError):
        opt2.load_state_dict({"m": m_rev[:8], "v": v_rev[:8], "t": 3, "param_groups": []})
