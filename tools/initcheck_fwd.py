import random, sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import engine, synthetic
from prismer_b200.prismer_caption import PrismerCaption
from tests.test_surface_gpu import _model, _experts, TINY_DEC
m = _model(PrismerCaption)
m.expert_encoder.train(); m.text_decoder.eval()
ex = _experts(2)
ids, mask = synthetic.synth_tokens(2, 8, TINY_DEC["vocab_size"], 5, ragged=True)
ids, mask = ids.cuda(), mask.cuda()
labels = ids.masked_fill(ids == 1, -100); labels[:, :3] = -100
random.seed(1)
loss = engine.train_loss(m, ex, ids, mask, labels)
if len(sys.argv) > 1:
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
