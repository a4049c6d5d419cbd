#!/usr/bin/env python
"""The reference's caption fine-tune loop (train_caption.py:111-136) running on prismer_b200 -- same loop body, same config
keys, string captions in, synthetic data instead of COCO (no datasets here).

    python examples/train_caption_synthetic.py --steps 20                       # 1 GPU
    torchrun --nproc-per-node 8 examples/train_caption_synthetic.py             # data parallel, one all-reduce per step
"""
import argparse
import math
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prismer_b200 import synthetic                                   # noqa: E402
from prismer_b200.accelerate_shim import Accelerator                 # reference: from accelerate import Accelerator
from prismer_b200.optim import FusedAdamW                            # reference: torch.optim.AdamW (also works)
from prismer_b200.prismer_caption import PrismerCaption              # reference: from model.prismer_caption import PrismerCaption


def cosine_lr_schedule(optimizer, epoch, max_epoch, init_lr, min_lr):
    """utils.py:13-17 of the reference."""
    lr = (init_lr - min_lr) * 0.5 * (1. + math.cos(math.pi * epoch / max_epoch)) + min_lr
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--seed', type=int, default=42)
    ap.add_argument('--torch_adamw', action='store_true')
    ap.add_argument('--compact', action='store_true', help="uint8 label maps + tables through the loader (prismer_b200.data)")
    args = ap.parse_args()
    config = {'experts': synthetic.DEFAULT_EXPERTS, 'image_resolution': 224, 'prismer_model': 'prismer_base', 'freeze': 'freeze_vision',
              'batch_size_train': args.batch, 'init_lr': 5e-5, 'weight_decay': 0.05, 'min_lr': 0, 'max_epoch': 1,
              'prefix': 'A picture of'}                                         # configs/caption.yaml
    torch.manual_seed(args.seed); np.random.seed(args.seed); random.seed(args.seed)

    accelerator = Accelerator(mixed_precision='bf16')
    model = PrismerCaption(config)
    if args.torch_adamw:
        model.to(accelerator.device)
        optimizer = torch.optim.AdamW(params=filter(lambda p: p.requires_grad, model.parameters()), lr=config['init_lr'],
                                      weight_decay=config['weight_decay'])
        model = accelerator.prepare(model)
    else:
        model = accelerator.prepare(model)
        optimizer = accelerator.prepare(FusedAdamW(model, lr=config['init_lr'], weight_decay=config['weight_decay']))

    # reference: train_dataset, test_dataset = create_dataset('caption', config); train_loader = create_loader(...)  (train_caption.py:49-52)
    dataset = synthetic.SyntheticCaptionDataset(args.steps * args.batch * accelerator.num_processes, config['experts'],
                                                config['image_resolution'], compact=args.compact, prefix=config['prefix'], seed=args.seed)
    train_loader = torch.utils.data.DataLoader(dataset, batch_size=config['batch_size_train'], num_workers=4, pin_memory=True,
                                               collate_fn=dataset.collate, shuffle=True, drop_last=True)
    train_loader = accelerator.prepare(train_loader)                 # this rank's batches, already on the device (train_caption.py:115)
    model.train()
    for i, (experts, caption) in enumerate(train_loader):
        cosine_lr_schedule(optimizer, i, len(train_loader), config['init_lr'], config['min_lr'])

        loss = model(experts, caption, prefix=config['prefix'])

        optimizer.zero_grad()
        accelerator.backward(loss)
        optimizer.step()
        accelerator.print(f"step {i:3d}  loss {loss.item():.4f}  lr {optimizer.param_groups[0]['lr']:.2e}")

    model.eval()
    with torch.no_grad():
        captions = model(experts, train=False, prefix=config['prefix'])
    accelerator.print("generated:", captions[:2])


if __name__ == '__main__':
    main()
