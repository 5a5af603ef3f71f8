"""Seeded synthetic inputs shaped like the BASELINE.json configs."""
import numpy as np
import torch


def class_map(h, w, n_cls, tile, gen):
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    small = torch.randint(0, n_cls, (th, tw), generator=gen)
    return small.repeat_interleave(tile, 0).repeat_interleave(tile, 1)[:h, :w].contiguous()


def streethazards_like(n_img, h, w, k=13, sigma=0.7, ood_label=13, n_discs=5, seed=1, tile=64, ignore_rows=0):
    """Embeddings x = 3 e_c + sigma N(0,1) on a block-constant class map; OOD discs (label `ood_label`,
    x = sigma N(0,1)); optional ignored (-1) top rows.  Returns (x [n,k,h,w] fp32, gt [n,h,w] int64)."""
    gen = torch.Generator().manual_seed(seed)
    xs, gts = [], []
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    for _ in range(n_img):
        cm = class_map(h, w, k, tile, gen)
        x = torch.randn(k, h, w, generator=gen) * sigma
        ood = torch.zeros(h, w, dtype=torch.bool)
        r = max(2, int(0.025 * min(h, w)) * 2)
        for _d in range(n_discs):
            cy = int(torch.randint(0, h, (1,), generator=gen))
            cx = int(torch.randint(0, w, (1,), generator=gen))
            ood |= (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        onehot = torch.nn.functional.one_hot(cm, k).permute(2, 0, 1).float() * 3.0
        x = x + onehot * (~ood).float()
        gt = cm.clone()
        gt[ood] = ood_label
        if ignore_rows:
            gt[:ignore_rows] = -1
        xs.append(x)
        gts.append(gt)
    return torch.stack(xs).contiguous(), torch.stack(gts).contiguous()
