"""Pins the summation order that ``dml_head(..., reference_order=True)`` reproduces (csrc/dml_head.cuh, parity mode).

The reference forms its distance logits with torch ops (anomaly/models/models.py:649-651): ``dists = features - centers``,
``dists ** 2``, ``torch.sum(..., 3)`` over a contiguous inner dimension of D floats.  On the CPU build of this image
torch adds, one rounded fp32 add at a time, for 8 <= D < 16 the tail elements 8 .. D-1 first and then 0 .. 7, and for
D < 8 element 0, then the tail 4 .. D-1, then 1 .. 3 (recovered by probing with 2^24 / 1 / 1 triples); the class-plane sum of the EDS score (``torch.sum(scores, dim=1)``,
anomaly/eval_ood_traditional.py:302) runs k = 0, 1, 2, ... for K <= 17 at realistic map sizes.  The kernel hard-codes
both orders; if a torch build ever changes them this test fails first (and the golden fixtures would have to be
regenerated with it)."""
import numpy as np
import pytest
import torch


def _kernel_order(d):
    if d >= 8:
        return list(range(8, d)) + list(range(8))
    if d > 4:
        return [0] + list(range(4, d)) + [1, 2, 3]
    return list(range(d))


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15])
def test_inner_sum_order_of_the_distance_logits(d):
    g = torch.Generator().manual_seed(d)
    x = torch.randn(20000, d, generator=g) * 3
    c = torch.eye(d) * 3
    diff = x.unsqueeze(1).expand(-1, d, -1) - c
    ref = torch.sum(diff ** 2, 2)
    sq = diff ** 2
    acc = None
    for i in _kernel_order(d):
        acc = sq[..., i].clone() if acc is None else acc + sq[..., i]
    assert torch.equal(acc, ref)


@pytest.mark.parametrize("k", [2, 13, 16, 17])
def test_class_plane_sum_order_of_the_eds_score(k):
    g = torch.Generator().manual_seed(k)
    z = -(torch.randn(1, k, 96, 160, generator=g) * 30 + 60)
    ref = torch.sum(z, dim=1)
    acc = z[:, 0].clone()
    for j in range(1, k):
        acc = acc + z[:, j]
    assert torch.equal(acc, ref)
