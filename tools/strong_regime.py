"""One-GPU emulation of the pooled stage in the strong-scaling regime: the negatives of 1500/N images are bucketed
against the distinct positive scores of ALL 1500 images (what every rank of an N-GPU run does).  Prints the
CUDA-event time of dml_ood_bucket_rank per shard size; run under ncu for the per-kernel split."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from dml_b200 import ood

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=1500)
ap.add_argument("--shards", default="187,375,750,1500")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--positive-passes", type=int, default=1, help="positives of this many x --images images (weak-scaling regime: every rank ranks against the positives of all ranks)")
a = ap.parse_args()
sys.argv = [sys.argv[0], "--images", str(a.images)]
args = bench.parse_args()
dev = torch.device("cuda:0")
pipe = bench.Pipeline(args, dev, 0, 1)
gen = torch.Generator(device=dev).manual_seed(1234)
ws = ood.OodWorkspace(dev)
all_pos = []
for pss in range(a.positive_passes):
    pipe.begin_step()
    for ci, (s, e) in enumerate(pipe.bounds):
        x, gt = bench.synth_chunk_torch(e - s, pipe.k, pipe.h, pipe.w, gen, dev)
        pipe.process_chunk(ci, x, gt)
        del x, gt
    pool = pipe.pool
    n_pos = int(pool.pos_count.view(-1)[0].item())
    all_pos.append(pool.pos[:n_pos].clone())
allp = torch.cat(all_pos)
n_pos = allp.numel()
del all_pos
srt = ood.sort_keys(allp, ws, "pr_sort", end_bit=31)
S, pc, G = ood.unique_groups(srt, ws)
print("positives", n_pos, "groups", G, flush=True)
for sub in [int(v) for v in a.shards.split(",")]:
    keys = pool.keys[: sub * pipe.hw]
    for _ in range(2):
        ood.bucket_rank_counters(keys, S, ood.KEY_BASE_NONNEG, ws)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.reps):
        ood.bucket_rank_counters(keys, S, ood.KEY_BASE_NONNEG, ws)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.reps
    print(f"shard {sub} images: bucket_rank {ms:.3f} ms  ({ms / sub * 1500:.2f} ms per 1500 images)", flush=True)
