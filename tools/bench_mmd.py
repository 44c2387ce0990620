#!/usr/bin/env python
"""Times the fused MMD (BASELINE.json configs[2]: 8192 latents vs 8192 samples, D = 5640) stage by stage:
spin extraction, the one-pass Hamming-histogram forward, the coefficient pass and the int8 GEMM of the backward."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_generation_b200 as B
from image_generation_b200 import _lib
from image_generation_b200.mmd import mmd_block_sums
from image_generation_b200.mmd_tc import (mmd_backward_i8, mmd_block_sums_i8, mmd_histograms_i8, pack_pair_i8, pack_rows_i8)

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=8192)
ap.add_argument("--d", type=int, default=5640)
ap.add_argument("--path", default="i8")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--bandwidth", type=float, default=0.0)
ap.add_argument("--stage", default="all", help="all | forward (only the Gram forward: what the ncu capture wants) | backward")
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
z = (torch.randint(0, 2, (2 * args.m, args.d), generator=g, dtype=torch.int8) * 2 - 1)
z[args.m:, : args.d // 8] = 1
z = z.to(dev)
kern = B.GaussianKernel(7, bandwidth=args.bandwidth if args.bandwidth > 0 else None).to(dev)


def timed(fn, iters=args.iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


out = {"m": 2 * args.m, "d": args.d, "path": args.path}
if args.path != "i8":
    zz = z.float()
    out["forward_ms"] = timed(lambda: mmd_block_sums(zz, args.m, kern, path=args.path))
    print(json.dumps(out)); sys.exit(0)

zi, _ = pack_rows_i8(z)                       # resident kernel layout: row pitch padded to whole 128-byte lines
if args.stage in ("all", "forward"):
    out["forward_ms"] = timed(lambda: mmd_block_sums_i8(zi, args.m, kern, d=args.d))
    hist = torch.zeros((3, args.d + 1), dtype=torch.int64, device=dev)
    out["gram_hist_ms"] = timed(lambda: mmd_histograms_i8(zi, args.m, args.d, hist=hist))
    m = 2 * args.m
    tiles = (m // 256) * (m // 256 + 1)
    out["executed_TOPs"] = 2.0 * tiles * 128 * 256 * zi.shape[1] / out["gram_hist_ms"] / 1e9
if args.stage in ("all", "backward"):
    x, y = z[: args.m].float(), z[args.m:].float()
    out["spin_extract_fwd_only_ms"] = timed(lambda: pack_pair_i8(x, y))
    out["spin_extract_with_transpose_ms"] = timed(lambda: pack_pair_i8(x, y, need_grad=True))
    pair = pack_pair_i8(x, y, need_grad=True)
    sums = mmd_block_sums_i8(pair.rows, args.m, kern, d=args.d)
    w_xx, w_xy = 2.0 / (args.m * (args.m - 1)), -2.0 / (args.m * args.m)
    one = torch.ones((), device=dev)
    for planes in (2, 3):
        out[f"backward_{planes}planes_ms"] = timed(lambda: mmd_backward_i8(pair.rows, args.d, args.m, kern, sums, w_xx, w_xy, one,
                                                                              zt=pair.zt, n_planes=planes))
    xg = x.clone().requires_grad_(True)

    def loss_call():
        xg.grad = None
        B.maximum_mean_discrepancy_loss(xg, y, kern).backward()
    out["loss_call_fwd_bwd_auto_ms"] = timed(loss_call)
print(json.dumps(out))
