#!/usr/bin/env python
"""Times the fused MMD (BASELINE.json configs[2]: 8192 latents vs 8192 samples, D = 5640)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_generation_b200 as B
from image_generation_b200.mmd import mmd_block_sums

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=8192)
ap.add_argument("--d", type=int, default=5640)
ap.add_argument("--path", default="i8")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--bandwidth", type=float, default=0.0)
ap.add_argument("--zeros", action="store_true", help="all-zero operands: same work, minimal switching power (is the kernel power-limited?)")
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
z = (torch.randint(0, 2, (2 * args.m, args.d), generator=g, dtype=torch.int8) * 2 - 1).to(dev)
if args.zeros:
    z.zero_()
kern = B.GaussianKernel(7, bandwidth=args.bandwidth if args.bandwidth > 0 else None).to(dev)
if args.path == "i8":
    from image_generation_b200.mmd_tc import mmd_block_sums_i8, pack_rows_i8
    zi, _ = pack_rows_i8(z)                       # resident kernel layout: row pitch padded to whole 128-byte lines
    mmd_block_sums = lambda zz, m_x, kern, path: mmd_block_sums_i8(zz, m_x, kern, d=args.d)
    zz = zi
else:
    zz = z.float()
for _ in range(2):
    s = mmd_block_sums(zz, args.m, kern, path=args.path)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
for a, b in ev:
    a.record(); s = mmd_block_sums(zz, args.m, kern, path=args.path); b.record()
torch.cuda.synchronize()
ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
m = 2 * args.m
passes = 1 if args.bandwidth > 0 else 2
flops = 2.0 * m * m * args.d * passes          # as the reference computes it (full stacked matrix), per pass
print(json.dumps({"path": args.path, "m": m, "d": args.d, "ms": ms, "passes": passes,
                  "tflops_full_matrix_equiv": flops / ms / 1e9, "input_GBps": m * args.d * (1 if args.path == "i8" else 4) / ms / 1e6,
                  "sums": s.cpu().tolist()}))
