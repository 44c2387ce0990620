#!/usr/bin/env python
"""Fuzz of the tcgen05 MMD path against the CPU oracle:  python tests/fuzz_mmd.py SEED SECONDS  (test infrastructure:
it runs the oracle)  (one B200).

Random shapes (1..700 rows per side incl. ragged tile edges, D = 5..900 incl. widths that are no multiple of 4 / 16 /
128), rows with structure (shared prefixes, duplicated rows, all-equal clouds), every kernel switch (squared, fixed /
auto bandwidth, reduce, estimator), tile sharding over 1..5 "ranks", the bit-row exchange layout.  Checked per case:
the three Hamming histograms equal the oracle's count for count; block sums / estimate against the float64 oracle;
the loss value and (2- and 3-plane) gradient through the reference's unmodified call against the oracle.
Round 2: 2953 cases in 160 s (seeds 1 and 2), all green; tests/test_mmd_gpu.py runs a 10-second slice."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import image_generation_b200 as B
from image_generation_b200 import mmd_tc
from image_generation_b200.dist import _DeviceOps as ops
from oracle import oracle as O

dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
t0 = time.time(); n_case = 0


def cloud(rows, d):
    z = rng.choice([-1.0, 1.0], size=(rows, d))
    kind = rng.integers(0, 5)
    if kind == 1:
        z[:, : d // 3] = 1.0                       # shared prefix: narrow distance distribution
    elif kind == 2 and rows > 2:
        z[rows // 2:] = z[: rows - rows // 2]      # duplicated rows: distance 0 off the diagonal
    elif kind == 3:
        z[:] = z[0]                                # one point
    elif kind == 4:
        z *= rng.choice([-1.0, 1.0], size=(1, d)) * np.sign(rng.normal(size=(rows, 1)) + 1.0)   # a cloud and its mirror
    return z.astype(np.float32)


while time.time() - t0 < budget:
    m_x = int(rng.choice([2, 3, 17, 127, 128, 129, 255, 300, 513, 700]))
    m_y = int(rng.choice([2, 5, 64, 128, 131, 256, 400, 640]))
    d = int(rng.choice([5, 16, 31, 77, 128, 130, 256, 333, 640, 900]))
    x, y = cloud(m_x, d), cloud(m_y, d)
    squared = bool(rng.integers(0, 2)); reduce = str(rng.choice(["sum", "mean"])); est = str(rng.choice(["unbiased", "biased"]))
    bw = None if rng.integers(0, 2) else float(rng.uniform(0.5, 3.0) * (2.0 * d if squared else np.sqrt(2.0 * d)))
    kern = B.GaussianKernel(7, bandwidth=bw, squared=squared, reduce=reduce).to(dev)
    ctx = dict(m_x=m_x, m_y=m_y, d=d, squared=squared, reduce=reduce, est=est, bw=bw)
    z_np = np.concatenate([x, y]).astype(np.int64)
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    pair = mmd_tc.pack_pair_i8(xt, yt)
    # (1) histograms, whole and sharded
    want_h = O.hamming_histograms(z_np, m_x)
    got_h = mmd_tc.mmd_histograms_i8(pair.rows, m_x, d).cpu().numpy()
    if not np.array_equal(got_h, want_h):
        print("HISTOGRAM MISMATCH", ctx); sys.exit(1)
    world = int(rng.integers(1, 6))
    parts = sum(mmd_tc.mmd_histograms_i8(pair.rows, m_x, d, (r, world)).cpu().numpy() for r in range(world))
    if not np.array_equal(parts, want_h):
        print("SHARD MISMATCH", ctx, world); sys.exit(1)
    # (2) the bit-row layout assembles the same matrix (rows dealt to `world` ranks when they divide)
    if m_x % world == 0 and m_y % world == 0:
        a, b = m_x // world, m_y // world
        every = torch.stack([ops.pack_bits(xt[r * a:(r + 1) * a], yt[r * b:(r + 1) * b]) for r in range(world)])
        if not torch.equal(ops.unpack_bits(every, a, b, d), pair.rows):
            print("BIT ROW MISMATCH", ctx, world); sys.exit(1)
    # (3) value and gradient through the reference's call (auto dispatch -> int8 tensor cores)
    degenerate = want_h[:, 1:].sum() == 0 and bw is None            # all distances zero: auto bandwidth 0 (0/0 in any form)
    if not degenerate:
        want_v, want_g = O.mmd(x, y, bandwidth=bw, squared=squared, reduce=reduce, estimator=est, return_grad=True)
        xg = xt.clone().requires_grad_(True)
        # fixed-point digits of the backward coefficients: 2 planes = 16 bits of the largest coefficient (the default; an
        # entry's error is bounded by rows x 2^-17 of it, so sums with heavy cancellation see up to a few 1e-3 of the
        # largest gradient entry), 3 planes = 24 bits
        mmd_tc.GRAD_PLANES = planes = int(rng.choice([2, 3]))
        val = B.maximum_mean_discrepancy_loss(x=xg, y=yt, kernel=kern, estimator=est)
        val.backward()
        scale = 7.0 if reduce == "sum" else 1.0
        if B.mmd.last_path != "i8" or abs(float(val.detach()) - want_v) > 2e-5 * scale:
            print("VALUE MISMATCH", ctx, float(val), want_v, B.mmd.last_path); sys.exit(1)
        gmax = np.abs(want_g).max()
        if gmax > 0 and np.abs(xg.grad.cpu().numpy() - want_g).max() > (1e-2 if planes == 2 else 5e-5) * gmax + 1e-12:
            print("GRADIENT MISMATCH", ctx, planes, np.abs(xg.grad.cpu().numpy() - want_g).max(), gmax); sys.exit(1)
    n_case += 1
print("fuzz OK:", n_case, "cases")
