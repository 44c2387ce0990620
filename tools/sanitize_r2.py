#!/usr/bin/env python
"""Tiny launches of the round-2 sweep kernels for compute-sanitizer (tools/sanitize.sh r2):
gibbs_wide_kernel Pegasus form (pre-drawn uniforms, split mbarrier round barrier), Zephyr form (one stage, two CTAs
per SM), gibbs_small_kernel (producer warps), unpack_state_kernel -- each against the generic kernel's samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_generation_b200 as B

if os.environ.get("B200GRBM_LIB"):          # experiment build (tools/build_variant.sh or a -D variant of gibbs_wide.cu)
    from image_generation_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(os.environ["B200GRBM_LIB"])
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {"p16": (B.IsingGraph.pegasus(16), 4096, 2, "wide"), "z15": (B.IsingGraph.zephyr(15), 8288, 1, "wide"),
         "p3": (B.IsingGraph.pegasus(3), 64, 5, "small")}
for name, (g, chains, sweeps, want) in cases.items():
    if which not in ("all", name):
        continue
    rng = np.random.default_rng(3)
    h = (0.05 * rng.uniform(-0.5, 0.5, g.n)).astype(np.float32)
    J = (0.05 * rng.uniform(-5, 5, g.n_edges)).astype(np.float32)
    s = B.BlockGibbsSampler(g, device=dev)
    a = s.sample_ising(h, J, num_reads=chains, num_sweeps=sweeps, seed=5).record.sample.copy()
    kernel = s.last_kernel
    os.environ["B200GRBM_WIDE"] = "0"
    os.environ["B200GRBM_SMALL"] = "0"
    b = B.BlockGibbsSampler(g, device=dev).sample_ising(h, J, num_reads=chains, num_sweeps=sweeps, seed=5).record.sample
    del os.environ["B200GRBM_WIDE"], os.environ["B200GRBM_SMALL"]
    print(name, "kernel", kernel, "plan", s.last_plan, "equal to generic:", bool(np.array_equal(a, b)))
    assert kernel == want and np.array_equal(a, b)
