#!/usr/bin/env python
"""Side configurations of BASELINE.json (not bench lines): Zephyr Z15 shard of cfg4, annealed cfg2."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_generation_b200 as B

ap = argparse.ArgumentParser()
ap.add_argument("--graph", default="z15")
ap.add_argument("--chains", type=int, default=32768)
ap.add_argument("--sweeps", type=int, default=100)
ap.add_argument("--accept", default="exact")
ap.add_argument("--anneal", action="store_true")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--cpl", type=int, default=0, help="chains per lane (0 = planner)")
ap.add_argument("--threads", type=int, default=0, help="CTA size (0 = planner)")
ap.add_argument("--lib", default="", help="experiment build of the library (tools/build_variant.sh)")
args = ap.parse_args()
if args.lib:
    from image_generation_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(args.lib)
dev = torch.device("cuda:0")
if args.graph == "cfg1":       # the 256-spin Advantage2 subgraph of the reference's default checkpoint (BASELINE configs[0])
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "grbm_checkpoints.npz"))
    name = "Advantage2_system1_10_epochs"
    g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
else:
    import re
    m = re.fullmatch(r"([pz])(\d+)", args.graph)
    g = (B.IsingGraph.zephyr if m.group(1) == "z" else B.IsingGraph.pegasus)(int(m.group(2)))
rng = np.random.default_rng(0)
h = (0.05 * rng.uniform(-0.05, 0.05, g.n)).astype(np.float32)
J = (0.05 * rng.uniform(-5, 5, g.n_edges)).astype(np.float32)
s = B.BlockGibbsSampler(g, device=dev, accept=args.accept, beta_range=(0.1, 1.0) if args.anneal else None)
s.device_graph.set_weights(torch.from_numpy(h).to(dev), torch.from_numpy(J).to(dev))
plan = None
if args.cpl or args.threads:
    from image_generation_b200.sampler import plan_launch
    auto = plan_launch(args.chains, np.diff(g.colour_start).tolist(), 148, g.n, g.ell_width)
    plan = (args.cpl or auto[0], args.threads or s.device_graph.default_threads)
out = (torch.empty((args.chains, g.n), dtype=torch.int8, device=dev), torch.empty(args.chains, dtype=torch.float64, device=dev))
for _ in range(2):
    s._run(args.chains, args.sweeps, None, None, None, None, None, None, out=out, plan=plan)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
for a, b in ev:
    a.record(); s._run(args.chains, args.sweeps, None, None, None, None, None, None, out=out, plan=plan); b.record()
torch.cuda.synchronize()
ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
print(json.dumps({"graph": args.graph, "n": g.n, "edges": g.n_edges, "chains": args.chains, "sweeps": args.sweeps,
                  "anneal": args.anneal, "accept": args.accept, "lib": os.path.basename(args.lib), "plan": s.last_plan, "kernel": s.last_kernel, "ms": ms,
                  "spin_updates_per_s": args.chains * args.sweeps * g.n / ms * 1e3}))
