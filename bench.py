#!/usr/bin/env python
"""bench.py -- GRBM spin-updates/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # CPU arm: the oracle port on host cores

Workload (config.workload): BASELINE.json configs[1] -- block-Gibbs sampling on the Pegasus
P16 fabric graph (5 640 spins, 40 484 couplers), 4 096 chains x 1 000 sweeps at beta = 1 per
GPU, synthetic h ~ U(-0.05, 0.05) * prefactor, J ~ U(-5, 5) * prefactor, prefactor 0.05
(SURVEY.md section 8d cfg2).  A *step* is one pass of the hot path over one batch of chains:
sampler.sample_grbm (sweeps + sample energies) followed by the integer edge statistics of the
samples; at N > 1 each rank owns 4 096 chains of the global chain-id space (weak scaling, no
data-path collective) and the step ends with the path's one exchange, an NCCL all-reduce of
the N + E int64 counters.

`value`  : device-resident throughput (h/J and state in HBM; CUDA events, max over ranks).
`e2e`    : the same metric through the reference-facing call sampler.sample_ising(h, J, num_reads, ...)
           with HOST h/J arrays in and HOST samples out (pinned H2D + D2H inside the timed region).
`roofline`: dominant kernel b200grbm::gibbs_wide_kernel (csrc/gibbs_wide.cu; gibbs_kernel for other geometries).  SURVEY.md section 8(d): the sweep is not HBM bound (one
           state read + write per launch); algorithmic bytes are (mean degree + 1) per update against the
           shared-memory roof n_SM x 128 B/clk x f_SM; the HBM view is reported next to it.
`cpu_baseline`: the oracle port's textbook double-precision sequential heat bath on the box's host cores (N = 1 only).

Blocks next to the headline, measured at EVERY N (they are what SCALE_rNN.json carries for BASELINE.json's other
multi-GPU configurations):
`cfg4`       : configs[3] -- Zephyr Z15, 262 144 chains split over the N ranks (STRONG scaling), 100 sweeps, integer
               statistics + the NCCL int64 all-reduce per step; at N > 1 rank 0 also runs the whole job alone, in the
               same process, for `strong_scaling_vs_n1`.
`mmd_sharded`: configs[2] with the 8 192 + 8 192 rows sharded over the ranks: rows exchanged as one bit per spin, pulled
               from the peers' memory over NVLink by the kernel that writes the int8 Gram operand, each rank contracts
               its share of the Gram tiles, one int64 all-reduce of the Hamming histograms; checked equal, count for
               count and bit for bit, to the single-GPU result.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

P16_MEAN_DEGREE = 2 * 40484 / 5640.0
CFG = dict(pegasus_m=16, chains=4096, sweeps=1000, prefactor=0.05, beta=1.0, seed=775321899904)
WORKLOAD = ("GRBM block-Gibbs, Pegasus P16 (5640 spins, 40484 couplers), 4096 chains x 1000 sweeps per GPU, beta=1, "
            "prefactor 0.05 (BASELINE.json configs[1])")


def make_problem(cfg):
    import image_generation_b200 as B

    g = B.IsingGraph.pegasus(cfg["pegasus_m"])
    rng = np.random.default_rng(cfg["seed"] % (2 ** 32))
    h = (cfg["prefactor"] * rng.uniform(-0.05, 0.05, g.n)).astype(np.float32)
    J = (cfg["prefactor"] * rng.uniform(-5.0, 5.0, g.n_edges)).astype(np.float32)
    return g, h, J


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def cpu_port_rate(g, h, J, beta, target_seconds, threads=None):
    """Oracle port (textbook double heat bath, all host threads) on a bounded sample of the workload."""
    from oracle import oracle as O

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is specified as "all the host threads it can use"
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    csr = O.PositionCSR(g.n, g.edge_i, g.edge_j, g.order)
    chains = 16 * cores            # the oracle threads over Philox blocks of 8 chains: two blocks per thread
    st = O.init_state(csr, chains, 1)
    t = time.perf_counter()
    O.gibbs(csr, h, J, st, [beta] * 4, seed=1, f64=True)                       # calibration pass
    rate = chains * 4 * g.n / (time.perf_counter() - t)
    sweeps = max(4, int(target_seconds * rate / (chains * g.n)))
    t = time.perf_counter()
    O.gibbs(csr, h, J, st, [beta] * sweeps, seed=2, f64=True)
    dt = time.perf_counter() - t
    return (chains * sweeps * g.n / dt, cores, f"Pegasus P16, {chains} chains x {sweeps} sweeps ({dt:.1f} s), oracle_gibbs_f64",
            {"chains": chains, "sweeps": sweeps, "seconds": dt})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, h, J = make_problem(CFG)
    rates, sample, cores, last = [], "", 1, {}
    per_step = max(1.0, min(15.0, 120.0 / max(1, args.steps + args.warmup)))
    for k in range(args.warmup + args.steps):
        r, cores, sample, last = cpu_port_rate(g, h, J, CFG["beta"], per_step)
        if k >= args.warmup:
            rates.append(r)
    val = float(np.mean(rates))
    full_updates = CFG["chains"] * CFG["sweeps"] * g.n
    line = {
        "impl": "reference", "metric": "grbm_spin_updates_per_s", "value": val, "unit": "spin-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * full_updates / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        # every step times a BOUNDED sample of the workload; ms_per_step is what the full 4096 x 1000 step would take
        # at the measured rate -- an extrapolation, not a measured full step
        "extrapolated_from": {"sample_chains": last.get("chains"), "sample_sweeps": last.get("sweeps"),
                              "sample_seconds": last.get("seconds"), "sample_updates": last.get("chains", 0) * last.get("sweeps", 0) * g.n,
                              "full_step_updates": full_updates,
                              "note": "ms_per_step = full_step_updates / measured rate; the timed region per step is the sample"},
        "cpu_baseline": {"value": val, "unit": "spin-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "spin-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference ships no CPU sampler source (arithmetic in un-vendored dwave-samplers); this is the oracle port "
                "of the reference-style sequential heat bath on all host threads, each step a bounded sample",
    }
    print(json.dumps(line))


def _load_profile_metrics():
    """ncu-derived constants of the dominant kernel, written by tools/ncu_summary.py --json from the committed capture
    (never literals in this file)."""
    for name in ("r2_gibbs_wide_ncu_metrics.json", "r2_gibbs_ncu_metrics.json", "r1_gibbs_v5_ncu_metrics.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                d = json.load(open(path))
                d["file"] = "profiles/" + name
                return d
            except (OSError, ValueError):
                pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import image_generation_b200 as B
    from image_generation_b200.dist import allreduce_statistics
    from image_generation_b200.stats import sample_statistics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g, h, J = make_problem(CFG)
    chains, sweeps = args.chains or CFG["chains"], args.sweeps or CFG["sweeps"]
    sampler = B.BlockGibbsSampler(g, device=dev, num_sweeps=sweeps, seed=CFG["seed"], accept=args.accept,
                                  chain_offset=rank * chains)
    dg = sampler.device_graph
    h_d, J_d = torch.from_numpy(h).to(dev), torch.from_numpy(J).to(dev)
    sum_s = torch.zeros(g.n, dtype=torch.int64, device=dev)
    sum_ss = torch.zeros(g.n_edges, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches = {"n": 0}

    # caller-owned output buffers: the steady state allocates nothing (a cudaMalloc between the
    # start event and the launch would be charged to the kernel)
    out_bufs = (torch.empty((chains, g.n), dtype=torch.int8, device=dev), torch.empty(chains, dtype=torch.float64, device=dev))

    def sample():
        # the public device-resident call (what GraphRestrictedBoltzmannMachine.sample makes): h, J already in HBM
        return sampler.sample_grbm(h_d, J_d, 1.0, num_reads=chains, num_sweeps=sweeps, out=out_bufs)

    def finish_step(ss):
        sum_s.zero_(); sum_ss.zero_()
        sample_statistics(ss, dg, out=(sum_s, sum_ss))
        if world > 1:
            a, b = allreduce_statistics([sum_s, sum_ss])
            sum_s.copy_(a); sum_ss.copy_(b)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    clocks = ClockSampler(local)
    if not args.no_clocks:
        clocks.start()                              # started before warm-up so the GPU never idles before step 0
    for _ in range(args.warmup):
        finish_step(sample())
    sync_all()
    t_wall0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if not args.no_flush:
            flush.fill_(k & 0xFF)                   # L2 flush between timed steps (256 MB write > 126 MB L2)
        sync_all()
        ev[k][0].record()
        kev[k][0].record()                          # the dominant kernel alone, on the stream it is launched on
        ss = sample()
        kev[k][1].record()
        n_l = 2 + sampler.last_launches             # set_weights (edge + node kernels) + sweeps + sample energies
        finish_step(ss)
        ev[k][1].record()
        launches["n"] += n_l + 2                    # edge + node statistics kernels (on the sampler's packed state)
    sync_all()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1)
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    kern_ms = torch.tensor([a.elapsed_time(b) for a, b in kev], dtype=torch.float64, device=dev)
    total_ms = step_ms.sum().reshape(1)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms) / 1e3
    updates_per_step = chains * sweeps * g.n * world
    value = updates_per_step * args.steps / total_s
    timed_plan = sampler.last_plan
    timed_kernel = {"wide": "b200grbm::gibbs_wide_kernel", "small": "b200grbm::gibbs_small_kernel"}.get(
        sampler.last_kernel, "b200grbm::gibbs_kernel")

    # ---- end to end through the reference-facing call, host buffers in and out
    h_pin, J_pin = torch.from_numpy(h).pin_memory(), torch.from_numpy(J).pin_memory()
    out_pin = torch.empty((chains, g.n), dtype=torch.int8).pin_memory()
    e_pin = torch.empty(chains, dtype=torch.float64).pin_memory()

    def step_e2e():
        ss = sampler.sample_ising(h_pin.numpy(), J_pin.numpy(), num_reads=chains, num_sweeps=sweeps, answer_mode="raw",
                                  auto_scale=False, annealing_time=1, label="bench")
        out_pin.copy_(ss.samples_tensor, non_blocking=True)
        e_pin.copy_(ss.energies_tensor, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_pin

    step_e2e()
    sync_all()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = updates_per_step * e2e_steps / float(e2e_s)

    # ---- BASELINE.json configs[3] and the cross-rank MMD: every rank takes part
    del out_bufs, flush
    torch.cuda.empty_cache()
    cfg4 = mmd_sh = None
    if not args.skip_extra:
        cfg4 = bench_cfg4(dev, rank, world)
        mmd_sh = bench_mmd_sharded(dev, rank, world)

    # every rank leaves the process group here, together; the CPU baseline and the side metrics
    # below are rank-0-only work with no collective in them
    if world > 1:
        from image_generation_b200.dist import release_peer_buffers
        release_peer_buffers()               # the NVLink exchange buffers of mmd_sharded (collective)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    kernel_s = float(kern_ms.mean()) / 1e3            # set_weights + gibbs_kernel + the small energy kernel behind it
    upd_per_launch = chains * sweeps * g.n
    alg_bytes = (P16_MEAN_DEGREE + 1.0) * upd_per_launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    f_sm = (clk.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
    smem_peak = sms * 128 * f_sm / 1e9
    achieved = alg_bytes / kernel_s / 1e9
    hbm_bytes = 2.0 * chains * g.n                      # int8 state written once (+ read when resuming chains)
    prof = _load_profile_metrics()
    roofline = {
        "bound": "smem", "kernel": timed_kernel, "achieved": achieved, "peak": smem_peak, "unit": "GB/s",
        "frac": achieved / smem_peak,
        "traffic": None, "algorithmic_bytes_per_update": P16_MEAN_DEGREE + 1.0, "updates_per_launch": upd_per_launch,
        "kernel_ms": kernel_s * 1e3,
        "peak_source": f"SURVEY.md 8(d): n_SM({sms}) x 128 B/clk x SM clock sampled under load ({f_sm / 1e6:.0f} MHz)",
        "hbm": {"achieved": hbm_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": hbm_bytes,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
        "note": "state is bit-packed (28 chains per word) so the kernel is issue-bound, not byte-bound; see DESIGN.md 5.1",
    }
    if prof:
        # read from the committed ncu capture (tools/ncu_summary.py --json): DRAM traffic of one launch and the warp
        # instructions per 32 spin-updates that set the issue roof
        roofline["traffic"] = prof.get("dram_bytes_per_launch")
        roofline["traffic_source"] = prof.get("file")
        wi = prof.get("warp_instr_per_32_updates")
        if wi:
            roofline["issue"] = {"warp_instr_per_32_updates": wi, "peak_updates_per_s": sms * 4 * 32 / wi * f_sm,
                                 "frac": (upd_per_launch / kernel_s) / (sms * 4 * 32 / wi * f_sm),
                                 "ncu_issue_active_frac": prof.get("issue_active_frac"), "source": prof.get("file")}
    line = {
        "metric": "grbm_spin_updates_per_s", "value": value, "unit": "spin-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "chains_per_gpu": chains, "sweeps": sweeps, "accept": args.accept,
                   "chains_per_lane": timed_plan[0], "threads": timed_plan[1],
                   "l2": "256 MB buffer written between timed steps", "parallelism": f"chains sharded x{world}",
                   "exchange": "int64 all-reduce of N+E counters per step" if world > 1 else "none",
                   "api": "BlockGibbsSampler.sample_grbm (device-resident h, J) + stats.sample_statistics"},
        "e2e": {"value": e2e_val, "unit": "spin-updates/s", "h2d_bytes_per_step": int(4 * (g.n + g.n_edges)),
                "d2h_bytes_per_step": int(chains * g.n + 8 * chains), "steps": e2e_steps},
        "gpu_launches": launches["n"], "clocks": clk, "roofline": roofline,
        "step_ms": [round(float(v), 3) for v in step_ms.tolist()],
    }
    if world == 1:
        cpu_rate, cores, sample_txt, _ = cpu_port_rate(g, h, J, CFG["beta"], args.cpu_seconds)
        line["cpu_baseline"] = {"value": cpu_rate, "unit": "spin-updates/s", "cores": cores, "kind": "port", "sample": sample_txt}
    if cfg4 is not None:
        line["cfg4"] = cfg4
        line["mmd_sharded"] = mmd_sh
    if not args.skip_extra:
        line["sweep_variants"] = bench_sweep_variants(dev, g, h, J)
        line["mmd"] = bench_mmd(dev, peaks)
        line["dvae_step"] = bench_dvae_step(dev)
    print(json.dumps(line))


def bench_cfg4(dev, rank, world, total_chains=262144, sweeps=100, steps=4, warmup=2):
    """BASELINE.json configs[3]: Zephyr Z15 (7 440 spins, 71 736 couplers), 262 144 chains split by global chain id over
    the ranks (strong scaling), 100 sweeps, the integer sufficient statistics of all chains and the NCCL all-reduce of
    the N + E int64 counters -- the h/J gradient exchange -- inside every timed step."""
    import torch
    import torch.distributed as dist

    import image_generation_b200 as B
    from image_generation_b200.dist import allreduce_statistics, shard_chains
    from image_generation_b200.stats import sample_statistics

    z = B.IsingGraph.zephyr(15)
    rng = np.random.default_rng(15)
    hz = torch.from_numpy((CFG["prefactor"] * rng.uniform(-0.05, 0.05, z.n)).astype(np.float32)).to(dev)
    Jz = torch.from_numpy((CFG["prefactor"] * rng.uniform(-5.0, 5.0, z.n_edges)).astype(np.float32)).to(dev)

    def run(off, cnt, reduce_over_ranks, n_steps):
        s = B.BlockGibbsSampler(z, device=dev, num_sweeps=sweeps, seed=CFG["seed"], chain_offset=off)
        out = (torch.empty((cnt, z.n), dtype=torch.int8, device=dev), torch.empty(cnt, dtype=torch.float64, device=dev))
        bufs = (torch.zeros(z.n, dtype=torch.int64, device=dev), torch.zeros(z.n_edges, dtype=torch.int64, device=dev))
        ex_ms = []

        def step(seed, timed):
            ss = s.sample_grbm(hz, Jz, 1.0, num_reads=cnt, num_sweeps=sweeps, seed=seed, out=out)
            bufs[0].zero_(); bufs[1].zero_()
            sample_statistics(ss, s.device_graph, out=bufs)
            if reduce_over_ranks:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                a, b = allreduce_statistics([bufs[0], bufs[1]])
                bufs[0].copy_(a); bufs[1].copy_(b)
                e1.record()
                if timed:
                    ex_ms.append((e0, e1))
        for k in range(warmup):
            step(100 + k, False)
        torch.cuda.synchronize(dev)
        if reduce_over_ranks:
            dist.barrier()
            torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for k, (a, b) in enumerate(ev):
            a.record()
            step(200 + k, True)
            b.record()
        torch.cuda.synchronize(dev)
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        ex = float(np.mean([a.elapsed_time(b) for a, b in ex_ms])) if ex_ms else 0.0
        step(999, False)                                   # identity check: fixed seed, whatever the sharding
        torch.cuda.synchronize(dev)
        stats_sum = int((bufs[1] * torch.arange(1, z.n_edges + 1, device=dev)).sum().item())   # weighted checksum of sum s_i s_j
        return ms, ex, s.last_plan, stats_sum

    off, cnt = shard_chains(total_chains, rank, world)
    ms, ex, plan, checksum = run(off, cnt, world > 1, steps)
    t = torch.tensor([ms, ex], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, ex_max = float(t[0]), float(t[1])
    out = {"workload": f"Zephyr Z15 ({z.n} spins, {z.n_edges} couplers), {total_chains} chains over {world} GPU(s), {sweeps} sweeps, "
                       "integer statistics + int64 all-reduce per step (BASELINE.json configs[3])",
           "scaling": "strong", "chains_total": total_chains, "chains_per_gpu": cnt, "sweeps": sweeps, "steps": steps,
           "ms_per_step": ms_max, "spin_updates_per_s": total_chains * sweeps * z.n / ms_max * 1e3,
           "exchange_ms": ex_max, "exchange_bytes": 8 * (z.n + z.n_edges), "plan": list(plan),
           "statistics_checksum": checksum}
    if world > 1:
        # the whole job on ONE GPU, same process, same seeds: the strong-scaling denominator and an identity check
        # (integer statistics of the same global chains must be equal whatever the sharding)
        if rank == 0:
            ms1, _, _, checksum1 = run(0, total_chains, False, 2)
            out["n1_ms_per_step_same_run"] = ms1
            out["strong_scaling_vs_n1"] = ms1 / ms_max
            out["statistics_equal_to_n1"] = bool(checksum1 == checksum)
            rest = ms_max - ms1 / world
            out["limiter"] = (f"per-GPU sweep launch of {cnt} chains ({ms_max - ex_max:.2f} ms incl. statistics) + exchange {ex_max:.2f} ms; "
                              f"ideal {ms1 / world:.2f} ms, lost {rest:.2f} ms to wave quantisation of "
                              f"{-(-cnt // plan[0])} chain groups over the SMs, the fixed per-launch work and the all-reduce")
        dist.barrier()
    else:
        out["strong_scaling_vs_n1"] = 1.0
    return out


def _mmd_exchange_report(world, m, d_pad, d):
    """What sharded_mmd_loss moved between the ranks: rows as one bit per spin pulled over NVLink by the unpack kernel
    ("p2p"), the same bit rows through an NCCL all-gather ("bits"), or int8 rows through NCCL ("int8")."""
    from image_generation_b200.dist import _DeviceOps
    mode = _DeviceOps.last_exchange
    if world == 1:
        return {"mode": mode, "row_bytes_per_rank_inbound": 0, "allreduce_int64_bytes": 0}
    total = m * d_pad if mode == "int8" else m * d_pad // 8
    return {"mode": mode, "collective_on_rows": "none (peer loads over NVLink, flag-synchronised)" if mode == "p2p" else "nccl all-gather",
            "row_bytes_total": int(total), "row_bytes_per_rank_inbound": int(total * (world - 1) // world),
            "allreduce_int64_bytes": int(3 * (d + 1) * 8)}


def bench_mmd_sharded(dev, rank, world, m_each=8192, d=5640, iters=8):
    """BASELINE.json configs[2] with its rows sharded over the ranks (m_each / N encoder rows and as many samples per
    rank): dist.sharded_mmd_loss = bit-row exchange over NVLink peer memory (csrc/peer_exchange.cu), every rank contracts its share of the Gram tiles into
    Hamming histograms, one int64 all-reduce, float64 evaluation; backward for the rank's own rows."""
    import torch
    import torch.distributed as dist

    import image_generation_b200 as B
    from image_generation_b200.dist import sharded_mmd_loss
    from image_generation_b200.mmd_tc import mmd_block_sums_i8, pack_rows_i8

    gen = torch.Generator(device=dev).manual_seed(11)                  # same seed on every rank: same global matrix
    z = torch.randint(0, 2, (2 * m_each, d), generator=gen, dtype=torch.int8, device=dev) * 2 - 1
    z[m_each:, : d // 8] = 1
    mx = m_each // world
    x_loc = z[rank * mx:(rank + 1) * mx].float().requires_grad_(True)
    y_loc = z[m_each + rank * mx: m_each + (rank + 1) * mx].contiguous()
    kern = B.GaussianKernel(7).to(dev)

    def call():
        x_loc.grad = None
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        val = sharded_mmd_loss(x_loc, y_loc, kern)
        e1.record()
        val.backward()
        e2.record()
        return val, e0, e1, e2

    for _ in range(3):
        call()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    fw, bw = [], []
    for _ in range(iters):
        val, e0, e1, e2 = call()
        torch.cuda.synchronize(dev)
        fw.append(e0.elapsed_time(e1)); bw.append(e1.elapsed_time(e2))
    t = torch.tensor([float(np.mean(fw)), float(np.mean(bw))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # single-GPU truth on this rank: the same global matrix in one piece
    x_all, y_all = z[: mx * world], z[m_each: m_each + mx * world]
    zi, _ = pack_rows_i8(torch.cat([x_all, y_all], 0))
    sums1 = mmd_block_sums_i8(zi, mx * world, kern, d=d)
    m_x = m_y = mx * world
    from image_generation_b200.mmd import _estimate
    val1 = _estimate(sums1, m_x, m_y, kern, "unbiased")[0].to(torch.float32)
    same = torch.tensor([int(torch.equal(val.detach().reshape(()), val1.reshape(())))], device=dev)
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    return {"workload": f"MMD {m_x} x {m_y} rows, D = {d}, 7 kernels, rows sharded over {world} GPU(s) (BASELINE.json configs[2])",
            "forward_ms": float(t[0]), "backward_ms": float(t[1]), "rows_per_gpu": mx,
            "exchange": _mmd_exchange_report(world, m_x + m_y, zi.shape[1], d),
            "value": float(val.detach()), "bit_identical_to_single_gpu_on_every_rank": bool(int(same.item()) == 1)}


def bench_sweep_variants(dev, g, h, J, iters=3):
    """The sweep kernels off the headline configuration: fast (MUFU, 16-bit uniform) acceptance, an annealed schedule on
    the same P16 problem, the per-GPU shard of BASELINE.json configs[3] at 8 GPUs (Zephyr Z15, 32 768 chains, 100 sweeps),
    the reference's own default call (256 reads on the 256-spin Advantage2 sub-graph, configs[0]) and the same graph with a
    GPU's share of configs[4]'s 1 M annealed chains (several chain groups per CTA on resident tables).  Device-resident,
    CUDA events, no L2 flush (state and tables live in shared memory / registers)."""
    import torch

    import image_generation_b200 as B

    def timed(graph, hh, JJ, chains, sweeps, stats=False, **kw):
        s = B.BlockGibbsSampler(graph, device=dev, **kw)
        hd, Jd = torch.from_numpy(hh).to(dev), torch.from_numpy(JJ).to(dev)
        out = (torch.empty((chains, graph.n), dtype=torch.int8, device=dev),
               torch.empty(chains, dtype=torch.float64, device=dev))
        run = lambda n_sweeps=sweeps: s.sample_grbm(hd, Jd, 1.0, num_reads=chains, num_sweeps=n_sweeps, out=out)
        run()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            a.record()
            run()
            b.record()
        torch.cuda.synchronize(dev)
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        res = {"ms": ms, "spin_updates_per_s": chains * sweeps * graph.n / ms * 1e3, "chains": chains,
               "sweeps": sweeps, "plan": list(s.last_plan), "kernel": s.last_kernel}
        if stats:
            # integer statistics (sum s_i, sum s_i s_j) straight from the sampler's packed final state
            from image_generation_b200.stats import sample_statistics
            ss = run(1)
            bufs = (torch.zeros(graph.n, dtype=torch.int64, device=dev),
                    torch.zeros(graph.n_edges, dtype=torch.int64, device=dev))
            sample_statistics(ss, s.device_graph, out=bufs)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                sample_statistics(ss, s.device_graph, out=bufs)
            b.record()
            torch.cuda.synchronize(dev)
            st_ms = a.elapsed_time(b) / 10
            packed_bytes = ss.packed.numel() * 4
            res["statistics"] = {"ms": st_ms, "packed_state_bytes": packed_bytes,
                                 "int8_state_bytes": chains * graph.n,
                                 "packed_GBps": packed_bytes / st_ms / 1e6,
                                 "note": "edge + node statistics kernels over the bit-packed state (N/8 bytes per "
                                         "chain instead of the N bytes of SURVEY 8(d)'s stand-alone int8 form)"}
            # a short launch: the int8 write-back (coalesced through shared memory) is most of its HBM traffic
            short = 5
            run(short)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                run(short)
            b.record()
            torch.cuda.synchronize(dev)
            res["short_launch_5_sweeps_ms"] = a.elapsed_time(b) / 3
        return res

    out = {"p16_fast_acceptance": timed(g, h, J, CFG["chains"], CFG["sweeps"], accept="fast"),
           "p16_annealed_0.1_to_1": timed(g, h, J, CFG["chains"], CFG["sweeps"], beta_range=(0.1, 1.0))}
    z = B.IsingGraph.zephyr(15)
    rng = np.random.default_rng(15)
    hz = (CFG["prefactor"] * rng.uniform(-0.05, 0.05, z.n)).astype(np.float32)
    Jz = (CFG["prefactor"] * rng.uniform(-5.0, 5.0, z.n_edges)).astype(np.float32)
    out["z15_shard_32768_chains"] = timed(z, hz, Jz, 32768, 100, stats=True)
    ck = np.load(os.path.join(ROOT, "tests", "golden", "grbm_checkpoints.npz"))
    name = "Advantage2_system1_10_epochs"
    gc = B.IsingGraph.build(256, ck[name + "/edge_i"], ck[name + "/edge_j"])
    hc = np.clip(np.float32(0.05) * ck[name + "/linear"], -4, 4).astype(np.float32)
    Jc = np.clip(np.float32(0.05) * ck[name + "/quadratic"], -1, 1).astype(np.float32)
    out["cfg1_256_reads_256_spins"] = timed(gc, hc, Jc, 256, 1000)
    # per-GPU share of BASELINE.json configs[4] on 8 GPUs: 1 M annealed chains of the 256-latent model, 100 sweeps per step
    out["cfg5_shard_131072_chains_256_spins_annealed"] = timed(gc, hc, Jc, 131072, 100, beta_range=(0.1, 1.0))
    return out


def bench_mmd(dev, peaks, m_each=8192, d=5640, iters=5):
    """BASELINE.json configs[2]: fused mixture-of-RBF MMD, 8192 encoder latents vs 8192 GRBM samples,
    latent dim = P16 graph size, +-1 rows on tensor cores (e2m1 operands at this size; auto bandwidth from ONE Gram pass)."""
    import torch

    import image_generation_b200 as B
    from image_generation_b200 import _lib as L
    from image_generation_b200 import mmd as M

    # int8 tcgen05 peak of THIS device (MEASURED_PEAKS.json has none): back-to-back kind::i8 MMAs from resident
    # zero operands -- an upper bound real data cannot reach under the power cap; 2 x the cuBLAS bf16 burst beside it
    i8_probe = L.tensor_peak("i8", 10000, dev) / 1e12
    gen = torch.Generator().manual_seed(1)
    z = (torch.randint(0, 2, (2 * m_each, d), generator=gen, dtype=torch.int8) * 2 - 1).to(dev)
    z[m_each:, : d // 8] = 1
    from image_generation_b200.mmd_tc import mmd_block_sums_i8, pack_rows_i8
    zi, _ = pack_rows_i8(z)          # resident in HBM in the kernel's layout (row pitch = whole 128-byte lines)
    out = {}
    for label, bw in (("auto_bandwidth", None), ("fixed_bandwidth", 75.0)):
        kern = B.GaussianKernel(7, bandwidth=bw).to(dev)
        for _ in range(3):
            mmd_block_sums_i8(zi, m_each, kern, d=d)
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            a.record()
            mmd_block_sums_i8(zi, m_each, kern, d=d)
            b.record()
        torch.cuda.synchronize(dev)
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        m = 2 * m_each
        tiles = (m // 256) * (m // 256 + 1)                     # upper triangle of 128 x 256 tiles
        from image_generation_b200.mmd_tc import use_fp4_gram
        fp4 = use_fp4_gram(m)
        k_pad = -(-d // 256) * 256 if fp4 else -(-d // 128) * 128
        executed = 2.0 * tiles * 128 * 256 * k_pad
        i8_2x = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        # what bounds the pass is shared-memory bandwidth: every k-block (48 KB of operands) is written once by TMA and
        # read once by the MMAs, 128 B/clk per SM
        operand_bytes = 2.0 * tiles * (128 + 256) * (k_pad // 2 if fp4 else k_pad)
        smem_peak = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e9
        out[label] = {
            "ms": ms, "gram_passes": 1, "operands": "e2m1 (tcgen05.mma.kind::mxf4, unit block scales)" if fp4 else "int8 (kind::i8)",
            "input_GBps": m * d / ms / 1e6,
            "tflops_as_reference_computes_it": 2.0 * m * m * d / ms / 1e9,
            "roofline": {"bound": "tensor", "achieved": executed / ms / 1e9, "peak": (2.0 if fp4 else 1.0) * i8_probe, "unit": "TOP/s",
                         "frac": executed / ms / 1e9 / ((2.0 if fp4 else 1.0) * i8_probe), "traffic": None,
                         "peak_source": ("2 x " if fp4 else "") + "int8 tcgen05 probe on this device (b200grbm_tensor_peak, resident zero "
                                        "operands)" + ("; e2m1 runs at twice the int8 rate" if fp4 else ""),
                         "frac_of_2x_bf16_burst": executed / ms / 1e9 / i8_2x, "peak_2x_bf16_burst": i8_2x,
                         "executed_ops": executed,
                         "shared_memory": {"achieved": operand_bytes / ms / 1e6, "peak": smem_peak, "unit": "GB/s",
                                           "frac": operand_bytes / ms / 1e6 / smem_peak,
                                           "note": "operand bytes written by TMA + read by the MMAs, against 148 SMs x 128 B/clk; "
                                                   "includes the packing pass and the histogram evaluation in the time"},
                         "note": "symmetric: only the upper triangle of tiles is contracted; the epilogue counts Hamming "
                                 "distances (integer histograms), one pass"}}
    # value + gradient wrt x through the reference's own call, no extra arguments (src/model_wrapper.py:320-326)
    x = z[:m_each].float().requires_grad_(True)
    y = z[m_each:].float()
    kern = B.GaussianKernel(7).to(dev)
    for _ in range(2):
        x.grad = None
        B.maximum_mean_discrepancy_loss(x=x, y=y, kernel=kern).backward()
    torch.cuda.synchronize(dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    x.grad = None
    e0.record()
    val = B.maximum_mean_discrepancy_loss(x=x, y=y, kernel=kern)
    e1.record()
    val.backward()
    e2.record()
    torch.cuda.synchronize(dev)
    out["loss_call"] = {"forward_ms": e0.elapsed_time(e1), "backward_ms": e1.elapsed_time(e2), "dispatched_to": M.last_path,
                        "note": "maximum_mean_discrepancy_loss(x=, y=, kernel=GaussianKernel(7)) from fp32 inputs, default path: fused spin "
                                "extraction (rows + transpose) + spin check + 1 Gram pass (e2m1 operands); backward = int8 Gram coefficient pass "
                                "(2 fixed-point digit planes) + tcgen05 int8 GEMM"}
    out["workload"] = f"MMD {m_each} x {m_each} rows, D = {d}, 7 kernels, int8 +-1 rows (BASELINE.json configs[2])"
    return out


def bench_dvae_step(dev, steps=20, warmup=5):
    """BASELINE.json configs[0] on the GPU path: one DVAE+GRBM training step, B=128, R=8, n_latents=256 on the
    Advantage2 checkpoint graph, 256 reads x 1000 sweeps, stock-PyTorch encoder / decoder."""
    import torch

    from image_generation_b200.dvae import HybridDVAE, synthetic_batch

    z = np.load(os.path.join(ROOT, "tests", "golden", "grbm_checkpoints.npz"))
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))

    def run(**kw):
        model = HybridDVAE(range(256), edges, device=dev, **kw)
        model.setup()
        model.train_init(n_epochs=1, n_batches=steps + warmup)
        batches = [(synthetic_batch(128, seed=k, device=dev), None) for k in range(4)]
        for k in range(warmup):
            model.step(batches[k % 4], epoch=0, record_losses=False)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(steps):
            model.step(batches[k % 4], epoch=0, record_losses=False)
        torch.cuda.synchronize(dev)
        return 1e3 * (time.perf_counter() - t0) / steps

    ms = run()
    out = {"ms_per_step": ms, "steps": steps, "workload": "DVAE+GRBM step, B=128, R=8, n_latents=256, 256 reads x 1000 sweeps, "
           "MMD on tcgen05 int8 path, NLL via packed statistics every 10th step (BASELINE.json configs[0], GPU path)"}
    try:
        out["persistent_chains_20_sweeps_ms_per_step"] = run(persistent=20)
    except Exception as exc:  # a side metric must never take the bench line down
        out["persistent_chains_error"] = repr(exc)
    try:    # the stock-PyTorch nets (forward + backward) replayed as two CUDA graphs: the step stops being launch-bound
        out["graphed_nets_ms_per_step"] = run(graphed=True)
        out["graphed_nets_persistent_chains_20_sweeps_ms_per_step"] = run(graphed=True, persistent=20)
    except Exception as exc:
        out["graphed_nets_error"] = repr(exc)
    try:
        out["cpu_baseline"] = cpu_dvae_step(z, name, edges)
    except Exception as exc:
        out["cpu_baseline"] = {"error": repr(exc)}
    return out


def cpu_dvae_step(z, name, edges):
    """BASELINE.json configs[0] as the reference runs it without a QPU: the same step on the host -- stock PyTorch
    encoder / decoder on CPU, the oracle port as the classical sampler (256 reads x 1000 sweeps, all host threads),
    the stock torch form of the MMD (cat -> cdist -> 7 x exp -> block means) with autograd."""
    import torch

    from image_generation_b200.dvae import Decoder, DiscreteVariationalAutoencoder, Encoder, synthetic_batch
    from oracle import oracle as O

    O.set_num_threads(len(os.sched_getaffinity(0)))
    torch.manual_seed(0)
    dvae = DiscreteVariationalAutoencoder(Encoder(256), Decoder(256))
    opt = torch.optim.Adam(dvae.parameters(), lr=1e-4, weight_decay=0.01)
    ei, ej = z[name + "/edge_i"], z[name + "/edge_j"]
    h = np.clip(np.float32(0.05) * z[name + "/linear"], -4, 4).astype(np.float32)
    J = np.clip(np.float32(0.05) * z[name + "/quadratic"], -1, 1).astype(np.float32)
    csr = O.PositionCSR(256, ei, ej, np.arange(256))
    images = synthetic_batch(128, seed=0)
    mult = 2.0 ** (torch.arange(7) - 3).float()
    times = []
    for it in range(2):
        t0 = time.perf_counter()
        _, spins, recon = dvae(images, 8)
        opt.zero_grad()
        mse = torch.nn.functional.mse_loss(recon, images.unsqueeze(1).expand(-1, 8, -1, -1, -1))
        samples = torch.from_numpy(O.gibbs(csr, h, J, O.init_state(csr, 256, it), [1.0] * 1000, seed=it, f64=True)).float()
        x = spins.reshape(-1, 256)
        zz = torch.cat([x, samples], 0)
        dmat = torch.cdist(zz, zz, p=2)
        bw = dmat.detach().sum() / (zz.shape[0] ** 2 - zz.shape[0])
        k = torch.exp(-dmat.unsqueeze(0) / (bw * mult).reshape(-1, 1, 1)).sum(0)
        mx, my = x.shape[0], samples.shape[0]
        kxx, kyy, kxy = k[:mx, :mx], k[mx:, mx:], k[:mx, mx:]
        mmd = ((kxx.sum() - kxx.trace()) / (mx * (mx - 1)) + (kyy.sum() - kyy.trace()) / (my * (my - 1)) - 2 * kxy.mean())
        (mse + mmd).backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return {"ms_per_step": 1e3 * times[-1], "cores": O.num_threads(), "kind": "port",
            "sample": "one full step (second of two), oracle_gibbs_f64 sampler, stock-torch CPU nets and MMD; NLL update not included"}


def main():
    if os.environ.get("B200_BENCH_FAULT_AFTER"):      # debugging aid: dump all stacks if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["B200_BENCH_FAULT_AFTER"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--accept", default="exact", choices=["exact", "fast"])
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample nvidia-smi during the timed region")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: skip the L2 flush between timed steps")
    ap.add_argument("--skip-extra", action="store_true", help="headline only: skip cfg4 / mmd_sharded / mmd / dvae_step / sweep_variants")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
