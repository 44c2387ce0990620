"""Host-side graph logic: generators, colourings, ELL tables, launch planning."""
import numpy as np
import pytest

import image_generation_b200 as B
from image_generation_b200.topology import greedy_colouring


def test_pegasus_p16_counts_and_colouring():
    n, ei, ej, col = B.pegasus_graph(16)
    assert (n, ei.size) == (5640, 40484)          # BASELINE.json cfg2
    deg = np.bincount(np.concatenate([ei, ej]), minlength=n)
    assert (deg.min(), deg.max()) == (6, 15) and abs(deg.mean() - 14.356) < 1e-3
    assert not np.any(col[ei] == col[ej])
    assert np.bincount(col).tolist() == [1410] * 4


def test_zephyr_z15_counts_and_colouring():
    n, ei, ej, col = B.zephyr_graph(15)
    assert (n, ei.size) == (7440, 71736)          # BASELINE.json cfg4
    deg = np.bincount(np.concatenate([ei, ej]), minlength=n)
    assert (deg.min(), deg.max()) == (10, 20)
    assert not np.any(col[ei] == col[ej])
    assert np.bincount(col).tolist() == [1860] * 4


def test_rounds_are_balanced_by_merging_colour_remainders():
    """Pegasus P16: four colours of 1410 = 2 x 640 + 130 spins.  The visit order keeps the full chunks and merges
    the four remainders into one independent class: 9 rounds of a 640-lane CTA instead of 8 of a 736-lane one
    (45 instead of 48 warp slots per scheduler and sweep)."""
    from image_generation_b200.topology import balance_rounds, round_cost
    n, ei, ej, col = B.pegasus_graph(16)
    g = B.IsingGraph.pegasus(16)
    sizes = np.diff(g.colour_start).tolist()
    assert sizes == [640] * 8 + [520]
    assert not np.any(g.colour[ei] == g.colour[ej])                         # still a proper colouring
    assert B.sampler.plan_threads(sizes, g.n, g.ell_width) == 640
    assert round_cost(sizes, 640) < round_cost([1410] * 4, 736)
    plain = B.IsingGraph.build(n, ei, ej, colour=col, balance=False)
    assert np.diff(plain.colour_start).tolist() == [1410] * 4 and B.sampler.plan_threads([1410] * 4, n, 15) == 736
    # Zephyr Z15 (four colours of 1860): merging would give 15 rounds of 512 lanes (60 slots, 4 warps per scheduler),
    # but the colouring as it is reaches 60 slots with two 384-lane CTAs per SM (6 warps): kept.  Small graphs: nothing to gain
    nz, zi, zj, zc = B.zephyr_graph(15)
    assert np.array_equal(balance_rounds(nz, zi, zj, zc, 20), zc)
    assert np.diff(B.IsingGraph.pegasus(4).colour_start).tolist() == [66] * 4
    assert np.array_equal(g.colour, B.IsingGraph.pegasus(16).colour)        # deterministic


def test_round_balancing_on_a_generic_four_colourable_lattice():
    """King's graph 75 x 75 (8 neighbours, colours (x & 1, y & 1) of 1444 / 1406 / 1406 / 1369 spins): the same
    arithmetic as Pegasus P16 -- every colour is two chunks of 640 plus a remainder, and the remainders can be
    chosen mutually non-adjacent -- on a graph the generators know nothing about."""
    from image_generation_b200.topology import balance_rounds
    L = 75
    idx = lambda x, y: x * L + y
    ei, ej = [], []
    for x in range(L):
        for y in range(L):
            for dx, dy in ((1, 0), (0, 1), (1, 1), (1, -1)):
                if 0 <= x + dx < L and 0 <= y + dy < L:
                    ei.append(idx(x, y))
                    ej.append(idx(x + dx, y + dy))
    ei, ej = np.array(ei, dtype=np.int32), np.array(ej, dtype=np.int32)
    col = np.array([2 * (x & 1) + (y & 1) for x in range(L) for y in range(L)], dtype=np.int32)
    new = balance_rounds(L * L, ei, ej, col, 8)
    sizes = np.bincount(new).tolist()
    assert not np.any(new[ei] == new[ej]) and new.min() == 0               # proper, every spin assigned
    assert sizes == [640] * 8 + [505]
    # a class is either a chunk of one colour or the merged remainders of all four
    for c in range(8):
        assert len(set(col[new == c].tolist())) == 1
    assert sorted(np.bincount(col[new == 8]).tolist()) == sorted([1444 - 1280, 1406 - 1280, 1406 - 1280, 1369 - 1280])
    g = B.IsingGraph.build(L * L, ei, ej, colour=col)
    assert np.diff(g.colour_start).tolist() == sizes and B.sampler.plan_threads(sizes, g.n, g.ell_width) == 640
    # the graph the sampler sees is the same graph: edges, degrees and a bijective visit order
    assert sorted(g.order.tolist()) == list(range(L * L)) and g.n_edges == ei.size


def test_ell_tables_are_consistent():
    g = B.IsingGraph.pegasus(3)
    assert g.n_pad % 32 == 0 and g.ell_width == g.degree.max()
    for e in range(g.n_edges):
        ka, pa = divmod(int(g.slot_a[e]), g.n_pad)
        kb, pb = divmod(int(g.slot_b[e]), g.n_pad)
        assert pa == g.pos[g.edge_i[e]] and g.ell_nbr[ka, pa] == g.pos[g.edge_j[e]]
        assert pb == g.pos[g.edge_j[e]] and g.ell_nbr[kb, pb] == g.pos[g.edge_i[e]]
    for p in range(g.n):
        row = g.ell_nbr[: g.degree[p], p]
        assert np.all(np.diff(row) > 0)                       # contract order: ascending position
        assert np.all(g.ell_nbr[g.degree[p]:, p] == p)        # padding points at self
    # colour blocks are contiguous and proper
    c_of_pos = g.colour[g.order]
    assert np.all(np.diff(c_of_pos) >= 0)
    assert g.colour_start.tolist() == [0] + np.cumsum(np.bincount(g.colour)).tolist()


def test_checkpoint_graphs_colour_with_few_colours(golden):
    z, meta = golden
    for name in meta:
        ei, ej = z[name + "/edge_i"], z[name + "/edge_j"]
        col = greedy_colouring(256, ei, ej)
        assert not np.any(col[ei] == col[ej])
        counts = np.bincount(col)
        assert len(counts) <= 5                    # 4-5 (iterated greedy; plain greedy gives 6
        assert counts.min() >= 30                  # on Advantage2, with a 4-spin class that costs a whole round)
        assert np.array_equal(col, greedy_colouring(256, ei, ej))      # deterministic
        assert greedy_colouring(256, ei, ej, refine=0).max() + 1 >= len(counts)
        g = B.IsingGraph.build(256, ei, ej)
        assert g.n_colours <= 5 and g.ell_width == np.bincount(np.concatenate([ei, ej])).max()


def test_build_rejects_bad_graphs():
    with pytest.raises(ValueError):
        B.IsingGraph.build(3, [0, 1], [1, 1])                 # self loop
    with pytest.raises(ValueError):
        B.IsingGraph.build(3, [0, 1], [1, 0])                 # duplicate edge
    with pytest.raises(ValueError):
        B.IsingGraph.build(3, [0], [5])                       # out of range
    with pytest.raises(ValueError):
        B.IsingGraph.build(3, [0], [1], colour=[0, 0, 1])     # improper colouring


def test_plan_launch():
    # cfg2: 4096 chains on 148 SMs -> 28 chains per CTA fills 147 SMs in one wave
    assert B.plan_launch(4096, [1410] * 4, 148, 5640, 15)[0] == 28
    assert B.plan_launch(4096, [1410] * 4, 148, 5640, 15) == (28, 736)
    assert B.plan_launch(4096, [640] * 8 + [520], 148, 5640, 15) == (28, 640)   # P16 as IsingGraph.pegasus(16) orders it
    # many groups: two narrow single-stage CTAs per SM when that also needs fewer warp slots per sweep
    assert B.plan_launch(32768, [1860] * 4, 148, 7440, 20) == (28, 384)       # Z15: 20 rounds x 3 < 16 rounds x 4
    assert B.plan_launch(4096, [1860] * 4, 148, 7440, 20) == (28, 480)        # too few groups for two CTAs per SM
    assert B.plan_launch(262144, [1410] * 4, 148, 5640, 15) == (28, 736)      # P16: 48 slots either way -> wide CTA
    for chains in (1, 5, 100, 4096, 262144):
        cpl, threads = B.plan_launch(chains, [7, 9], 148)
        assert cpl in (4, 8, 16, 24, 28, 32) and 64 <= threads <= 768


def test_planner_invariants_over_random_graph_shapes():
    """Whatever the class sizes, chain count and degree: the plan is launchable (CTA size a multiple of 32 in
    [64, 768], two tile stages -- or, for the narrow two-CTA plans, one -- fit in shared memory, chains per lane
    supported) and never costs more than the widest feasible CTA."""
    from image_generation_b200.sampler import SMEM_LIMIT, SMEM_PER_SM, plan_threads, sweep_smem_bytes
    from image_generation_b200.topology import round_cost
    rng = np.random.default_rng(42)
    for _ in range(300):
        k = int(rng.integers(1, 12))
        sizes = [int(x) for x in rng.integers(1, 2500, size=k)]
        n, width = sum(sizes), int(rng.integers(1, 24))
        if n * 4 > 100 * 1024:
            continue
        chains = int(rng.choice([1, 7, 256, 4096, 50000, 262144]))
        cpl, t = B.plan_launch(chains, sizes, 148, n, width)
        assert cpl in (4, 8, 16, 24, 28, 32) and t % 32 == 0 and 64 <= t <= 768
        rounds = sum(-(-s // t) for s in sizes)
        two_stage = sweep_smem_bytes(n, width, t, rounds)
        if two_stage > SMEM_LIMIT:                      # only the single-stage two-CTA plan may exceed it
            assert t <= 384 and 2 * (two_stage - (width + 1) * t * 8 + 1024) <= SMEM_PER_SM
        t1 = plan_threads(sizes, n, width)
        assert sweep_smem_bytes(n, width, t1, sum(-(-s // t1) for s in sizes)) <= SMEM_LIMIT
        feasible = [c for c in range(64, 769, 32) if sweep_smem_bytes(n, width, c, sum(-(-s // c) for s in sizes)) <= SMEM_LIMIT]
        assert round_cost(sizes, t1) <= min(round_cost(sizes, c) for c in feasible) + 1e-9


def test_beta_schedule():
    assert np.all(B.beta_schedule(5) == 1.0)
    s = B.beta_schedule(4, (0.1, 1.0))
    assert s[0] == pytest.approx(0.1) and s[-1] == pytest.approx(1.0) and np.all(np.diff(np.log(s)) > 0)
    with pytest.raises(ValueError):
        B.beta_schedule(4, (0.0, 1.0), "geometric")
    with pytest.raises(ValueError):
        B.beta_schedule(4, (0.1, 1.0), "cubic")
    r = B.beta_schedule(6, (0.5, 2.0), "linear", num_sweeps_per_beta=2)
    assert r.tolist() == [0.5, 0.5, 1.25, 1.25, 2.0, 2.0]
    with pytest.raises(ValueError):
        B.beta_schedule(5, (0.5, 2.0), "linear", num_sweeps_per_beta=2)


def test_greedy_get_subgraph_is_deterministic_and_connected():
    n, ei, ej, _ = B.pegasus_graph(4)
    adj = {v: [] for v in range(n)}
    for a, b in zip(ei.tolist(), ej.tolist()):
        adj[a].append(b)
        adj[b].append(a)
    a = B.greedy_get_subgraph(40, 775321899904, range(n), adj)
    b = B.greedy_get_subgraph(40, 775321899904, range(n), adj)
    assert a == b and len(set(a)) == 40
    chosen = set(a)
    seen, stack = {a[0]}, [a[0]]
    while stack:
        v = stack.pop()
        for u in adj[v]:
            if u in chosen and u not in seen:
                seen.add(u)
                stack.append(u)
    assert seen == chosen
    mapping = B.get_graph_mapping(a)
    assert sorted(mapping.values()) == list(range(40))


def test_plan_launch_small_graphs_with_resident_tables():
    """Small graphs (tables resident, several CTAs per SM): chains per lane follow the occupancy model, reproducing the
    choices measured on a B200 for the 256-spin Advantage2 sub-graph (tools/bench_configs.py --graph cfg1 --cpl ...)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grbm_checkpoints.npz"))
    name = "Advantage2_system1_10_epochs"
    g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
    sizes = np.diff(g.colour_start).tolist()
    want = {256: 4, 1024: 4, 2048: 4, 4096: 8, 8192: 16, 16384: 28, 20000: 28, 33333: 28, 131072: 28}
    for chains, cpl in want.items():
        assert B.plan_launch(chains, sizes, 148, g.n, g.ell_width) == (cpl, 64), chains
    # the big fabric graphs stream their tables: the wave rule is untouched
    assert B.plan_launch(4096, [640] * 8 + [520], 148, 5640, 15) == (28, 640)
