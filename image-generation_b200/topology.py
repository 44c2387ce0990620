"""Qubit-graph construction for the GRBM sampler: Pegasus / Zephyr generators, colourings,
the reference's greedy sub-graph selection, and the colour-major visit order the
sm_100a sweep kernel runs in.

Reference sites this module stands in for:
  * ``DWaveSampler(...).to_networkx_graph()`` at src/utils/common.py:120-121 -- the QPU's
    working graph.  Without a QPU (or dwave_networkx) the ideal-yield Pegasus ``P_m`` /
    Zephyr ``Z_m`` fabric graphs are generated here from their coordinate rules
    (node/edge counts are asserted in tests/test_topology.py).
  * ``greedy_get_subgraph`` (src/utils/common.py:22-84) and ``get_graph_mapping``
    (src/utils/common.py:86-100) -- restated in :func:`greedy_get_subgraph` /
    :func:`get_graph_mapping` with the same ``random.Random(seed)`` call sequence.

Everything here is host-side numpy; nothing touches CUDA at import time (the reference
runs inside a spawned Dash worker, app.py:37-43).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

__all__ = [
    "pegasus_graph",
    "zephyr_graph",
    "greedy_colouring",
    "IsingGraph",
    "greedy_get_subgraph",
    "greedy_get_subgraph_nx",
    "get_graph_mapping",
]

_PEGASUS_O0 = (2, 2, 2, 2, 10, 10, 10, 10, 6, 6, 6, 6)
_PEGASUS_O1 = (6, 6, 6, 6, 2, 2, 2, 2, 10, 10, 10, 10)


def _largest_component(n: int, ei: np.ndarray, ej: np.ndarray) -> np.ndarray:
    """Boolean mask of the nodes in the largest connected component (union-find)."""
    parent = np.arange(n)

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in zip(ei.tolist(), ej.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
    roots = np.array([find(a) for a in range(n)])
    vals, counts = np.unique(roots, return_counts=True)
    return roots == vals[np.argmax(counts)]


def _finish(n_all: int, edges: set, colour_all: np.ndarray, fabric_only: bool):
    e = np.array(sorted(edges), dtype=np.int64).reshape(-1, 2)
    ei, ej = e[:, 0], e[:, 1]
    if fabric_only:
        keep = _largest_component(n_all, ei, ej)
    else:
        keep = np.ones(n_all, dtype=bool)
    relabel = -np.ones(n_all, dtype=np.int64)
    relabel[keep] = np.arange(int(keep.sum()))
    m = keep[ei] & keep[ej]
    ei, ej = relabel[ei[m]], relabel[ej[m]]
    lo, hi = np.minimum(ei, ej), np.maximum(ei, ej)
    return int(keep.sum()), lo.astype(np.int32), hi.astype(np.int32), colour_all[keep].astype(np.int32)


def pegasus_graph(m: int, fabric_only: bool = True):
    """Ideal Pegasus ``P_m``.  Returns ``(n, edge_i, edge_j, colour)`` with ``edge_i < edge_j``
    and a proper 4-colouring ``colour = 2u + ((k + z) & 1)``.  P16: n = 5640, E = 40484."""
    def lab(u, w, k, z):
        return ((u * m + w) * 12 + k) * (m - 1) + z

    n_all = 2 * m * 12 * (m - 1)
    colour = np.zeros(n_all, dtype=np.int32)
    edges = set()
    for u in range(2):
        for w in range(m):
            for k in range(12):
                for z in range(m - 1):
                    a = lab(u, w, k, z)
                    colour[a] = 2 * u + ((k + z) & 1)
                    if z + 1 < m - 1:                       # external
                        edges.add((a, lab(u, w, k, z + 1)))
                    if k % 2 == 0:                          # odd
                        edges.add((a, lab(u, w, k + 1, z)))
    for w in range(m):                                      # internal
        for k in range(12):
            for z in range(m - 1):
                a = lab(0, w, k, z)
                for kk in range(12):
                    w1 = z + (1 if kk < _PEGASUS_O0[k] else 0)
                    z1 = w - (1 if k < _PEGASUS_O1[kk] else 0)
                    if 0 <= w1 < m and 0 <= z1 < m - 1:
                        b = lab(1, w1, kk, z1)
                        edges.add((min(a, b), max(a, b)))
    return _finish(n_all, edges, colour, fabric_only)


def zephyr_graph(m: int, t: int = 4, fabric_only: bool = True):
    """Ideal Zephyr ``Z_m`` (tile parameter t).  Returns ``(n, edge_i, edge_j, colour)``;
    proper 4-colouring ``colour = j + ((w + 2(z+u) + j) & 2)``.  Z15: n = 7440, E = 71736."""
    W = 2 * m + 1

    def lab(u, w, k, j, z):
        return (((u * W + w) * t + k) * 2 + j) * m + z

    n_all = 2 * W * t * 2 * m
    colour = np.zeros(n_all, dtype=np.int32)
    edges = set()
    for u in range(2):
        for w in range(W):
            for k in range(t):
                for j in range(2):
                    for z in range(m):
                        a = lab(u, w, k, j, z)
                        colour[a] = j + ((w + 2 * (z + u) + j) & 2)
                        if z + 1 < m:                       # external
                            edges.add((a, lab(u, w, k, j, z + 1)))
                        if j == 0:                          # odd
                            edges.add((a, lab(u, w, k, 1, z)))
                        elif z + 1 < m:
                            edges.add((a, lab(u, w, k, 0, z + 1)))
    for w in range(W):                                      # internal
        for k in range(t):
            for j in range(2):
                for z in range(m):
                    a = lab(0, w, k, j, z)
                    for w1 in (2 * z + j, 2 * z + j + 1):
                        if not 0 <= w1 < W:
                            continue
                        for k1 in range(t):
                            for j1 in range(2):
                                for z1 in range(m):
                                    if w in (2 * z1 + j1, 2 * z1 + j1 + 1):
                                        b = lab(1, w1, k1, j1, z1)
                                        edges.add((min(a, b), max(a, b)))
    return _finish(n_all, edges, colour, fabric_only)


def _adjacency(n: int, ei: np.ndarray, ej: np.ndarray):
    adj = [[] for _ in range(n)]
    for a, b in zip(ei.tolist(), ej.tolist()):
        adj[a].append(b)
        adj[b].append(a)
    return adj


def _greedy_in_order(adj, order, n: int) -> np.ndarray:
    colour = -np.ones(n, dtype=np.int32)
    for v in order:
        used = {int(colour[u]) for u in adj[v] if colour[u] >= 0}
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    return colour


def _smallest_last_order(adj, n: int) -> list:
    deg = [len(a) for a in adj]
    removed = [False] * n
    buckets: dict[int, set] = {}
    for v, d in enumerate(deg):
        buckets.setdefault(d, set()).add(v)
    order = []
    for _ in range(n):
        d = min(k for k, s in buckets.items() if s)
        v = min(buckets[d])            # deterministic choice inside a bucket
        buckets[d].discard(v)
        removed[v] = True
        order.append(v)
        for u in adj[v]:
            if not removed[u]:
                buckets[deg[u]].discard(u)
                deg[u] -= 1
                buckets.setdefault(deg[u], set()).add(u)
    order.reverse()
    return order


def greedy_colouring(n: int, ei: np.ndarray, ej: np.ndarray, refine: Optional[int] = None) -> np.ndarray:
    """Proper colouring for graphs without a closed form: smallest-last greedy, then *iterated greedy*
    (Culberson): re-colour greedily with the vertices grouped by their current colour class, classes taken in
    a new order -- the number of colours never grows and usually shrinks.  Every colour is one barrier-separated
    round of the sweep kernel, so on the reference's 256-spin QPU sub-graphs a colour less is ~1/5 less latency
    per sweep (the Advantage2 sub-graph goes from 6 classes, one of them 4 spins, to 4-5).  Deterministic
    (fixed-seed permutations); ``refine`` = number of re-colouring passes (default scales down with size)."""
    adj = _adjacency(n, ei, ej)
    best = _greedy_in_order(adj, _smallest_last_order(adj, n), n)
    if n == 0:
        return best
    if refine is None:
        refine = int(min(200, max(0, 400_000 // (n + 2 * len(ei) + 1))))
    rng = np.random.default_rng(0x6C6F7572)

    def score(col):
        counts = np.bincount(col)
        return (len(counts), int(counts.max()) - int(counts.min()))

    cur = best
    for it in range(refine):
        k = int(cur.max()) + 1
        if k <= 2:
            break
        counts = np.bincount(cur, minlength=k)
        if it % 3 == 0:
            classes = list(np.argsort(-counts, kind="stable"))        # largest class first
        elif it % 3 == 1:
            classes = list(range(k - 1, -1, -1))                      # reverse
        else:
            classes = list(rng.permutation(k))
        order = [v for c in classes for v in np.flatnonzero(cur == c).tolist()]
        cur = _greedy_in_order(adj, order, n)
        if score(cur) < score(best):
            best = cur
    return best


#: relative cost of a colour round with w warps per scheduler (ceil(threads / 128)): fewer warps hide less latency.
#: Measured on B200, Pegasus P16, 48 warp slots per sweep in every case: 736 threads (6 warps) 34.6 ms, 480 (4) 36.4,
#: 384 (3) 39.4; 640 threads (5 warps, 45 slots) 32.7 ms.
WARP_PENALTY = {1: 1.6, 2: 1.3, 3: 1.14, 4: 1.05, 5: 1.02, 6: 1.0}
_SMEM_LIMIT = 227 * 1024


def round_cost(class_sizes: Sequence[int], threads: int, ctas: int = 1) -> float:
    """Cost of one sweep for a CTA of ``threads`` lanes: rounds x warps per scheduler x latency-hiding penalty.
    Every independent class of the visit order is split into rounds of at most ``threads`` spins, and each round
    costs every scheduler ``ceil(threads / 128)`` warp slots whatever the number of active lanes.  ``ctas`` CTAs
    resident per SM multiply the warps that hide each other's latency, not the slots per chain."""
    w = -(-threads // 128)
    return sum(-(-s // threads) for s in class_sizes if s > 0) * w * WARP_PENALTY.get(min(6, w * ctas), 1.0)


def _sweep_smem(n: int, width: int, threads: int, n_tiles: int) -> int:
    # mirrors sampler.sweep_smem_bytes / b200grbm_sweep_smem_bytes (2 tile stages)
    return 128 + (n_tiles * 8 + 127) // 128 * 128 + (n * 4 + 127) // 128 * 128 + 2 * (width + 1) * threads * 8


def balance_rounds(n: int, ei: np.ndarray, ej: np.ndarray, colour: np.ndarray, width: int) -> np.ndarray:
    """Refine a proper colouring into independent classes that fill the sweep kernel's rounds better.

    A colour of s spins costs ``ceil(s / T)`` rounds of a T-lane CTA, the last one mostly empty: Pegasus P16 has four
    colours of 1410 = 2 x 640 + 130 spins.  Here every colour keeps its full chunks of T and the *remainders of
    different colours are merged* into common classes -- chosen greedily, staggered over the index range, so that no
    two remainder spins are adjacent.  P16: 8 classes of 640 + one of 520 = 9 rounds x 5 warps per scheduler = 45 warp
    slots per sweep instead of 8 x 6 = 48 (measured 34.0 -> 32.7 ms for 4096 chains x 1000 sweeps).  Returns
    ``colour`` itself when no T gains at least 2 % or the remainders cannot be made independent.  Any refinement of a
    proper colouring is a proper colouring, so the parallel round still equals the sequential sweep in visit order.
    """
    colour = np.asarray(colour, dtype=np.int32)
    if n == 0 or ei.size == 0:
        return colour
    sizes = np.bincount(colour).tolist()
    k = len(sizes)
    cands = [t for t in range(64, 769, 32) if _sweep_smem(n, width, t, sum(-(-s // t) for s in sizes)) <= _SMEM_LIMIT]
    if not cands or k < 2:
        return colour
    plain = min(round_cost(sizes, t) for t in cands)
    # the launcher can also run two narrow single-stage CTAs per SM (many chain groups): twice the warps per
    # scheduler at the same slot count.  If the colouring as it is does at least as well that way, keep it.
    for t in range(64, 385, 32):
        single = _sweep_smem(n, width, t, sum(-(-s // t) for s in sizes)) - (width + 1) * t * 8
        if 2 * (single + 1024) <= 228 * 1024:
            plain = min(plain, round_cost(sizes, t, ctas=2))
    best = None
    for t in cands:
        if min(sizes) < t:
            continue                    # every colour must keep at least one full chunk
        rem = sum(s % t for s in sizes)
        merged = [t] * sum(s // t for s in sizes) + [t] * (rem // t) + ([rem % t] if rem % t else [])
        c = round_cost(merged, t)
        if best is None or c < best[0] - 1e-9:
            best = (c, t)
    if best is None or best[0] > 0.98 * plain:
        return colour
    t = best[1]
    adj = _adjacency(n, ei, ej)
    # spatial order: breadth-first rank from node 0 (restarted per component).  Remainders taken from segments of
    # this order that lie far apart are compact clusters -- their neighbourhoods overlap, so few spins of the other
    # colours get blocked -- and clusters of different colours do not touch.
    rank = -np.ones(n, dtype=np.int64)
    r_next = 0
    for root in range(n):
        if rank[root] >= 0:
            continue
        rank[root] = r_next
        r_next += 1
        frontier = [root]
        while frontier:
            nxt_frontier = []
            for v in frontier:
                for u in adj[v]:
                    if rank[u] < 0:
                        rank[u] = r_next
                        r_next += 1
                        nxt_frontier.append(u)
            frontier = nxt_frontier
    blocked = np.zeros(n, dtype=bool)
    new = -np.ones(n, dtype=np.int32)
    nxt, leftovers = 0, []
    for ci in range(k):
        idx = np.flatnonzero(colour == ci)
        r = idx.size % t
        by_rank = idx[np.argsort(rank[idx], kind="stable")]
        start = (idx.size * ci) // k                      # stagger the remainders of different colours
        rot = np.concatenate([by_rank[start:], by_rank[:start]])
        chosen = [int(v) for v in rot if not blocked[v]][:r]
        if len(chosen) < r:
            return colour                                 # not enough mutually independent remainder spins
        for v in chosen:
            blocked[adj[v]] = True
        leftovers += chosen
        taken = np.zeros(n, dtype=bool)
        taken[chosen] = True
        rest = idx[~taken[idx]]
        for j in range(0, rest.size, t):
            new[rest[j:j + t]] = nxt
            nxt += 1
    for j in range(0, len(leftovers), t):
        new[np.asarray(leftovers[j:j + t], dtype=np.int64)] = nxt
        nxt += 1
    if np.any(new < 0) or np.any(new[ei] == new[ej]):     # cannot happen; keep the proven colouring if it does
        return colour
    return new


@dataclass
class IsingGraph:
    """An Ising graph in the layout the sweep kernel consumes.

    ``order[p]`` is the node visited p-th in a sweep (colour-major: all of colour 0, then
    colour 1, ...).  Because same-colour spins are never adjacent, updating a whole colour
    block in parallel is *identical* to the sequential index-order sweep of the reference-
    style annealer run on the graph relabelled by ``pos``.

    ELL tables are in visit-position space: row p, slot k -> neighbour position
    ``ell_nbr[k, p]`` (ascending in k, padded with p itself) -- the order the field sum of
    include/b200grbm_spec.h is taken in.  ``slot_a/slot_b[e]`` are the two flat ELL slots
    that carry edge e's coupling.
    """

    n: int
    edge_i: np.ndarray
    edge_j: np.ndarray
    colour: np.ndarray
    order: np.ndarray
    pos: np.ndarray
    colour_start: np.ndarray
    ell_width: int
    n_pad: int
    ell_nbr: np.ndarray
    ell_edge: np.ndarray
    slot_a: np.ndarray
    slot_b: np.ndarray
    degree: np.ndarray

    @property
    def n_edges(self) -> int:
        return int(self.edge_i.shape[0])

    @property
    def n_colours(self) -> int:
        return int(self.colour_start.shape[0] - 1)

    @classmethod
    def build(cls, n: int, edge_i: Sequence[int], edge_j: Sequence[int],
              colour: Optional[Sequence[int]] = None, balance: bool = True) -> "IsingGraph":
        ei = np.asarray(edge_i, dtype=np.int32).reshape(-1)
        ej = np.asarray(edge_j, dtype=np.int32).reshape(-1)
        if ei.shape != ej.shape:
            raise ValueError("edge_i and edge_j must have the same length")
        if ei.size and (min(ei.min(), ej.min()) < 0 or max(ei.max(), ej.max()) >= n):
            raise ValueError("edge endpoint out of range")
        if np.any(ei == ej):
            raise ValueError("self-loops are not allowed in an Ising graph")
        key = np.minimum(ei, ej).astype(np.int64) * n + np.maximum(ei, ej)
        if np.unique(key).size != key.size:
            raise ValueError("duplicate edges")
        if colour is None:
            colour = greedy_colouring(n, ei, ej)
        colour = np.asarray(colour, dtype=np.int32)
        if ei.size and np.any(colour[ei] == colour[ej]):
            raise ValueError("colouring is not proper")
        if balance and n and ei.size:
            # refine into classes that fill the kernel's rounds (see balance_rounds); `colour` is from here on the
            # class of the visit order, still a proper colouring
            width0 = int(np.bincount(np.concatenate([ei, ej]), minlength=n).max())
            colour = balance_rounds(n, ei, ej, colour, width0)
        n_colours = int(colour.max()) + 1 if n else 0
        # colour-major visit order; inside a colour keep node order (stable)
        order = np.argsort(colour, kind="stable").astype(np.int32)
        pos = np.empty(n, dtype=np.int32)
        pos[order] = np.arange(n, dtype=np.int32)
        colour_start = np.zeros(n_colours + 1, dtype=np.int32)
        np.cumsum(np.bincount(colour, minlength=n_colours), out=colour_start[1:])
        # directed entries in position space, sorted by (row, neighbour position)
        pa, pb = pos[ei], pos[ej]
        rows = np.concatenate([pa, pb]).astype(np.int64)
        cols = np.concatenate([pb, pa]).astype(np.int64)
        eid = np.concatenate([np.arange(ei.size), np.arange(ei.size)]).astype(np.int64)
        srt = np.lexsort((cols, rows))
        rows, cols, eid_s = rows[srt], cols[srt], eid[srt]
        degree = np.bincount(rows, minlength=n).astype(np.int32)
        width = int(degree.max()) if n and ei.size else 1
        width = max(width, 1)
        n_pad = ((n + 31) // 32) * 32
        rowstart = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(degree, out=rowstart[1:])
        k = np.arange(rows.size, dtype=np.int64) - rowstart[rows]
        ell_nbr = np.tile(np.arange(n_pad, dtype=np.int32), (width, 1))
        ell_nbr[:, n:] = 0
        ell_edge = -np.ones((width, n_pad), dtype=np.int32)
        ell_nbr[k, rows] = cols
        ell_edge[k, rows] = eid_s
        flat = (k * n_pad + rows).astype(np.int32)
        inv = np.empty_like(srt)
        inv[srt] = np.arange(srt.size)
        slot_a = flat[inv[: ei.size]]
        slot_b = flat[inv[ei.size:]]
        return cls(n=n, edge_i=ei, edge_j=ej, colour=colour, order=order, pos=pos,
                   colour_start=colour_start, ell_width=width, n_pad=n_pad, ell_nbr=ell_nbr,
                   ell_edge=ell_edge, slot_a=slot_a.astype(np.int32), slot_b=slot_b.astype(np.int32),
                   degree=degree)

    @classmethod
    def pegasus(cls, m: int) -> "IsingGraph":
        n, ei, ej, col = pegasus_graph(m)
        return cls.build(n, ei, ej, col)

    @classmethod
    def zephyr(cls, m: int) -> "IsingGraph":
        n, ei, ej, col = zephyr_graph(m)
        return cls.build(n, ei, ej, col)


def greedy_get_subgraph(n_nodes: int, random_seed: Optional[int], nodes: Sequence[int],
                        adjacency: dict) -> list:
    """Densely connected ``n_nodes``-node selection from a QPU graph.

    Same procedure and the same ``random.Random(random_seed)`` draw sequence as the
    reference's ``greedy_get_subgraph`` (src/utils/common.py:22-84), expressed over a node
    list + adjacency dict (insertion-ordered neighbour lists) instead of a networkx graph.
    Returns the selected nodes in selection order.
    """
    gen = random.Random(random_seed)
    nodes = list(nodes)
    chosen = [gen.choice(nodes)]
    chosen_set = {chosen[0]}
    cap = max(len(adjacency[v]) for v in nodes)
    while len(chosen) < n_nodes:
        best_conn, best = 0, None
        want = min(cap, len(chosen))
        hit = False
        gen.shuffle(chosen)
        for v in chosen:
            nbrs = list(adjacency[v])
            gen.shuffle(nbrs)
            for u in nbrs:
                if u in chosen_set:
                    continue
                conn = sum(1 for x in adjacency[u] if x in chosen_set)
                if conn >= want:
                    hit, best = True, u
                    break
                if conn > best_conn:
                    best_conn, best = conn, u
            if hit:
                break
        chosen.append(best)
        chosen_set.add(best)
    return chosen


def greedy_get_subgraph_nx(n_nodes: int, random_seed: Optional[int], graph):
    """networkx front end with the reference's signature minus the QPU lookup
    (src/utils/common.py:22-84): returns ``graph.subgraph(selected)``, so iterating its nodes --
    and therefore :func:`get_graph_mapping` -- follows networkx's own view order exactly as in
    the reference."""
    adjacency = {v: list(graph.neighbors(v)) for v in graph.nodes()}
    chosen = greedy_get_subgraph(n_nodes, random_seed, list(graph.nodes()), adjacency)
    return graph.subgraph(chosen)


def get_graph_mapping(sub_nodes: Iterable[int]) -> dict:
    """physical qubit -> logical index 0..n-1 in iteration order (src/utils/common.py:86-100)."""
    return {phys: logical for logical, phys in enumerate(sub_nodes)}
