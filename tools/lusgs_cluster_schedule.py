#!/usr/bin/env python
"""Prototype of the clustered LU-SGS schedule planned for round 2 — host logic only, runs on CPU.

For an n^3 block in OpenFOAM cell order (i fastest) it builds the (slab, intra-slab level) schedule, checks the two
properties the device protocol relies on, and replays a scalar Gauss-Seidel forward sweep in schedule order to show that
the result is bit-identical to the sequential sweep in cell order (every row subtracts its lower neighbours in ascending
face id, exactly like lusgs.C:230-290):

  (P1) inside a slab every lower neighbour of a row of level l lies in level l - 1 of the same slab  -> one cluster barrier
       per level is enough;
  (P2) the lower neighbour across the slab face lies in the previous slab at intra-slab level l + s - 1, i.e. the previous
       slab produced it s - 1 levels before this slab needs it when slab m runs s levels behind slab m - 1.

    python tools/lusgs_cluster_schedule.py [n] [n_clusters]
"""
import sys

import numpy as np


def schedule(n, n_clusters):
    s = int(np.ceil(n / n_clusters))
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    cell = (i + n * (j + n * k)).ravel()
    slab = (k // s).ravel()
    level = (i + j + k % s).ravel()
    # global step at which a row is processed: slab m starts s levels after slab m - 1
    step = level + slab * s
    return s, cell, slab, level, step, (i.ravel(), j.ravel(), k.ravel())


def check_and_replay(n=24, n_clusters=6, seed=0):
    s, cell, slab, level, step, (i, j, k) = schedule(n, n_clusters)
    idx = lambda a, b, c: a + n * (b + n * c)
    lower = []                      # lower neighbours in ascending face id: (i,j,k-1), (i,j-1,k), (i-1,j,k)
    for (di, dj, dk) in ((0, 0, -1), (0, -1, 0), (-1, 0, 0)):
        ok = (i + di >= 0) & (j + dj >= 0) & (k + dk >= 0)
        nb = np.where(ok, idx(i + di, j + dj, k + dk), -1)
        lower.append(nb)
    pos_of_cell = np.empty(n ** 3, np.int64)
    pos_of_cell[cell] = np.arange(n ** 3)
    for nb in lower:
        has = nb >= 0
        q = pos_of_cell[nb[has]]
        same = slab[q] == slab[has]
        assert (level[q][same] == level[has][same] - 1).all(), "P1 violated"
        assert (slab[q][~same] == slab[has][~same] - 1).all() and (level[q][~same] == level[has][~same] + s - 1).all(), "P2 violated"
        assert (step[q] < step[has]).all(), "a neighbour would be needed before it exists"
    # replay: x_c = (b_c - sum_lower a_cq x_q) * rD_c, sequential in cell order vs in schedule order
    rng = np.random.default_rng(seed)
    a = [rng.standard_normal(n ** 3) for _ in lower]
    b, rD = rng.standard_normal(n ** 3), 1.0 / (4.0 + rng.random(n ** 3))

    def sweep(order):
        x = np.full(n ** 3, np.nan)
        for c in order:
            acc = b[c]
            for nb, coef in zip(lower, a):
                if nb[c] >= 0:
                    acc -= coef[c] * x[nb[c]]
            x[c] = acc * rD[c]
        return x

    seq = sweep(np.arange(n ** 3))
    sched = sweep(cell[np.lexsort((cell, step))])
    assert np.array_equal(seq, sched), "schedule order changes the result"
    chain = step.max() + 1
    return s, chain


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    nc = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    s, chain = check_and_replay(n, nc)
    print(f"n = {n}: slab thickness {s}, {int(np.ceil(n / s))} slabs, dependent chain {chain} steps (hyperplane sweep: {3 * n - 2} levels through L2); "
          f"P1, P2 hold and the replayed sweep is bit-identical to the sequential one")


# ------------------------------------------------------------------------------------------------------------------
# General meshes: the same schedule from the LDU addressing alone (owner < neighbour on every internal face), which is what
# the device set-up will have to build.  Slabs are contiguous cell-id ranges — any such cut is dependency-monotone because a
# row's lower neighbours have smaller ids — intra-slab levels are longest paths inside the slab, and every slab gets the
# earliest start step that (a) satisfies all of its cross-slab dependencies with `lag` steps of slack for the trip through L2
# and (b) follows the previous slab of the same cluster.
def aligned_bounds(n_cells, owner, neighbour, n_slabs):
    """Slab boundaries near the equal-size cuts, moved (by at most half a slab) to where the fewest faces cross the cut.  In
    OpenFOAM's structured numbering those are the k-planes: a cut inside a plane would make rows depend on same-plane rows of
    the previous slab that sit many levels up, which delays the whole slab."""
    owner, neighbour = np.asarray(owner), np.asarray(neighbour)
    d = np.zeros(n_cells + 2, np.int64)
    np.add.at(d, owner + 1, 1)            # a face (o, n) crosses every cut position b with o < b <= n
    np.add.at(d, neighbour + 1, -1)
    crossing = np.cumsum(d)[: n_cells + 1]
    nominal = (np.arange(n_slabs + 1) * n_cells) // n_slabs
    half = max(n_cells // n_slabs // 2, 1)
    bounds = [0]
    for b in nominal[1:-1]:
        lo, hi = max(bounds[-1] + 1, b - half), min(n_cells - 1, b + half)
        if lo > hi:
            continue
        cand = np.arange(lo, hi + 1)
        best = cand[crossing[cand] == crossing[cand].min()]
        bounds.append(int(best[np.argmin(np.abs(best - b))]))
    bounds.append(n_cells)
    return np.array(bounds, np.int64)


def general_schedule(n_cells, owner, neighbour, n_slabs, n_clusters=18, lag=2, align=True):
    owner, neighbour = np.asarray(owner), np.asarray(neighbour)
    assert (owner < neighbour).all()
    bounds = aligned_bounds(n_cells, owner, neighbour, n_slabs) if align else (np.arange(n_slabs + 1) * n_cells) // n_slabs
    n_slabs = len(bounds) - 1
    slab = np.searchsorted(bounds, np.arange(n_cells), side="right") - 1
    # lower neighbours of every cell in ascending face id (faces are sorted by owner, so that is ascending neighbour id too)
    order = np.argsort(neighbour, kind="stable")
    lo_ptr = np.zeros(n_cells + 1, np.int64)
    np.add.at(lo_ptr, neighbour + 1, 1)
    lo_ptr = np.cumsum(lo_ptr)
    lo_cell, lo_face = owner[order], order
    level = np.zeros(n_cells, np.int64)
    for c in range(n_cells):
        for q in lo_cell[lo_ptr[c]:lo_ptr[c + 1]]:
            if slab[q] == slab[c] and level[q] + 1 > level[c]:
                level[c] = level[q] + 1
    depth = np.array([level[bounds[m]:bounds[m + 1]].max() + 1 if bounds[m + 1] > bounds[m] else 0 for m in range(n_slabs)])
    off = np.zeros(n_slabs, np.int64)
    for m in range(n_slabs):
        start = 0 if m < n_clusters else off[m - n_clusters] + depth[m - n_clusters]      # a cluster works through its slabs in turn
        for c in range(bounds[m], bounds[m + 1]):
            for q in lo_cell[lo_ptr[c]:lo_ptr[c + 1]]:
                if slab[q] != m:
                    start = max(start, off[slab[q]] + level[q] + lag - level[c])
        off[m] = start
    step = off[slab] + level
    return {"slab": slab, "level": level, "step": step, "off": off, "depth": depth, "lo_ptr": lo_ptr, "lo_cell": lo_cell, "lo_face": lo_face,
            "chain": int(step.max()) + 1, "lag": lag}


def check_general(n_cells, owner, neighbour, sched, seed=0):
    """P1': intra-slab lower neighbours sit at a strictly lower level (one cluster barrier per level suffices);
    P2': cross-slab lower neighbours were produced at least `lag` steps earlier; replay == sequential sweep, bit for bit."""
    slab, level, step, lo_ptr, lo_cell, lo_face = (sched[k] for k in ("slab", "level", "step", "lo_ptr", "lo_cell", "lo_face"))
    for c in range(n_cells):
        for q in lo_cell[lo_ptr[c]:lo_ptr[c + 1]]:
            if slab[q] == slab[c]:
                assert level[q] < level[c], "P1' violated"
            else:
                assert slab[q] < slab[c] and step[q] + sched["lag"] <= step[c], "P2' violated"
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(len(owner))
    b, rD = rng.standard_normal(n_cells), 1.0 / (4.0 + rng.random(n_cells))

    def sweep(order):
        x = np.full(n_cells, np.nan)
        for c in order:
            acc = b[c]
            for q, f in zip(lo_cell[lo_ptr[c]:lo_ptr[c + 1]], lo_face[lo_ptr[c]:lo_ptr[c + 1]]):
                acc -= a[f] * x[q]
            x[c] = acc * rD[c]
        return x

    seq = sweep(np.arange(n_cells))
    sched_order = np.lexsort((np.arange(n_cells), step))
    assert np.array_equal(seq, sweep(sched_order)), "schedule order changes the result"
    # hyperplane levels of the whole mesh (today's sweep) for comparison
    glob = np.zeros(n_cells, np.int64)
    for c in range(n_cells):
        lows = lo_cell[lo_ptr[c]:lo_ptr[c + 1]]
        if len(lows):
            glob[c] = glob[lows].max() + 1
    widths = np.bincount(step)
    return {"chain": sched["chain"], "global_levels": int(glob.max()) + 1, "max_rows_per_step": int(widths.max()),
            "max_rows_per_slab_level": int(max(np.bincount(level[slab == m]).max() for m in range(len(sched["off"])) if (slab == m).any()))}
