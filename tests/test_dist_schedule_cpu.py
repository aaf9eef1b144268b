"""The launch schedule and the role mapping of the row-partitioned LU with the exchange inside the kernels (csrc/lu_dist.cu),
replayed on CPU through the library's own host code (lls_dist_debug_schedule / lls_dist_debug_role: the functions the launcher and
the kernel use).  No GPU: what is checked is the bookkeeping that the multi-GPU runs depend on --

  * every tile (i, j) of every rank receives every panel p < min(i, j) exactly once, in increasing order, and only after that panel
    has been factored (its launch, on whatever rank owns it, precedes in the global step order);
  * every panel is factored exactly once (DIAG on the owner of its block row), every U-row tile solved once on that owner, every
    L-column tile solved once on the rank that owns its block row, each after ALL its updates;
  * a rank with tiles in a launch has an L-panel CTA in the launch that factored the newest panel it reads (that CTA is what keeps
    the launch alive until the whole panel has arrived: the update path never polls);
  * the panels that are live at the same time fall into different mailbox slots."""
import ctypes

import numpy as np
import pytest

from machupx_b200 import _native

DIAG, UROW, LCOL, OTHER, NONE = range(5)


def _schedule(lib, nblk, world, rank, pair_min):
    n = lib.lls_dist_debug_schedule(nblk, world, rank, pair_min, None, 0)
    assert n > 0
    out = (ctypes.c_int * (9 * n))()
    assert lib.lls_dist_debug_schedule(nblk, world, rank, pair_min, out, n) == n
    a = np.frombuffer(out, dtype=np.int32).reshape(n, 9)
    keys = ("k", "passes", "thin", "base", "rem", "lb0", "nloc", "own", "grid")
    return [dict(zip(keys, map(int, row))) for row in a]


def _role(lib, b, L):
    lb, bj = ctypes.c_int(0), ctypes.c_int(0)
    r = lib.lls_dist_debug_role(b, L["own"], L["rem"], L["base"], L["lb0"], L["nloc"], L["thin"], ctypes.byref(lb), ctypes.byref(bj))
    return r, lb.value, bj.value


@pytest.mark.parametrize("world,nblk,pair_min", [(1, 7, 0), (1, 9, 2), (2, 8, 0), (2, 10, 3), (3, 12, 2), (4, 12, 0), (4, 16, 5), (8, 16, 0),
                                                 (8, 24, 2), (8, 40, -1), (8, 40, -2)])
def test_every_tile_gets_every_panel_once_and_in_order(world, nblk, pair_min):
    lib = _native.load()
    plans = [_schedule(lib, nblk, world, r, pair_min) for r in range(world)]
    assert len({len(p) for p in plans}) == 1                      # every rank makes the same number of launches: the global steps line up
    nl = len(plans[0])
    applied = {}                                                  # (i, j) -> list of panels applied so far
    factored_at = {}                                              # panel -> launch index in which it was factored
    solved = {}                                                   # (i, j) panel tile -> launch index of its solve
    slots = 8
    for e in range(nl):
        # all ranks are in launch e "at the same time": first the updates of this launch, then its roles
        new_panel = None
        for rank in range(world):
            L = plans[rank][e]
            assert L["k"] == plans[0][e]["k"] and L["passes"] == plans[0][e]["passes"] and L["thin"] == plans[0][e]["thin"]
            k, passes, base = L["k"], L["passes"], L["base"]
            new_panel = base
            panels = list(range(k, k + passes)) if k >= 0 else []
            for p in panels:
                assert p in factored_at and factored_at[p] < e, "panel %d read in launch %d before it was factored" % (p, e)
            # panels whose slots are live in this launch: the ones read and the one written
            live = panels + [base]
            assert len({p % slots for p in live}) == len(live)           # slot = global step index modulo the slot count
            seen = set()
            for b in range(L["grid"]):
                role, lb, bj = _role(lib, b, L)
                if role == NONE:
                    continue
                bi = lb * world + rank
                assert base <= bi < nblk and base <= bj < nblk and (bi, bj) not in seen
                seen.add((bi, bj))
                got = applied.setdefault((bi, bj), [])
                for p in panels:
                    assert got == list(range(p)), ((bi, bj), got, p)          # exactly the panels 0 .. p - 1 so far, in order
                    got.append(p)
                if role == DIAG:
                    assert (bi, bj) == (base, base) and base % world == rank and got == list(range(base))
                    assert base not in factored_at
                    factored_at[base] = e
                elif role == UROW:
                    assert bi == base and bj > base and base % world == rank and got == list(range(base))
                    assert (bi, bj) not in solved
                    solved[(bi, bj)] = e
                elif role == LCOL:
                    assert bj == base and bi > base and bi % world == rank and got == list(range(base))
                    assert (bi, bj) not in solved
                    solved[(bi, bj)] = e
                else:
                    assert bi > base and bj > base and not L["thin"]
            # the no-polling rule: a rank that will read panel `base` later has an L-panel CTA (or is its owner) in this launch
            later_rows = [i for i in range(base + 1, nblk) if i % world == rank]
            if later_rows:
                has_lcol = any(_role(lib, b, L)[0] == LCOL for b in range(min(L["grid"], L["own"] * L["rem"] + len(later_rows) + 1)))
                assert has_lcol
        assert factored_at.get(new_panel) == e
    # at the end: every panel factored, every panel tile solved, every trailing tile fully updated
    assert sorted(factored_at) == list(range(nblk))
    for i in range(nblk):
        for j in range(nblk):
            if i == j:
                continue
            assert (i, j) in solved, (i, j)
            assert applied[(i, j)] == list(range(min(i, j))), ((i, j), applied[(i, j)])
    for p in range(nblk):
        assert applied.get((p, p), []) == list(range(p))


def test_default_pair_threshold_follows_the_exchange():
    lib = _native.load()
    # with a multicast mapping pairs are used from ~ 64 sqrt(P) block rows on; without one they are switched off beyond four ranks
    def pairs(world, nblk, pm):
        return sum(1 for L in _schedule(lib, nblk, world, 0, pm) if L["passes"] == 2)
    assert pairs(1, 128, -1) == (128 - 64) // 2 + (0 if (128 - 64) % 2 == 0 else 1) or pairs(1, 128, -1) > 20
    assert pairs(8, 512, -1) > 100 and pairs(8, 512, -2) == 0
    assert pairs(4, 512, -2) > 100 and pairs(8, 64, -1) == 0
    assert lib.lls_dist_debug_schedule(10, 4, 0, 0, None, 0) == -1            # block rows must divide evenly over the ranks
