"""A happens-before model of the overlapped-frames schedule with the split halo flags
(csrc/bmfr_pipeline.cu, run_frame, "Strips"): the operations a rank enqueues per frame, the edges that order
them (stream order, CUDA events, the early / late flags between neighbours) and what each operation reads and
writes.  The test derives, for a line of ranks and a run of frames, that

  * the order of submission is a schedule (the graph is acyclic: nothing waits for something submitted later),
  * every read sees the version (frame) of the data it is meant to see, and no write of another version can
    land between that version's write and the read, or race with it.

Writes of the SAME version to the same rows by two operations (a rank's own reprojection and its neighbour's push
recompute identical bits for the rows both cover) are allowed to be unordered; that is the one race the design
accepts, in order as well as overlapped.

This is a restatement of the protocol, not the C++ itself: it checks the design, the GPU tests
(tests/test_sharding.py) check the implementation against the whole-frame run bit for bit.
"""
import itertools

import pytest


def build(n_ranks, n_frames, split_flags=True, wait_late_before_r=True, wait_early_push_before_r=True):
    """-> (ops, edges, accesses).  op = (rank, kind, frame); accesses[op] = [(mode, location, version)],
    location = (rank, buffer, parity, region)."""
    ops, edges, acc = [], set(), {}

    def op(r, kind, f):
        o = (r, kind, f)
        ops.append(o)
        acc[o] = []
        return o

    def nb(r):
        return [x for x in (r - 1, r + 1) if 0 <= x < n_ranks]

    for r, f in itertools.product(range(n_ranks), range(n_frames)):
        for kind in ("WR", "R", "E", "SE", "F", "WP", "P", "L", "SL"):
            op(r, kind, f)
    for r, f in itertools.product(range(n_ranks), range(n_frames)):
        q = f & 1
        o = {k: (r, k, f) for k in ("WR", "R", "E", "SE", "F", "WP", "P", "L", "SL")}
        # streams: s_r = WR R | halo = E SE | s_f = F | s_p = WP P L SL
        edges |= {(o["WR"], o["R"]), (o["E"], o["SE"]), (o["WP"], o["P"]), (o["P"], o["L"]), (o["L"], o["SL"])}
        if f + 1 < n_frames:
            edges |= {(o["R"], (r, "WR", f + 1)), (o["SE"], (r, "E", f + 1)), (o["F"], (r, "F", f + 1)), (o["SL"], (r, "WP", f + 1))}
        # events: e_r (R -> E, F), e_f (F -> the wait + post pass), e_p / e_e of frame f-2 before this frame's s_r work
        edges |= {(o["R"], o["E"]), (o["R"], o["F"]), (o["F"], o["WP"])}
        if f >= 2:
            edges.add(((r, "SL", f - 2), o["WR"]))
            if wait_early_push_before_r:
                edges.add(((r, "SE", f - 2), o["WR"]))
        # flags raised by the neighbours
        for m in nb(r):
            if split_flags:
                if f >= 1:
                    edges.add(((m, "SE", f - 1), o["WR"]))  # early >= f
                    edges.add(((m, "SL", f - 1), o["WP"]))  # late >= f
                if f >= 2 and wait_late_before_r:
                    edges.add(((m, "SL", f - 2), o["WR"]))  # late >= f - 1
            elif f >= 1:  # the in-order protocol's single flag, waited for before the reprojection
                edges.add(((m, "SL", f - 1), o["WR"]))
                edges.add(((m, "SE", f - 1), o["WR"]))
        # what the kernels and pushes touch.  state1 = accumulated noisy colour + spp, state2 = accumulated
        # filtered colour + TAA result, tmp = prev_pixels / accept / noise tile / counter / weights / min-max.
        # Regions of a rank's buffer: own rows; per neighbour m the halo rows near the boundary that the rank's
        # own kernels also compute ("k1": straddling blocks, the +-1 ring) and the far ones only pushes write.
        R, F, P, E, L = o["R"], o["F"], o["P"], o["E"], o["L"]
        halos = [(m, part) for m in nb(r) for part in ("k1", "far")]
        if f >= 1:
            acc[R] += [("r", (r, "state1", 1 - q, "own"), f - 1)] + [("r", (r, "state1", 1 - q, ("halo", m, part)), f - 1) for m, part in halos]
            acc[P] += [("r", (r, "state2", 1 - q, "own"), f - 1)] + [("r", (r, "state2", 1 - q, ("halo", m, "k1")), f - 1) for m in nb(r)]
        acc[R] += [("w", (r, "state1", q, "own"), f), ("w", (r, "tmp", q, "own"), f)] + [("w", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)]
        acc[F] += [("r", (r, "state1", q, "own"), f), ("r", (r, "tmp", q, "own"), f), ("w", (r, "tmp2", q, "own"), f)]
        acc[F] += [("r", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)]
        acc[P] += [("r", (r, "state1", q, "own"), f), ("r", (r, "tmp", q, "own"), f), ("r", (r, "tmp2", q, "own"), f), ("w", (r, "state2", q, "own"), f)]
        acc[P] += [("r", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)] + [("w", (r, "state2", q, ("halo", m, "k1")), f) for m in nb(r)]
        acc[E] += [("r", (r, "state1", q, "own"), f)]
        acc[L] += [("r", (r, "state2", q, "own"), f)]
        for m in nb(r):  # the pushes land in the neighbour's halo rows "from r"
            acc[E] += [("w", (m, "state1", q, ("halo", r, part)), f) for part in ("k1", "far")]
            acc[L] += [("w", (m, "state2", q, ("halo", r, "k1")), f)]
    return ops, edges, acc


def closure(ops, edges):
    """reach[a] = set of operations that happen after a; raises on a cycle."""
    succ = {o: [] for o in ops}
    indeg = {o: 0 for o in ops}
    for a, b in edges:
        succ[a].append(b)
        indeg[b] += 1
    order, ready = [], [o for o in ops if indeg[o] == 0]
    while ready:
        a = ready.pop()
        order.append(a)
        for b in succ[a]:
            indeg[b] -= 1
            if indeg[b] == 0:
                ready.append(b)
    if len(order) != len(ops):
        raise ValueError("cycle: an operation waits for one submitted later")
    reach = {o: set() for o in ops}
    for a in reversed(order):
        for b in succ[a]:
            reach[a].add(b)
            reach[a] |= reach[b]
    return reach


def hazards(ops, edges, acc):
    reach = closure(ops, edges)
    before = lambda a, b: b in reach[a]  # noqa: E731
    by_loc = {}
    for o in ops:
        for mode, loc, ver in acc[o]:
            by_loc.setdefault(loc, []).append((o, mode, ver))
    found = []
    for loc, items in by_loc.items():
        writes = [(o, v) for o, m, v in items if m == "w"]
        for o, m, v in items:
            if m == "r":
                if not any(wv == v and before(w, o) for w, wv in writes):
                    found.append(("read without its producer", o, loc, v))
                for w, wv in writes:
                    if wv == v:
                        continue  # the same version: ordered or identical bits
                    ok = before(o, w) if wv > v else any(w2v == v and before(w, w2) and before(w2, o) for w2, w2v in writes)
                    if not ok:
                        found.append(("read v%d can see / race with the write of v%d" % (v, wv), o, w, loc))
        for (w1, v1), (w2, v2) in itertools.combinations(writes, 2):
            if v1 != v2 and not (before(w1, w2) or before(w2, w1)):
                found.append(("unordered writes of different versions", w1, w2, loc))
    return found


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4])
def test_overlapped_schedule_is_hazard_free(n_ranks):
    ops, edges, acc = build(n_ranks, 7)
    assert hazards(ops, edges, acc) == []


def test_the_model_finds_what_the_extra_waits_are_for():
    # without "late >= f-1" before R(f) a neighbour's early push of frame f can overwrite rows that this rank's
    # post pass of frame f-2 still reads (they share the physical buffer of that parity)
    h = hazards(*build(3, 7, wait_late_before_r=False))
    assert h and any(x[1][1] in ("F", "P") for x in h)
    # the local wait for the early push of frame f-2 before R(f) turns out to be implied by the flags (that push ->
    # its signal -> the neighbour's R(f-1) -> its push and signal -> this rank's wait before R(f)); the code keeps
    # it so that a context's own write-after-read does not lean on its neighbours
    assert hazards(*build(3, 7, wait_early_push_before_r=False)) == []


def test_single_flag_in_the_overlapped_schedule_would_serialise_the_frames():
    """With the in-order protocol's single flag (raised after the post pass) R(f+1) is ordered after the
    neighbour's P(f): correct, but the overlap between a frame's post pass and the next reprojection is gone."""
    ops, edges, acc = build(2, 5, split_flags=False)
    reach = closure(ops, edges)
    assert (0, "R", 3) in reach[(1, "P", 2)]
    ops, edges, acc = build(2, 5, split_flags=True)
    reach = closure(ops, edges)
    assert (0, "R", 3) not in reach[(1, "P", 2)] and (0, "R", 3) not in reach[(0, "P", 2)]
    assert (0, "R", 4) in reach[(1, "P", 2)]  # two frames later it is
