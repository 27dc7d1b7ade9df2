"""A happens-before model of the in-kernel halo exchange of strip contexts (csrc/bmfr_pipeline.cu fill_halo / run_frame,
HaloK in csrc/bmfr_kernels.h): the kernels a rank enqueues per frame — reprojection and post pass split into their zone
CTAs (which poll, push and signal) and their interior CTAs — the edges that order them (stream order, CUDA events, the
early / late flags between neighbours) and what each of them reads and writes.  The test derives, for a line of ranks and a run of frames, that

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


def build(n_ranks, n_frames, overlapped=True, wait_late_before_r=True, wait_early_before_r=True):
    """-> (ops, edges, accesses).  op = (rank, kind, frame); accesses[op] = [(mode, location, version)],
    location = (rank, buffer, parity, region).

    Kinds: Rz / Ri = the zone CTAs (rows within halo_rows of a strip edge: they poll the flags, gather from halo rows,
    push their boundary rows into the neighbours and, when the last of them is done, raise the neighbours' early flag)
    and the interior CTAs of the reprojection; F = the fit; Pz / Pi = the same split of the post pass (late flag)."""
    KINDS = ("Rz", "Ri", "F", "Pz", "Pi")
    ops, edges, acc = [], set(), {}

    def nb(r):
        return [x for x in (r - 1, r + 1) if 0 <= x < n_ranks]

    for r, f in itertools.product(range(n_ranks), range(n_frames)):
        for kind in KINDS:
            ops.append((r, kind, f))
            acc[(r, kind, f)] = []
    for r, f in itertools.product(range(n_ranks), range(n_frames)):
        q = f & 1
        o = {k: (r, k, f) for k in KINDS}
        R, P = (o["Rz"], o["Ri"]), (o["Pz"], o["Pi"])
        # within a frame: R -> F -> P (events between the three streams, or the one in-order stream)
        edges |= {(a, o["F"]) for a in R} | {(o["F"], b) for b in P}
        # across frames on a stream a kernel starts when the previous one has completed: R(f) -> R(f+1), F -> F, P -> P
        if f + 1 < n_frames:
            edges |= {(a, (r, k, f + 1)) for a in R for k in ("Rz", "Ri")} | {(o["F"], (r, "F", f + 1))}
            edges |= {(a, (r, k, f + 1)) for a in P for k in ("Pz", "Pi")}
            if not overlapped:  # one stream: the next frame's reprojection follows this frame's post pass
                edges |= {(a, (r, k, f + 1)) for a in P for k in ("Rz", "Ri")}
        if f >= 2:  # overlapped: the buffers of this parity were last read by frame f-2 (event e_p)
            edges |= {((r, k, f - 2), b) for k in ("Pz", "Pi") for b in R}
        # flags raised by the neighbours' zone CTAs, polled by this rank's zone CTAs
        for m in nb(r):
            if f >= 1:
                if wait_early_before_r:
                    edges.add(((m, "Rz", f - 1), o["Rz"]))  # early >= f
                edges.add(((m, "Pz", f - 1), o["Pz"]))      # late >= f
            if f >= 2 and wait_late_before_r:
                edges.add(((m, "Pz", f - 2), o["Rz"]))      # late >= f - 1
        # what the CTAs touch.  state1 = accumulated noisy colour + spp, state2 = accumulated filtered colour + TAA
        # result, tmp = prev_pixels / accept / noise tile / counter, tmp2 = weights / min-max.  Regions of a rank's
        # buffer: its own rows next to a neighbour ("edge", mirrored there) and the rest ("int"); per neighbour m the halo
        # rows the rank's own kernels also compute ("k1": straddling blocks, the +-1 ring) and the far ones only pushes write.
        own = ["int"] + [("edge", m) for m in nb(r)]
        halos = [(m, part) for m in nb(r) for part in ("k1", "far")]
        Rz, Ri, F, Pz, Pi = o["Rz"], o["Ri"], o["F"], o["Pz"], o["Pi"]
        if f >= 1:
            acc[Rz] += [("r", (r, "state1", 1 - q, reg), f - 1) for reg in own] + [("r", (r, "state1", 1 - q, ("halo", m, part)), f - 1) for m, part in halos]
            acc[Ri] += [("r", (r, "state1", 1 - q, reg), f - 1) for reg in own]
            acc[Pz] += [("r", (r, "state2", 1 - q, reg), f - 1) for reg in own] + [("r", (r, "state2", 1 - q, ("halo", m, "k1")), f - 1) for m in nb(r)]
            acc[Pi] += [("r", (r, "state2", 1 - q, reg), f - 1) for reg in own]
        acc[Rz] += [("w", (r, "state1", q, ("edge", m)), f) for m in nb(r)] + [("w", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)]
        acc[Rz] += [("w", (r, "tmp", q, "zone"), f)]
        acc[Ri] += [("w", (r, "state1", q, "int"), f), ("w", (r, "tmp", q, "int"), f)]
        acc[F] += [("r", (r, "state1", q, reg), f) for reg in own] + [("r", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)]
        acc[F] += [("r", (r, "tmp", q, "zone"), f), ("r", (r, "tmp", q, "int"), f), ("w", (r, "tmp2", q, "own"), f)]
        for Px, regs, tmpreg in ((Pz, [("edge", m) for m in nb(r)], "zone"), (Pi, ["int"], "int")):
            acc[Px] += [("r", (r, "state1", q, reg), f) for reg in regs] + [("r", (r, "tmp", q, tmpreg), f), ("r", (r, "tmp2", q, "own"), f)]
            acc[Px] += [("w", (r, "state2", q, reg), f) for reg in regs]
        acc[Pz] += [("r", (r, "state1", q, ("halo", m, "k1")), f) for m in nb(r)] + [("w", (r, "state2", q, ("halo", m, "k1")), f) for m in nb(r)]
        for m in nb(r):  # the pushes of the zone CTAs land in the neighbour's halo rows "from r"
            acc[Rz] += [("w", (m, "state1", q, ("halo", r, part)), f) for part in ("k1", "far")]
            acc[Pz] += [("w", (m, "state2", q, ("halo", r, "k1")), f)]
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


@pytest.mark.parametrize("overlapped", [True, False], ids=["overlapped", "in_order"])
@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4])
def test_schedule_is_hazard_free(n_ranks, overlapped):
    ops, edges, acc = build(n_ranks, 7, overlapped=overlapped)
    assert hazards(ops, edges, acc) == []


def test_the_model_finds_what_the_waits_are_for():
    # without "late >= f-1" before the zone of R(f) a neighbour's push of frame f can overwrite halo rows that this
    # rank's fit / post pass of frame f-2 still read (they share the physical buffer of that parity)
    h = hazards(*build(3, 7, wait_late_before_r=False))
    assert h and any(x[1][1] in ("F", "Pz") for x in h)
    # without "early >= f" the zone of R(f) gathers from halo rows that have not arrived, and its pushes can land under
    # the neighbour's R(f-1)
    h = hazards(*build(3, 7, wait_early_before_r=False))
    assert h and any(x[1][1] == "Rz" for x in h)
    # on one in-order stream the late wait before R is implied (the neighbour's post pass of frame f-2 precedes its
    # reprojection of frame f-1, whose flag R(f) waits for) — the kernels poll it all the same, it costs one load
    assert hazards(*build(3, 7, overlapped=False, wait_late_before_r=False)) == []


def test_frames_overlap_across_ranks():
    """The reprojection of frame f+1 waits for the neighbours' reprojection of frame f, not for their post pass of
    frame f: consecutive frames overlap on every rank."""
    ops, edges, acc = build(2, 5)
    reach = closure(ops, edges)
    assert (0, "Rz", 3) not in reach[(1, "Pz", 2)] and (0, "Rz", 3) not in reach[(0, "Pz", 2)]
    assert (0, "Rz", 4) in reach[(1, "Pz", 2)]  # two frames later it is ordered (late >= f - 1)
    assert (0, "Rz", 3) in reach[(1, "Rz", 2)]
