"""Stand-in for `pywrapfst` (OpenFst 1.5.4 Python wrapper) -- TEST INFRASTRUCTURE ONLY.

The reference's Viterbi search is `openfst.compose(T, J)` + `openfst.shortestpath(...)` over two
acceptors it writes as AT&T text into `openfst.Compiler()`
(/root/reference/script/fst_functions_wrapped.py:28-58, 172-217, 368, 387-408).  OpenFst and its
wrapper are third-party code pinned at 1.5.4 (/root/reference/README_FULL.md:45-53) and absent from
this image, so this module restates the PUBLISHED algorithms of the four entry points the
reference touches, for the tropical semiring with float32 weights (`TropicalWeight<float>`, the
`standard` arc type fstcompile produces):

  Compiler            fstcompile's text format: "src dst ilabel olabel [weight]" / "final [weight]";
                      the start state is the source of the first line; a missing weight is One() = 0;
                      weights are parsed from decimal text straight into float32.
  Fst.arcsort         stable sort of every state's arcs by ilabel / olabel.
  compose             epsilon-aware composition of two transducers matching fst1's output labels with
                      fst2's input labels (Mohri, Pereira, Riley: "Weighted automata in text and speech
                      processing", the sequence filter of OpenFst's compose.h); arc weight =
                      Times(w1, w2) = w1 + w2 in float32; result trimmed like Connect() (pywrapfst's
                      default `connect=True`).
  shortestpath        OpenFst's SingleShortestPath (shortest-path.h): generic single-source relaxation
                      with the queue AutoQueue picks for an acyclic machine (states in topological order
                      = reverse DFS finishing order, TopOrderQueue); a distance is replaced only when
                      Plus(nd, w) != nd, i.e. on a STRICT improvement; Times(d, w) is a float32 add
                      applied arc by arc; the 1-best path is rebuilt backwards, so its final state is
                      state 0 and its start state has the highest id, which is what the reference's text
                      parser relies on (fst_functions_wrapped.py:395-403).
  Fst.text            FstPrinter's tab-separated listing, start state first, One() weights omitted.

Tie-breaking between equal-cost paths follows from the state / arc order above; OpenFst's own order
depends on hash-map iteration in the caller and on its matcher choice, so ties are where this model
may differ from the real library (BASELINE.json exempts ties within 1e-6 relative).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
_last_path_weight = None
ZERO = F32(np.inf)    # tropical Zero
ONE = F32(0.0)        # tropical One


def _parse_weight(tok, py2_str=False):
    """decimal text -> float32.  py2_str: first squeeze the text through Python 2's str(float64), 12 significant
    digits, the form the reference's '%s' % weight produced under its interpreter (Python 2.7, numpy < 1.14)."""
    x = float(tok)
    if py2_str:
        x = float("%.12g" % x)
    return F32(x)


class Fst:
    """A mutable vector FST over the float32 tropical semiring."""

    def __init__(self):
        self.arcs = []          # per state: list of (ilabel, olabel, weight f32, nextstate)
        self.final = []         # per state: weight (ZERO = not final)
        self.start = -1
        self.path_weight = None   # set on shortestpath results: Times-accumulated float32 weight of the path

    # -- construction
    def add_state(self):
        self.arcs.append([])
        self.final.append(ZERO)
        return len(self.arcs) - 1

    def _reserve(self, s):
        while len(self.arcs) <= s:
            self.add_state()

    def set_start(self, s):
        self._reserve(s)
        self.start = s

    def set_final(self, s, w=ONE):
        self._reserve(s)
        self.final[s] = F32(w)

    def add_arc(self, s, ilabel, olabel, w, nextstate):
        self._reserve(max(s, nextstate))
        self.arcs[s].append((int(ilabel), int(olabel), F32(w), int(nextstate)))

    def num_states(self):
        return len(self.arcs)

    def num_arcs(self):
        return sum(len(a) for a in self.arcs)

    # -- pywrapfst surface used by the reference
    def arcsort(self, st="ilabel"):
        key = (lambda a: a[0]) if st == "ilabel" else (lambda a: a[1])
        for a in self.arcs:
            a.sort(key=key)      # OpenFst uses std::sort: the order of equal labels is unspecified there, stable here
        return self

    def verify(self):
        return True

    def weight_type(self):
        return "tropical"

    def text(self, **_kw):
        if self.start < 0:
            return ""
        order = [self.start] + [s for s in range(len(self.arcs)) if s != self.start]
        out = []
        for s in order:
            for (il, ol, w, ns) in self.arcs[s]:
                if w == ONE:
                    out.append("%d\t%d\t%d\t%d" % (s, ns, il, ol))
                else:
                    out.append("%d\t%d\t%d\t%d\t%s" % (s, ns, il, ol, _fmt(w)))
            if self.final[s] != ZERO:
                out.append("%d" % s if self.final[s] == ONE else "%d\t%s" % (s, _fmt(self.final[s])))
        return "\n".join(out) + "\n"

    def __str__(self):
        return self.text()


def _fmt(w):
    return "%g" % float(w) if np.isfinite(w) else "Infinity"


class Compiler:
    """File-like sink for AT&T text; `print >> compiler, line` in the reference writes here."""

    py2_str = False     # class-wide switch: parse weights as Python 2 would have printed them (12 digits)

    def __init__(self, *args, **kwargs):
        self._buf = []

    def write(self, s):
        self._buf.append(s)

    def compile(self):
        fst = Fst()
        text = "".join(self._buf)
        self._buf = []
        first = True
        for line in text.split("\n"):
            f = line.split()
            if not f:
                continue
            if len(f) >= 4:
                s, d = int(f[0]), int(f[1])
                w = _parse_weight(f[4], self.py2_str) if len(f) >= 5 else ONE
                fst.add_arc(s, int(f[2]), int(f[3]), w, d)
            elif len(f) == 3:        # acceptor form "src dst label"
                s, d = int(f[0]), int(f[1])
                fst.add_arc(s, int(f[2]), int(f[2]), ONE, d)
            else:
                s = int(f[0])
                fst.set_final(s, _parse_weight(f[1], self.py2_str) if len(f) == 2 else ONE)
            if first:
                fst.set_start(s)
                first = False
        return fst


def _connect(fst):
    """Connect(): keep states that are accessible from the start AND coaccessible to a final state; surviving
    states keep their relative order."""
    n = fst.num_states()
    if fst.start < 0:
        return Fst()
    acc = [False] * n
    stack = [fst.start]
    acc[fst.start] = True
    while stack:
        s = stack.pop()
        for (_, _, _, ns) in fst.arcs[s]:
            if not acc[ns]:
                acc[ns] = True
                stack.append(ns)
    rev = [[] for _ in range(n)]
    for s in range(n):
        for (_, _, _, ns) in fst.arcs[s]:
            rev[ns].append(s)
    co = [fst.final[s] != ZERO for s in range(n)]
    stack = [s for s in range(n) if co[s]]
    while stack:
        s = stack.pop()
        for p in rev[s]:
            if not co[p]:
                co[p] = True
                stack.append(p)
    keep = [acc[s] and co[s] for s in range(n)]
    if not keep[fst.start]:
        return Fst()
    remap, out = {}, Fst()
    for s in range(n):
        if keep[s]:
            remap[s] = out.add_state()
    for s in range(n):
        if not keep[s]:
            continue
        for (il, ol, w, ns) in fst.arcs[s]:
            if keep[ns]:
                out.arcs[remap[s]].append((il, ol, w, remap[ns]))
        out.final[remap[s]] = fst.final[s]
    out.start = remap[fst.start]
    return out


def compose(fst1, fst2, connect=True, **_kw):
    """fst1 o fst2 with the sequence filter: filter state 0 = anything goes, 1 = the last move was an fst1
    output-epsilon taken alone (then fst2 may not take an input-epsilon alone: the canonical order is
    "fst2's epsilons first").  States are numbered in order of discovery (breadth first from the start pair)."""
    out = Fst()
    if fst1.start < 0 or fst2.start < 0:
        return out
    ids = {}
    queue = []

    def state(q1, q2, f):
        key = (q1, q2, f)
        s = ids.get(key)
        if s is None:
            s = out.add_state()
            ids[key] = s
            queue.append(key)
            w1, w2 = fst1.final[q1], fst2.final[q2]
            if w1 != ZERO and w2 != ZERO:
                out.final[s] = F32(w1 + w2)
        return s

    out.start = state(fst1.start, fst2.start, 0)
    # input-label index of fst2's arcs per state
    by_ilabel = []
    for arcs in fst2.arcs:
        d = {}
        for a in arcs:
            d.setdefault(a[0], []).append(a)
        by_ilabel.append(d)
    head = 0
    while head < len(queue):
        q1, q2, f = queue[head]
        head += 1
        s = ids[(q1, q2, f)]
        idx2 = by_ilabel[q2]
        for (il1, ol1, w1, n1) in fst1.arcs[q1]:
            if ol1 == 0:
                # fst1 moves alone on its output epsilon
                out.arcs[s].append((il1, 0, w1, state(n1, q2, 1)))
                continue
            for (il2, ol2, w2, n2) in idx2.get(ol1, ()):
                out.arcs[s].append((il1, ol2, F32(w1 + w2), state(n1, n2, 0)))
        if f == 0:
            for (il2, ol2, w2, n2) in idx2.get(0, ()):
                # fst2 moves alone on its input epsilon
                out.arcs[s].append((0, ol2, F32(ONE + w2), state(q1, n2, 0)))
    return _connect(out) if connect else out


def _top_order(fst):
    """TopOrderVisitor: DFS from the start, arcs in order; order = reverse finishing order.  Returns None on a
    cycle (AutoQueue would then pick another discipline; the reference's lattices are acyclic)."""
    n = fst.num_states()
    WHITE, GREY, BLACK = 0, 1, 2
    color = [WHITE] * n
    finish = []
    roots = [fst.start] + [s for s in range(n) if s != fst.start]
    for root in roots:
        if color[root] != WHITE:
            continue
        color[root] = GREY
        stack = [(root, 0)]
        while stack:
            s, i = stack.pop()
            arcs = fst.arcs[s]
            if i < len(arcs):
                stack.append((s, i + 1))
                ns = arcs[i][3]
                if color[ns] == WHITE:
                    color[ns] = GREY
                    stack.append((ns, 0))
                elif color[ns] == GREY:
                    return None
            else:
                color[s] = BLACK
                finish.append(s)
    finish.reverse()
    return finish


def shortestpath(fst, weight=None, nshortest=1, **_kw):
    """SingleShortestPath + backtrace.  An empty machine (no start, or no successful path) gives an empty Fst."""
    assert nshortest == 1, "the reference only asks for the 1-best path"
    out = Fst()
    n = fst.num_states()
    if fst.start < 0 or n == 0:
        return out
    order = _top_order(fst)
    if order is None:
        raise NotImplementedError("cyclic machine: the reference's T o J is acyclic")
    dist = [ZERO] * n
    parent = [None] * n          # (previous state, arc position)
    dist[fst.start] = ONE
    f_dist, f_parent = ZERO, -1
    for s in order:
        sd = dist[s]
        if sd == ZERO:
            continue             # never enqueued: not reached
        if fst.final[s] != ZERO:
            w = F32(sd + fst.final[s])
            if w < f_dist:       # f_distance != Plus(f_distance, w)
                f_dist, f_parent = w, s
        for pos, (il, ol, aw, ns) in enumerate(fst.arcs[s]):
            w = F32(sd + aw)
            if w < dist[ns]:     # nd != Plus(nd, w): strict improvement only
                dist[ns] = w
                parent[ns] = (s, pos)
    if f_parent < 0:
        return out
    # SingleShortestPathBacktrace: walk back from the best final state; the first state added is the final one
    s, d_p, arc_to_prev = f_parent, -1, None
    while True:
        s_p = out.add_state()
        if d_p < 0:
            out.final[s_p] = fst.final[f_parent]
        else:
            il, ol, aw, _ = arc_to_prev
            out.arcs[s_p].append((il, ol, aw, d_p))
        d_p = s_p
        if parent[s] is None:
            break
        ps, pos = parent[s]
        arc_to_prev = fst.arcs[ps][pos]
        s = ps
    out.start = d_p
    out.path_weight = f_dist
    global _last_path_weight
    _last_path_weight = f_dist
    return out
