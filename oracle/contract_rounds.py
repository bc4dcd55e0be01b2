"""TEST INFRASTRUCTURE (oracle/): the PARALLEL formulation of the reference's contraction / dead-end fix-point
(OverlapGraph.cpp:211-215, 669-785, 931-988) that the CUDA kernels of csrc/ogb_contract.cuh implement, written out in plain Python
so that its equivalence with the sequential restatement (oracle/contract_oracle.py, itself pinned to the unmodified reference's
--dump2 output) can be checked on every fixture and on random graphs without a GPU. Nothing in the product imports it.

The reference's contractCompositePaths is a sequential sweep over the nodes in ascending index with guards on the CURRENT graph
(degree 2, no self-loop, matchEdgeType, !isEdgePresent(far1, far2)). The parallel form keeps that order as a priority:

  * the graph is a CSR whose rows never grow: a contraction REPLACES the entry (A -> x) by the composite (A -> B) and (B -> x) by
    (B -> A); dead-end removal only tombstones entries. So a node's degree changes only when the node itself is contracted or when
    a dead-end pass removes edges;
  * a pass works in rounds. A node of degree 2 that has not had its turn is READY when no node of degree 2 with a smaller index
    that has not had its turn can still change what it will see: neither of its two far ends (their contraction rewrites its edges)
    nor any neighbour of a far end (a path of contractions that makes the two far ends adjacent before its turn starts at such a
    neighbour). Ready nodes take their turn together: guards on the current graph, merge or skip. Two ready nodes never share a far
    end, so no two of them touch the same row;
  * composite read lists are ropes: every contracted read owns one record per direction, an edge keeps head / tail / count / sum of
    offsets, concatenation is O(1); the lists are walked once at the end.
"""
TWIN = {0: 3, 1: 1, 2: 2, 3: 0}
DEAD_END_LENGTH = 10
MERGED = {(0, 0): 0, (0, 1): 1, (1, 2): 0, (1, 3): 1, (2, 0): 2, (2, 1): 3, (3, 2): 2, (3, 3): 3}


def match_edge_type(o1, o2):
    return (o1 in (1, 3) and o2 in (2, 3)) or (o1 in (0, 2) and o2 in (0, 1))


class Entry:
    __slots__ = ("dst", "orient", "off", "head", "tail", "count", "sumoffs", "tpos", "valid")

    def __init__(self, dst, orient, off):
        self.dst, self.orient, self.off = dst, orient, off
        self.head = self.tail = -1                 # rope of interior reads: indices into Graph.rec
        self.count = 0
        self.sumoffs = 0
        self.tpos = -1                             # position of the twin entry in row[dst]
        self.valid = True


class Graph:
    def __init__(self, edges, lengths):
        n = len(lengths)
        self.n = n
        self.row = [[] for _ in range(n + 1)]
        es = sorted((int(s), int(o), int(d), int(t)) for s, d, o, t in edges)          # canonical (src, offset, dst, orient)
        for s, o, d, t in es:
            self.row[s].append(Entry(d, t, o))
        for s in range(1, n + 1):                                                     # twin links (OverlapGraph.cpp:405-417)
            for p, e in enumerate(self.row[s]):
                if e.tpos >= 0:
                    continue
                want = ((lengths[e.dst - 1] + e.off - lengths[s - 1]) & 0xFFFF, TWIN[e.orient])
                for q, f in enumerate(self.row[e.dst]):
                    if f.tpos < 0 and f.dst == s and (f.off, f.orient) == want and not (e.dst == s and q == p):
                        e.tpos, f.tpos = q, p
                        break
                assert e.tpos >= 0, "edge without twin"
        self.rec = []                                                                  # (read, off, ori, next)

    def valid_entries(self, x):
        return [p for p, e in enumerate(self.row[x]) if e.valid]

    def append_rope(self, a_head, a_tail, b_head, b_tail):
        if a_head < 0:
            return b_head, b_tail
        if b_head < 0:
            return a_head, a_tail
        self.rec[a_tail][3] = b_head
        return a_head, b_tail

    def contract_pass(self):
        """One contractCompositePaths sweep (:669-694) as priority rounds. Returns the number of merges."""
        cand = {x for x in range(1, self.n + 1) if len(self.valid_entries(x)) == 2}   # static within the pass (rows never grow)
        pending = set(cand)
        merges = 0
        while pending:
            ready = []
            for x in pending:
                p1, p2 = self.valid_entries(x)
                A, B = self.row[x][p1].dst, self.row[x][p2].dst
                # x waits for every lower-index node of degree 2 still to come that is one of its far ends or a neighbour of one:
                # a far end that contracts changes x's own edges; a path of contractions from A to B that completes before x's turn makes
                # A and B adjacent, and such a path starts at a neighbour of A with a smaller index. (Waiting for more than strictly
                # necessary is harmless: the smallest pending index is always ready.) Consequence: two ready nodes never share a far end,
                # so the rows a ready node reads and writes belong to it alone in this round.
                blocked = False
                for F in (A, B):
                    if F != x and F in pending and F < x:
                        blocked = True
                    for e in self.row[F]:
                        if e.valid and e.dst != x and e.dst in pending and e.dst < x:
                            blocked = True
                if not blocked:
                    ready.append(x)
            assert ready, "no ready node: the smallest pending index is always ready"
            # the ready nodes take their turn "together": decisions on the graph as it is now, then all merges
            todo = []
            for x in ready:
                p1, p2 = self.valid_entries(x)
                e1, e2 = self.row[x][p1], self.row[x][p2]
                A, B = e1.dst, e2.dst
                if any(f.valid and f.dst == B for f in self.row[A]):                    # isEdgePresent(e1.dst, e2.dst) (:679)
                    continue
                t1 = self.row[A][e1.tpos]                                              # e1's twin: A -> x
                if match_edge_type(t1.orient, e2.orient) and A != x:                    # (:681)
                    todo.append((x, p1, p2))
            for x, p1, p2 in todo:
                self.merge(x, p1, p2)
                merges += 1
            pending.difference_update(ready)
        return merges

    def merge(self, x, p1, p2):
        """mergeEdges(e1.twin, e2) (:702-752) in place: (A -> x) becomes (A -> B), (B -> x) becomes (B -> A); x loses both edges."""
        e1, e2 = self.row[x][p1], self.row[x][p2]
        A, B = e1.dst, e2.dst
        pa, pb = e1.tpos, e2.tpos
        t1, t2 = self.row[A][pa], self.row[B][pb]                                     # A -> x, B -> x
        # forward: list(t1) + [x] + list(e2)   (mergeList :760-785)
        self.rec.append([x, (t1.off - t1.sumoffs) & 0xFFFF, 1 if t1.orient in (1, 3) else 0, -1])
        r = len(self.rec) - 1
        h, t = self.append_rope(t1.head, t1.tail, r, r)
        h, t = self.append_rope(h, t, e2.head, e2.tail)
        fwd = Entry(B, MERGED[(t1.orient, e2.orient)], t1.off + e2.off)
        fwd.head, fwd.tail, fwd.count = h, t, t1.count + 1 + e2.count
        fwd.sumoffs = t1.sumoffs + ((t1.off - t1.sumoffs) & 0xFFFF) + e2.sumoffs
        # reverse: list(t2) + [x] + list(e1)   (mergeEdges: mergeList(edge2->getReverseEdge(), edge1->getReverseEdge()))
        self.rec.append([x, (t2.off - t2.sumoffs) & 0xFFFF, 1 if t2.orient in (1, 3) else 0, -1])
        r = len(self.rec) - 1
        h, t = self.append_rope(t2.head, t2.tail, r, r)
        h, t = self.append_rope(h, t, e1.head, e1.tail)
        rev = Entry(A, TWIN[fwd.orient], t2.off + e1.off)
        rev.head, rev.tail, rev.count = h, t, t2.count + 1 + e1.count
        rev.sumoffs = t2.sumoffs + ((t2.off - t2.sumoffs) & 0xFFFF) + e1.sumoffs
        fwd.tpos, rev.tpos = pb, pa
        self.row[A][pa], self.row[B][pb] = fwd, rev
        e1.valid = e2.valid = False

    def dead_end_pass(self):
        """removeDeadEndNodes (:931-988): decided on a snapshot, then every edge of the chosen nodes goes together with its twin."""
        nodes = []
        for i in range(1, self.n + 1):
            es = [e for e in self.row[i] if e.valid]
            if not es:
                continue
            if any(e.count > DEAD_END_LENGTH or e.dst == i for e in es):
                continue
            inn = sum(e.orient in (0, 1) for e in es)
            if inn == 0 or inn == len(es):
                nodes.append(i)
        for i in nodes:
            for e in self.row[i]:
                if e.valid:
                    e.valid = False
                    self.row[e.dst][e.tpos].valid = False
        return len(nodes)

    def simplify(self):
        while True:
            c = self.contract_pass()
            c += self.dead_end_pass()
            if c == 0:
                return self

    def edge_records(self):
        out = []
        for s in range(1, self.n + 1):
            for e in self.row[s]:
                if not e.valid:
                    continue
                reads, offs, ors, r = [], [], [], e.head
                while r >= 0:
                    reads.append(self.rec[r][0]); offs.append(self.rec[r][1]); ors.append(self.rec[r][2])
                    r = self.rec[r][3]
                assert len(reads) == e.count
                out.append((s, e.dst, e.orient, e.off, tuple(reads), tuple(offs), tuple(ors)))
        return sorted(out)
