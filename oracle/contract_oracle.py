"""TEST INFRASTRUCTURE (oracle/): CPU restatement, in plain Python, of the stage that follows the hot path in the reference --
the fix-point  do { contractCompositePaths(); removeDeadEndNodes(); } while (counter > 0)  of
OverlapGraph::buildOverlapGraphFromHashTable (OverlapGraph.cpp:211-215). It exists to pin the NEXT row of SURVEY.md 8(f)
(rank 1) before any CUDA is written for it; nothing in the product imports it. Only for small graphs (pure Python loops).

Followed statement for statement (file:line relative to MetaGenomics/OverlapGraph.cpp):
    contractCompositePaths   :669-694     matchEdgeType :19-26     isEdgePresent :1599-1607
    mergeEdges               :702-752     mergeList     :760-785   mergedEdgeOrientation :803-829
    insertEdge(Edge*)        (push_back)  removeEdge    :867-899 (swap-with-last)   removeDeadEndNodes :931-988
Not restated: the per-read location lists (updateReadLocations / removeReadLocations), which no graph field depends on.

The start state is the graph at :210 as (src, dst, overlapOffset, orientation) tuples. The reference's list order at :210 is
an artefact of its swap-with-last deletions, so this restatement starts from the canonical order
(offset, dst, orientation); tests/test_oracle_golden.py shows that the result -- the multiset of composite edges with
their read lists -- equals the reference's own on every fixture, i.e. that the stage does not depend on that order."""

DEAD_END_LENGTH = 10          # Common.h:42
TWIN = {0: 3, 1: 1, 2: 2, 3: 0}


class Edge:
    __slots__ = ("src", "dst", "orient", "off", "reads", "offs", "ors", "twin", "flow")

    def __init__(self, src, dst, orient, off, reads=(), offs=(), ors=()):
        self.src, self.dst, self.orient, self.off = src, dst, orient, off
        self.reads, self.offs, self.ors = list(reads), list(offs), list(ors)
        self.twin, self.flow = None, 0


def match_edge_type(e1, e2):                                      # :19-26
    return (e1.orient in (1, 3) and e2.orient in (2, 3)) or (e1.orient in (0, 2) and e2.orient in (0, 1))


def merged_orientation(o1, o2):                                   # :803-829
    table = {(0, 0): 0, (0, 1): 1, (1, 2): 0, (1, 3): 1, (2, 0): 2, (2, 1): 3, (3, 2): 2, (3, 3): 3}
    return table[(o1, o2)]


def merge_list(e1, e2):                                           # :760-785
    reads, offs, ors = list(e1.reads), list(e1.offs), list(e1.ors)
    reads.append(e1.dst)
    offs.append((e1.off - sum(e1.offs)) & 0xFFFF)                # vector<UINT16>
    ors.append(1 if e1.orient in (1, 3) else 0)
    return reads + e2.reads, offs + e2.offs, ors + e2.ors


class Graph:
    def __init__(self, edges, lengths):
        """edges: iterable of (src, dst, overlapOffset, orientation) at :210; lengths[id-1] = read length."""
        n = len(lengths)
        self.adj = [[] for _ in range(n + 1)]
        es = sorted((int(s), int(o), int(d), int(t)) for s, d, o, t in edges)          # canonical (src, offset, dst, orient)
        pool = {}
        for s, o, d, t in es:
            e = Edge(s, d, t, o)
            self.adj[s].append(e)
            pool.setdefault((s, d, o, t), []).append(e)
        for s, o, d, t in es:                                     # twin links (:405-417): (d, s, (UINT16)(L_d + off - L_s), twin(t))
            for e in pool[(s, d, o, t)]:
                if e.twin is not None:
                    continue
                key = (d, s, (lengths[d - 1] + o - lengths[s - 1]) & 0xFFFF, TWIN[t])
                tw = next(x for x in pool[key] if x.twin is None and x is not e)
                e.twin, tw.twin = tw, e

    def is_edge_present(self, a, b):                              # :1599-1607
        return any(e.dst == b for e in self.adj[a])

    def remove_edge(self, e):                                     # :867-899
        for lst, x in ((self.adj[e.dst], e.twin), (self.adj[e.src], e)):
            for i, y in enumerate(lst):
                if y is x:
                    lst[i] = lst[-1]
                    lst.pop()
                    break

    def merge_edges(self, e1, e2):                                # :702-752, flow == 0 before the flow is computed
        r, o, t = merge_list(e1, e2)
        fwd = Edge(e1.src, e2.dst, merged_orientation(e1.orient, e2.orient), e1.off + e2.off, r, o, t)
        r, o, t = merge_list(e2.twin, e1.twin)
        rev = Edge(e2.dst, e1.src, TWIN[fwd.orient], e2.twin.off + e1.twin.off, r, o, t)
        fwd.twin, rev.twin = rev, fwd
        self.adj[fwd.src].append(fwd)
        self.adj[rev.src].append(rev)
        self.remove_edge(e1)
        self.remove_edge(e2)

    def contract_composite_paths(self):                           # :669-694
        counter = 0
        for index in range(1, len(self.adj)):
            if len(self.adj[index]) == 2:
                e1, e2 = self.adj[index][0], self.adj[index][1]
                if not self.is_edge_present(e1.dst, e2.dst):
                    if match_edge_type(e1.twin, e2) and e1.src != e1.dst:
                        self.merge_edges(e1.twin, e2)
                        counter += 1
        return counter

    def remove_dead_end_nodes(self):                              # :931-988
        nodes = []
        for i in range(1, len(self.adj)):
            if self.adj[i]:
                flag = inn = out = 0
                for e in self.adj[i]:
                    if len(e.reads) > DEAD_END_LENGTH or e.src == e.dst:
                        flag = 1
                        break
                    if e.orient in (0, 1):
                        inn += 1
                    else:
                        out += 1
                if flag == 0 and ((inn > 0 and out == 0) or (inn == 0 and out > 0)):
                    nodes.append(i)
        for i in nodes:
            for e in list(self.adj[i]):
                self.remove_edge(e)
        return len(nodes)

    def simplify(self):                                           # :211-215
        while True:
            counter = self.contract_composite_paths()
            counter += self.remove_dead_end_nodes()
            if counter == 0:
                return self

    def edge_records(self):
        """Canonical multiset: sorted tuples (src, dst, orient, offset, reads, offsets, orientations)."""
        return sorted((e.src, e.dst, e.orient, e.off, tuple(e.reads), tuple(e.offs), tuple(e.ors)) for lst in self.adj for e in lst)
