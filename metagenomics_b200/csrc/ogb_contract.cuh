// ogb_contract.cuh -- the stage that follows the graph build in the reference: the fix-point
//     do { counter = contractCompositePaths(); counter += removeDeadEndNodes(); } while (counter > 0);
// of OverlapGraph::buildOverlapGraphFromHashTable (OverlapGraph.cpp:211-215; contractCompositePaths :669-694, mergeEdges :702-752,
// mergeList :760-785, mergedEdgeOrientation :803-829, removeDeadEndNodes :931-988, matchEdgeType :19-26, isEdgePresent :1599-1607).
//
// The reference's contraction is one sequential sweep over the nodes in ascending index with guards on the CURRENT graph, so its
// result depends on that order (two parallel chains between the same end nodes: the one whose last node comes first in the sweep is
// contracted completely, the other keeps its last node because the end nodes are adjacent by then). The device form keeps the order
// as a PRIORITY instead of a sequence:
//
//   * the graph is a CSR of 32-byte entries whose rows never grow: contracting x between A and B REPLACES the entry (A -> x) by the
//     composite (A -> B) and (B -> x) by (B -> A) in place and retires x's two entries; dead-end removal only retires entries. A
//     node's degree therefore changes only when the node itself is contracted or in a dead-end pass, and the set of nodes of
//     degree 2 is fixed for the whole of a sweep;
//   * a sweep runs in ROUNDS over the nodes of degree 2 that have not had their turn. Such a node x with far ends A and B is READY
//     when no node of degree 2 with a smaller index that has not had its turn is A, B, or a neighbour of A or B. (A far end that
//     is contracted first rewrites x's own entries; a path of contractions that makes A and B adjacent before x's turn starts at a
//     neighbour of A with a smaller index. Waiting for more than strictly necessary is harmless: the smallest index is always
//     ready.) Two ready nodes are never adjacent and never share a far end, so a round is one launch with a thread per ready node
//     that evaluates the reference's guards on the current graph and merges: no two threads touch the same row, no atomics;
//   * the read lists of composite edges are ROPES: a contracted read owns one record per direction (record 2x forward, 2x+1
//     reverse; offset and orientation as mergeList computes them), an entry keeps head / tail / count / offset sum of its rope and
//     concatenation is O(1). When the fix-point is reached the ropes are ranked by in-place pointer jumping (each record holds
//     (next, hops to next) in one 64-bit word, so a jump is one consistent read and one write, no double buffer) and every record
//     is written at list_start[owner] + count - hops_to_end.
//
// The test suite states the same algorithm in plain Python (contract_rounds.py) and shows it equal to a sequential restatement and
// to the unmodified reference's dump on every fixture. The per-thread bodies below are `__host__ __device__` so that
// tests/contract_emul.cpp can run these very functions thread by thread on the CPU (a test of the logic without a GPU; the
// product only ever launches the kernels).
#ifndef OGB_CONTRACT_CUH_
#define OGB_CONTRACT_CUH_

#include <stdint.h>

#if defined(__CUDACC__)
#define OGB_HD __host__ __device__ __forceinline__
#else
#define OGB_HD inline
#endif

typedef unsigned int cu32;
typedef unsigned long long cu64;

#define OGB_C_NIL 0xFFFFFFFFu        // empty rope / end of a rope nobody owns
#define OGB_C_OWNER 0x80000000u      // next-field of a rope's last record once the owner is known: OWNER | entry index
#define OGB_C_DEAD_END_LENGTH 10     // Common.h:42
#define OGB_C_REC_USED 0x80000000u   // rec_info bit: the record is part of a rope

// One directed, possibly composite, edge in its source's row.
struct CEntry {
	cu32 dst;
	cu32 twin;       // index of the entry dst -> src (the reference's reverseEdge)
	cu64 off;        // overlapOffset (UINT64, Edge.h:25): sum over the contracted hops (mergeEdges :711)
	cu32 head, tail; // rope of the reads inside the edge: record indices, OGB_C_NIL = none
	cu32 count;      // ... how many
	uint16_t sumoffs;// sum of the rope's offsets mod 2^16: mergeList only needs (off - sum) as UINT16 (:771)
	uint8_t orient;
	uint8_t valid;
};

struct CRec { cu32 next; cu32 hops; };   // one 64-bit word: rope link, and (from ranking on) the hops that link spans

struct CGraph {
	CEntry *E;
	const cu32 *rowptr;   // n + 2: entries of read s (1-based) are E[rowptr[s] .. rowptr[s+1])
	cu32 n;               // reads
	cu64 n_entries;
	CRec *rec;            // 2 (n + 1) records
	cu32 *rec_info;       // offset | orientation << 16 | OGB_C_REC_USED
	uint8_t *state;       // n + 1: 1 = degree 2 and turn still to come in this sweep
	cu32 *cp;             // 2 (n + 1): the node's two entries in this sweep
	cu32 *blocker;        // n + 1: the node that made x wait when it was last checked (0 = none yet)
	const cu64 *meta;     // read lengths as the read store keeps them (low 16 bits of meta[id-1]); null = all uniform_len
	cu32 uniform_len;
};

OGB_HD cu32 c_len(const CGraph &G, cu32 id) { return G.meta ? (cu32)(G.meta[id - 1] & 0xFFFF) : G.uniform_len; }
OGB_HD cu32 c_twin_orient(cu32 o) { return o == 0 ? 3u : (o == 3 ? 0u : o); }
OGB_HD bool c_match_edge_type(cu32 o1, cu32 o2) { return ((o1 == 1 || o1 == 3) && (o2 == 2 || o2 == 3)) || ((o1 == 0 || o1 == 2) && (o2 == 0 || o2 == 1)); }   // :19-26
// mergedEdgeOrientation (:803-829) for the pairs matchEdgeType lets through: the left side of the first edge, the right side of the second
OGB_HD cu32 c_merged_orient(cu32 o1, cu32 o2) { return (o1 & 2u) | (o2 & 1u); }

// ---- set-up: rows from the sorted final edge list, entries, twin links (OverlapGraph.cpp:405-417) ----

// thread i = final edge i; the list is sorted by (src, offset, dst, orient)
OGB_HD void cb_rowptr(cu64 i, const ogb_edge *fin, cu64 ne, cu32 n, cu32 *rowptr)
{
	const cu32 s = fin[i].src, prev = i ? fin[i - 1].src : 0;
	for (cu32 u = prev + 1; u <= s; u++) rowptr[u] = (cu32)i;
	if (i == 0) rowptr[0] = 0;
	if (i + 1 == ne) for (cu32 u = s + 1; u <= n + 1; u++) rowptr[u] = (cu32)ne;
}

OGB_HD void cb_init_entry(cu64 i, const ogb_edge *fin, const CGraph &G)
{
	CEntry e;
	e.dst = fin[i].dst; e.twin = OGB_C_NIL; e.off = fin[i].offset; e.head = e.tail = OGB_C_NIL; e.count = 0; e.sumoffs = 0;
	e.orient = fin[i].orient; e.valid = 1;
	G.E[i] = e;
}

OGB_HD bool c_same_edge(const ogb_edge &a, const ogb_edge &b) { return a.src == b.src && a.dst == b.dst && a.offset == b.offset && a.orient == b.orient; }

// The twin of (s -> d, off, t) is (d -> s, (UINT16)(len(d) + off - len(s)), twin(t)). Identical records exist only as the two halves
// of a palindromic self-overlap (s == d, t in {1, 2}): they are each other's twins. Returns false when there is none.
OGB_HD bool cb_twin(cu64 i, const ogb_edge *fin, const CGraph &G)
{
	const ogb_edge e = fin[i];
	const cu64 r0 = G.rowptr[e.src];
	cu32 rank = 0;
	while (i - rank > r0 && c_same_edge(fin[i - rank - 1], e)) rank++;
	ogb_edge w;
	w.src = e.dst; w.dst = e.src; w.orient = (uint8_t)c_twin_orient(e.orient); w.reserved = 0;
	w.offset = (uint16_t)((c_len(G, e.dst) + e.offset - c_len(G, e.src)) & 0xFFFF);
	const cu64 a = G.rowptr[e.dst], b = G.rowptr[e.dst + 1];
	for (cu64 q = a; q < b; q++) {
		if (!c_same_edge(fin[q], w)) continue;
		const cu64 t = c_same_edge(w, e) ? q + (rank ^ 1u) : q + rank;      // q = first of the run
		if (t >= b || !c_same_edge(fin[t], w)) return false;
		G.E[i].twin = (cu32)t;
		return true;
	}
	return false;
}

// ---- one sweep of contractCompositePaths ----

// thread x = read x (1..n): the nodes of degree 2 enter the sweep. Returns true when x does (the caller appends it to the list).
OGB_HD bool cb_candidate(cu32 x, const CGraph &G)
{
	cu32 k = 0, p[2] = {0, 0};
	for (cu32 q = G.rowptr[x]; q < G.rowptr[x + 1]; q++) {
		if (!G.E[q].valid) continue;
		if (k < 2) p[k] = q;
		k++;
	}
	const bool in = k == 2;
	G.state[x] = in ? 1 : 0;
	if (in) { G.cp[2 * (cu64)x] = p[0]; G.cp[2 * (cu64)x + 1] = p[1]; G.blocker[x] = 0; }
	return in;
}

// the LARGEST node of degree 2 with an index below x and its turn still to come that is F or a neighbour of F (0 = none). Any such
// node would do; the largest one tends to wait longest itself, so x re-reads its neighbourhood less often (2.5 instead of 3.0 full
// checks per node at config 3)
OGB_HD cu32 c_row_blocker(const CGraph &G, cu32 F, cu32 x)
{
	cu32 b = (F < x && G.state[F]) ? F : 0;
	for (cu32 q = G.rowptr[F]; q < G.rowptr[F + 1]; q++) {
		const CEntry &e = G.E[q];
		if (e.valid && e.dst < x && e.dst > b && G.state[e.dst]) b = e.dst;
	}
	return b;
}

// thread k = pending node list[k]: is it ready in this round? (reads the graph only; the launch does not modify it)
// A node that had to wait remembers for whom: while that node's turn has not come it stays a far end or a neighbour of a far end of
// x (neither x's far ends nor their rows can change without a contraction of a node that x waits for anyway), so x is still
// blocked and the two rows need not be read again.
OGB_HD bool cb_ready(cu32 x, const CGraph &G)
{
	const cu32 last = G.blocker[x];
	if (last && G.state[last]) return false;
	const cu32 A = G.E[G.cp[2 * (cu64)x]].dst, B = G.E[G.cp[2 * (cu64)x + 1]].dst;
	const cu32 ba = c_row_blocker(G, A, x), bb = c_row_blocker(G, B, x), b = ba > bb ? ba : bb;
	if (b) G.blocker[x] = b;
	return b == 0;
}

OGB_HD void c_rope_append(const CGraph &G, cu32 &h, cu32 &t, cu32 bh, cu32 bt)
{
	if (bh == OGB_C_NIL) return;
	if (h == OGB_C_NIL) { h = bh; t = bt; return; }
	G.rec[t].next = bh;
	t = bt;
}

// thread = one READY node: the body of the reference's loop (:674-690) for index x on the current graph. Returns true when it merged.
OGB_HD bool cb_turn(cu32 x, const CGraph &G)
{
	G.state[x] = 0;
	const cu32 p1 = G.cp[2 * (cu64)x], p2 = G.cp[2 * (cu64)x + 1];
	const CEntry e1 = G.E[p1], e2 = G.E[p2];
	const cu32 A = e1.dst, B = e2.dst;
	for (cu32 q = G.rowptr[A]; q < G.rowptr[A + 1]; q++)                         // isEdgePresent(edge1->dst, edge2->dst) (:679)
		if (G.E[q].valid && G.E[q].dst == B) return false;
	const cu32 pa = e1.twin, pb = e2.twin;
	const CEntry t1 = G.E[pa], t2 = G.E[pb];                                    // A -> x, B -> x
	if (!(c_match_edge_type(t1.orient, e2.orient) && A != x)) return false;      // (:681)
	// mergeEdges(edge1->getReverseEdge(), edge2) (:702-752): forward A -> B = list(A->x) + [x] + list(x->B) (mergeList :760-785)
	const cu32 rf = 2 * x, rr = 2 * x + 1;
	CEntry f;
	f.dst = B; f.twin = pb; f.off = t1.off + e2.off; f.orient = (uint8_t)c_merged_orient(t1.orient, e2.orient); f.valid = 1;
	const cu32 of = (cu32)((t1.off - t1.sumoffs) & 0xFFFF);
	G.rec[rf].next = OGB_C_NIL; G.rec[rf].hops = 1;
	G.rec_info[rf] = of | ((t1.orient == 1 || t1.orient == 3) ? 0x10000u : 0u) | OGB_C_REC_USED;
	f.head = t1.head; f.tail = t1.tail;
	c_rope_append(G, f.head, f.tail, rf, rf);
	c_rope_append(G, f.head, f.tail, e2.head, e2.tail);
	f.count = t1.count + 1 + e2.count;
	f.sumoffs = (uint16_t)(t1.sumoffs + of + e2.sumoffs);
	// reverse B -> A = list(B->x) + [x] + list(x->A)  (mergeList(edge2->getReverseEdge(), edge1->getReverseEdge()), :716-718)
	CEntry r;
	r.dst = A; r.twin = pa; r.off = t2.off + e1.off; r.orient = (uint8_t)c_twin_orient(f.orient); r.valid = 1;
	const cu32 orv = (cu32)((t2.off - t2.sumoffs) & 0xFFFF);
	G.rec[rr].next = OGB_C_NIL; G.rec[rr].hops = 1;
	G.rec_info[rr] = orv | ((t2.orient == 1 || t2.orient == 3) ? 0x10000u : 0u) | OGB_C_REC_USED;
	r.head = t2.head; r.tail = t2.tail;
	c_rope_append(G, r.head, r.tail, rr, rr);
	c_rope_append(G, r.head, r.tail, e1.head, e1.tail);
	r.count = t2.count + 1 + e1.count;
	r.sumoffs = (uint16_t)(t2.sumoffs + orv + e1.sumoffs);
	G.E[pa] = f; G.E[pb] = r;
	G.E[p1].valid = 0; G.E[p2].valid = 0;
	return true;
}

// ---- removeDeadEndNodes (:931-988): decided for every node on the graph as it is, then all removed together ----

OGB_HD bool cb_dead_end(cu32 x, const CGraph &G)
{
	cu32 k = 0, in = 0;
	for (cu32 q = G.rowptr[x]; q < G.rowptr[x + 1]; q++) {
		const CEntry &e = G.E[q];
		if (!e.valid) continue;
		if (e.count > OGB_C_DEAD_END_LENGTH || e.dst == x) return false;        // a long edge or a self-loop keeps the node (:944-951)
		k++;
		in += e.orient <= 1;                                                     // 0 u<--<v and 1 u<-->v enter u (:955-958)
	}
	return k != 0 && (in == 0 || in == k);
}

OGB_HD void cb_dead_remove(cu32 x, const CGraph &G)
{
	for (cu32 q = G.rowptr[x]; q < G.rowptr[x + 1]; q++) {
		if (!G.E[q].valid) continue;
		G.E[q].valid = 0;
		G.E[G.E[q].twin].valid = 0;      // both halves of an edge between two dead ends write the same value
	}
}

// ---- result: surviving entries in row order, their ropes ranked and written out ----

// thread i = entry: what it contributes to the two scans; the rope's last record learns its owner
OGB_HD void cb_survivor(cu64 i, const CGraph &G, cu32 *keep, cu32 *items)
{
	const CEntry &e = G.E[i];
	keep[i] = e.valid ? 1 : 0;
	items[i] = e.valid ? e.count : 0;
	if (e.valid && e.count) G.rec[e.tail].next = OGB_C_OWNER | (cu32)i;
}

// thread r = record: one pointer jump. (next, hops) is read and written as one 64-bit word, so any interleaving of jumps keeps
// "hops = links from r to next" true. Returns true while r has not reached the end of its rope.
OGB_HD bool cb_jump(cu64 r, const CGraph &G)
{
	if (!(G.rec_info[r] & OGB_C_REC_USED)) return false;
	cu64 *W = reinterpret_cast<cu64 *>(G.rec);
	const cu64 a = W[r];
	const cu32 nx = (cu32)a, hops = (cu32)(a >> 32);
	if (nx & OGB_C_OWNER) return false;
	const cu64 b = W[nx];
	W[r] = (cu64)(cu32)b | ((cu64)(hops + (cu32)(b >> 32)) << 32);
	return !((cu32)b & OGB_C_OWNER);
}

// thread i = surviving entry
OGB_HD void cb_emit_edge(cu64 i, cu32 src, const CGraph &G, const cu64 *epos, const cu64 *lpos, ogb_cedge *out)
{
	const CEntry &e = G.E[i];
	ogb_cedge o;
	o.src = src; o.dst = e.dst; o.offset = e.off; o.list_start = lpos[i]; o.count = e.count; o.twin = (cu32)epos[e.twin];
	o.orient = e.orient; o.reserved[0] = o.reserved[1] = o.reserved[2] = 0; o.reserved2 = 0;
	out[epos[i]] = o;
}

// thread r = record. Returns false when the rope was not ranked to its end (more jumps needed: reported as an error by the caller).
OGB_HD bool cb_emit_item(cu64 r, const CGraph &G, const cu64 *lpos, ogb_clist_item *out)
{
	const cu32 info = G.rec_info[r];
	if (!(info & OGB_C_REC_USED)) return true;
	const CRec w = G.rec[r];
	if (w.next == OGB_C_NIL) return true;                       // rope of an edge a dead-end pass removed
	if (!(w.next & OGB_C_OWNER)) return false;
	const cu32 owner = w.next & ~OGB_C_OWNER;
	ogb_clist_item it;
	it.read = (cu32)(r >> 1); it.offset = (uint16_t)(info & 0xFFFF); it.orient = (uint8_t)((info >> 16) & 1); it.reserved = 0;
	out[lpos[owner] + G.E[owner].count - w.hops] = it;
	return true;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// Kernels: grid-stride loops over the bodies above.
// ------------------------------------------------------------------------------------------------
#define OGB_C_LOOP(i, n) for (cu64 i = (cu64)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (cu64)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(256) k_c_rowptr(const ogb_edge *__restrict__ fin, cu64 ne, cu32 n, cu32 *rowptr)
{
	OGB_C_LOOP(i, ne) cb_rowptr(i, fin, ne, n, rowptr);
}
__global__ void __launch_bounds__(256) k_c_entries(const ogb_edge *__restrict__ fin, CGraph G)
{
	OGB_C_LOOP(i, G.n_entries) cb_init_entry(i, fin, G);
}
__global__ void __launch_bounds__(256) k_c_twins(const ogb_edge *__restrict__ fin, CGraph G, cu64 *err)
{
	OGB_C_LOOP(i, G.n_entries) if (!cb_twin(i, fin, G)) atomicAdd((unsigned long long *)err, 1ull);
}
__global__ void __launch_bounds__(256) k_c_records(CGraph G)
{
	OGB_C_LOOP(r, 2 * ((cu64)G.n + 1)) { G.rec[r].next = OGB_C_NIL; G.rec[r].hops = 0; G.rec_info[r] = 0; }
}
// appends with one atomic per warp
__device__ __forceinline__ void c_append(bool take, cu32 x, cu32 *list, cu32 *cursor)
{
	const unsigned m = __ballot_sync(__activemask(), take);
	if (!take) return;
	const unsigned lane = threadIdx.x & 31, leader = __ffs(m) - 1;
	cu32 base = 0;
	if (lane == leader) base = atomicAdd(cursor, (cu32)__popc(m));
	base = __shfl_sync(m, base, leader);
	list[base + __popc(m & ((1u << lane) - 1))] = x;
}
__global__ void __launch_bounds__(256) k_c_candidates(CGraph G, cu32 *list, cu32 *cursor)
{
	const cu64 total = ((cu64)G.n + 255) / 256 * 256;                              // whole warps reach the ballot
	OGB_C_LOOP(i, total) {
		const cu32 x = (cu32)i + 1;
		const bool in = x <= G.n && cb_candidate(x, G);
		c_append(in, x, list, cursor);
	}
}
// the length of the work list stays on the device (*n_list_p), so that several rounds can be queued without a host round trip
__global__ void __launch_bounds__(256) k_c_ready(CGraph G, const cu32 *__restrict__ list, const cu32 *__restrict__ n_list_p, uint8_t *ready)
{
	const cu32 n_list = *n_list_p;
	OGB_C_LOOP(k, n_list) ready[k] = cb_ready(list[k], G) ? 1 : 0;
}
// counters: [0] merges of the sweep
__global__ void __launch_bounds__(256) k_c_turns(CGraph G, const cu32 *__restrict__ list, const cu32 *__restrict__ n_list_p, const uint8_t *__restrict__ ready, cu32 *next_list, cu32 *cursor, cu64 *counters)
{
	const cu32 n_list = *n_list_p;
	const cu64 total = ((cu64)n_list + 255) / 256 * 256;
	OGB_C_LOOP(k, total) {
		const bool live = k < n_list;
		const cu32 x = live ? list[k] : 0;
		const bool go = live && ready[k];
		if (go && cb_turn(x, G)) atomicAdd((unsigned long long *)counters, 1ull);
		c_append(live && !go, x, next_list, cursor);
	}
}
__global__ void __launch_bounds__(256) k_c_dead_ends(CGraph G, uint8_t *flag, cu64 *counters)
{
	OGB_C_LOOP(i, G.n) {
		const bool d = cb_dead_end((cu32)i + 1, G);
		flag[i + 1] = d;
		if (d) atomicAdd((unsigned long long *)counters + 1, 1ull);
	}
}
__global__ void __launch_bounds__(256) k_c_dead_remove(CGraph G, const uint8_t *__restrict__ flag)
{
	OGB_C_LOOP(i, G.n) if (flag[i + 1]) cb_dead_remove((cu32)i + 1, G);
}
__global__ void __launch_bounds__(256) k_c_survivors(CGraph G, cu32 *keep, cu32 *items, cu64 *max_count)
{
	OGB_C_LOOP(i, G.n_entries) {
		cb_survivor(i, G, keep, items);
		if (items[i]) atomicMax((unsigned long long *)max_count, (unsigned long long)items[i]);
	}
}
__global__ void __launch_bounds__(256) k_c_jump(CGraph G, cu64 *unfinished)
{
	bool more = false;
	OGB_C_LOOP(r, 2 * ((cu64)G.n + 1)) more |= cb_jump(r, G);
	if (__any_sync(__activemask(), more) && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)unfinished, 1ull);
}
__global__ void __launch_bounds__(256) k_c_emit_edges(CGraph G, const cu32 *__restrict__ keep, const cu64 *__restrict__ epos, const cu64 *__restrict__ lpos, ogb_cedge *out)
{
	OGB_C_LOOP(s, G.n) for (cu32 q = G.rowptr[s + 1]; q < G.rowptr[s + 2]; q++) if (keep[q]) cb_emit_edge(q, (cu32)s + 1, G, epos, lpos, out);
}
__global__ void __launch_bounds__(256) k_c_emit_items(CGraph G, const cu64 *__restrict__ lpos, ogb_clist_item *out, cu64 *err)
{
	OGB_C_LOOP(r, 2 * ((cu64)G.n + 1)) if (!cb_emit_item(r, G, lpos, out)) atomicAdd((unsigned long long *)err, 1ull);
}
#endif

#endif
