// TEST INFRASTRUCTURE (oracle/): the reference's simplification fix-point at BASELINE.json sizes --
//     do { counter = contractCompositePaths(); counter += removeDeadEndNodes(); } while (counter > 0);
// (OverlapGraph.cpp:211-215) restated SEQUENTIALLY, in the reference's own order, so that it can serve as the golden for graphs of
// tens of millions of edges. It follows oracle/contract_oracle.py (the plain-Python restatement that tests/test_oracle_golden.py
// pins to the unmodified reference's --dump2 output) statement for statement:
//     contractCompositePaths :669-694   mergeEdges :702-752   mergeList :760-785   mergedEdgeOrientation :803-829
//     insertEdge (push_back)            removeEdge :867-899 (swap with last)        removeDeadEndNodes :931-988
//     matchEdgeType :19-26              isEdgePresent :1599-1607
// with one difference of representation only: the three lists of an edge are a linked list of records instead of three vectors
// (the reference copies the vectors at every merge, which is quadratic in the length of a chain: 10 s for 0.4 M reads, hours
// at 8.6 M). Nothing in it is shared with the device formulation (no rounds, no in-place rows): the sweep visits the nodes in
// ascending index and really inserts and removes edges. tests/test_contract.py pins it to the Python restatement and to the
// reference's fixtures. Nothing in the product loads it.
//
// Output: the edges in adjacency order with their lists, and an order-independent checksum of the graph (see cseq_checksum and
// tests/contract_lib.py::simplified_checksum for the same figure in numpy).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef uint32_t u32;
typedef uint64_t u64;

namespace {

const int DEAD_END_LENGTH = 10;                       // Common.h:42
const u32 TWIN[4] = {3, 1, 2, 0};

struct Rec { u32 read; uint16_t off; uint8_t ori; int next; };

struct SEdge {
	u32 src, dst;
	u64 off;
	u64 sum;            // sum of the list's offsets (mergeList :771 subtracts it from the edge's offset)
	int head, tail;     // list of reads inside the edge
	u32 count;
	int twin;
	uint8_t orient;
	bool alive;
};

struct Graph {
	std::vector<std::vector<int> > adj;
	std::vector<SEdge> e;
	std::vector<Rec> rec;
};

bool match_edge_type(const SEdge &a, const SEdge &b)  // :19-26
{
	return ((a.orient == 1 || a.orient == 3) && (b.orient == 2 || b.orient == 3)) || ((a.orient == 0 || a.orient == 2) && (b.orient == 0 || b.orient == 1));
}

int merged_orientation(int o1, int o2)               // :803-829
{
	if ((o1 == 1 || o1 == 3) && (o2 == 2 || o2 == 3)) return (o1 == 1 ? 0 : 2) + (o2 == 2 ? 0 : 1);
	if ((o1 == 0 || o1 == 2) && (o2 == 0 || o2 == 1)) return (o1 == 0 ? 0 : 2) + (o2 == 0 ? 0 : 1);
	return -1;
}

bool is_edge_present(const Graph &g, u32 a, u32 b)   // :1599-1607
{
	for (size_t i = 0; i < g.adj[a].size(); i++) if (g.e[g.adj[a][i]].dst == b) return true;
	return false;
}

void remove_from(std::vector<int> &lst, int x)
{
	for (size_t i = 0; i < lst.size(); i++)
		if (lst[i] == x) { lst[i] = lst.back(); lst.pop_back(); return; }
}

void remove_edge(Graph &g, int x)                     // :867-899: the twin out of its list, then the edge out of its own (swap with last)
{
	const int t = g.e[x].twin;
	remove_from(g.adj[g.e[x].dst], t);
	remove_from(g.adj[g.e[x].src], x);
	g.e[x].alive = false; g.e[t].alive = false;
}

// mergeList (:760-785): list(e1) + [e1.dst] + list(e2); the lists of e1 and e2 are consumed (both edges are removed right after)
void merge_list(Graph &g, const SEdge &e1, const SEdge &e2, SEdge &out)
{
	Rec r;
	r.read = e1.dst; r.off = (uint16_t)((e1.off - e1.sum) & 0xFFFF); r.ori = (e1.orient == 1 || e1.orient == 3) ? 1 : 0; r.next = e2.head;
	g.rec.push_back(r);
	const int k = (int)g.rec.size() - 1;
	if (e1.head >= 0) { g.rec[e1.tail].next = k; out.head = e1.head; } else out.head = k;
	out.tail = e2.head >= 0 ? e2.tail : k;
	out.count = e1.count + 1 + e2.count;
	out.sum = e1.sum + r.off + e2.sum;
}

void merge_edges(Graph &g, int i1, int i2)            // :702-752 (flow == 0 before the flow is computed)
{
	const SEdge e1 = g.e[i1], e2 = g.e[i2], t2 = g.e[e2.twin], t1 = g.e[e1.twin];
	SEdge f, r;
	f.src = e1.src; f.dst = e2.dst; f.orient = (uint8_t)merged_orientation(e1.orient, e2.orient); f.off = e1.off + e2.off; f.alive = true;
	merge_list(g, e1, e2, f);
	r.src = e2.dst; r.dst = e1.src; r.orient = (uint8_t)TWIN[f.orient]; r.off = t2.off + t1.off; r.alive = true;
	merge_list(g, t2, t1, r);
	const int fi = (int)g.e.size(), ri = fi + 1;
	f.twin = ri; r.twin = fi;
	g.e.push_back(f); g.e.push_back(r);
	g.adj[f.src].push_back(fi);
	g.adj[r.src].push_back(ri);
	remove_edge(g, i1);
	remove_edge(g, i2);
}

u64 contract_composite_paths(Graph &g)                // :669-694
{
	u64 counter = 0;
	for (size_t index = 1; index < g.adj.size(); index++) {
		if (g.adj[index].size() != 2) continue;
		const int a = g.adj[index][0], b = g.adj[index][1];
		if (is_edge_present(g, g.e[a].dst, g.e[b].dst)) continue;
		if (match_edge_type(g.e[g.e[a].twin], g.e[b]) && g.e[a].src != g.e[a].dst) { merge_edges(g, g.e[a].twin, b); counter++; }
	}
	return counter;
}

u64 remove_dead_end_nodes(Graph &g)                   // :931-988
{
	std::vector<u32> nodes;
	for (size_t i = 1; i < g.adj.size(); i++) {
		if (g.adj[i].empty()) continue;
		int flag = 0, in = 0, out = 0;
		for (size_t k = 0; k < g.adj[i].size(); k++) {
			const SEdge &x = g.e[g.adj[i][k]];
			if (x.count > (u32)DEAD_END_LENGTH || x.src == x.dst) { flag = 1; break; }
			if (x.orient == 0 || x.orient == 1) in++; else out++;
		}
		if (flag == 0 && ((in > 0 && out == 0) || (in == 0 && out > 0))) nodes.push_back((u32)i);
	}
	for (size_t k = 0; k < nodes.size(); k++) {
		const std::vector<int> copy = g.adj[nodes[k]];
		for (size_t j = 0; j < copy.size(); j++) if (g.e[copy[j]].alive) remove_edge(g, copy[j]);
	}
	return nodes.size();
}

inline u64 mix(u64 x) { x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32; return x; }

}  // namespace

struct cseq_result {
	u64 n_edges, n_items, merges, dead_ends, iterations, ck_xor, ck_sum;
};

// Order-independent checksum of a simplified graph: per edge  mix( mix(src*K1 ^ dst*K2 ^ offset*K3 ^ orient*K4 ^ count*K5) +
// sum over its list, position k = 1.., of mix(read*K1 ^ item_offset*K2 ^ item_orient*K3 ^ k*K4) ), then xor and sum over the edges.
static void checksum_edge(u64 src, u64 dst, u64 off, u64 orient, u64 count, u64 list_sum, u64 &x_or, u64 &x_sum)
{
	const u64 h = mix(mix(src * 0x9E3779B97F4A7C15ULL ^ dst * 0xC2B2AE3D27D4EB4FULL ^ off * 0x165667B19E3779F9ULL ^ orient * 0x27D4EB2F165667C5ULL ^ count * 0x94D049BB133111EBULL) + list_sum);
	x_or ^= h; x_sum += h;
}
static u64 checksum_item(u64 read, u64 off, u64 ori, u64 k)
{
	return mix(read * 0x9E3779B97F4A7C15ULL ^ off * 0xC2B2AE3D27D4EB4FULL ^ ori * 0x165667B19E3779F9ULL ^ k * 0x27D4EB2F165667C5ULL);
}

// edges: ne x 4 u32 (src, dst, overlapOffset, orientation), any order (sorted here into the canonical (src, offset, dst, orient));
// lens[id-1] = read length. out_edges (optional): malloc'd n_edges x 6 u64 (src, dst, orient, offset, count, list_start);
// out_items (optional): malloc'd n_items x 3 u32 (read, offset, orientation). Returns 0, or 1 when an edge has no twin.
extern "C" int cseq_simplify(const u32 *edges, u64 ne, const uint16_t *lens, u32 n, cseq_result *res, u64 **out_edges, u32 **out_items)
{
	Graph g;
	g.adj.resize((size_t)n + 1);
	g.e.reserve(2 * ne + 16);
	// canonical order: counting sort by src, then (offset, dst, orient) inside a node
	{
		std::vector<u64> first((size_t)n + 2, 0);
		for (u64 i = 0; i < ne; i++) first[edges[4 * i] + 1]++;
		for (size_t i = 1; i < first.size(); i++) first[i] += first[i - 1];
		std::vector<u64> order(ne);
		{
			std::vector<u64> cur(first.begin(), first.end() - 1);
			for (u64 i = 0; i < ne; i++) order[cur[edges[4 * i]]++] = i;
		}
		for (u32 s = 1; s <= n; s++) {
			u64 a = first[s], b = first[s + 1];
			for (u64 i = a + 1; i < b; i++) {                                    // insertion sort: rows are short
				const u64 x = order[i];
				u64 j = i;
				while (j > a) {
					const u32 *p = edges + 4 * order[j - 1], *q = edges + 4 * x;
					const bool gt = p[2] != q[2] ? p[2] > q[2] : (p[1] != q[1] ? p[1] > q[1] : p[3] > q[3]);
					if (!gt) break;
					order[j] = order[j - 1]; j--;
				}
				order[j] = x;
			}
			for (u64 i = a; i < b; i++) {
				const u32 *p = edges + 4 * order[i];
				SEdge x;
				x.src = p[0]; x.dst = p[1]; x.off = p[2]; x.orient = (uint8_t)p[3]; x.sum = 0; x.head = x.tail = -1; x.count = 0; x.twin = -1; x.alive = true;
				g.adj[s].push_back((int)g.e.size());
				g.e.push_back(x);
			}
		}
	}
	// twin links (:405-417): (d, s, (UINT16)(L_d + off - L_s), twin(t)), the first one not yet paired and not the edge itself
	for (size_t i = 0; i < g.e.size(); i++) {
		SEdge &x = g.e[i];
		if (x.twin >= 0) continue;
		const u64 woff = (u64)((lens[x.dst - 1] + x.off - lens[x.src - 1]) & 0xFFFF);
		const std::vector<int> &row = g.adj[x.dst];
		for (size_t k = 0; k < row.size(); k++) {
			SEdge &y = g.e[row[k]];
			if (y.twin < 0 && (size_t)row[k] != i && y.dst == x.src && y.off == woff && y.orient == TWIN[x.orient]) { x.twin = row[k]; y.twin = (int)i; break; }
		}
		if (x.twin < 0) return 1;
	}
	memset(res, 0, sizeof *res);
	for (;;) {                                                                  // :211-215
		u64 c = contract_composite_paths(g);
		res->merges += c;
		const u64 d = remove_dead_end_nodes(g);
		res->dead_ends += d; c += d;
		res->iterations++;
		if (c == 0) break;
	}
	for (u32 s = 1; s <= n; s++) for (size_t k = 0; k < g.adj[s].size(); k++) { res->n_edges++; res->n_items += g.e[g.adj[s][k]].count; }
	u64 *oe = out_edges ? (u64 *)malloc((res->n_edges ? res->n_edges : 1) * 6 * sizeof(u64)) : 0;
	u32 *oi = out_items ? (u32 *)malloc((res->n_items ? res->n_items : 1) * 3 * sizeof(u32)) : 0;
	u64 ei = 0, ii = 0;
	for (u32 s = 1; s <= n; s++)
		for (size_t k = 0; k < g.adj[s].size(); k++) {
			const SEdge &x = g.e[g.adj[s][k]];
			u64 list_sum = 0, pos = 0;
			if (oe) { u64 *p = oe + 6 * ei; p[0] = x.src; p[1] = x.dst; p[2] = x.orient; p[3] = x.off; p[4] = x.count; p[5] = ii; }
			for (int r = x.head; r >= 0 && pos < x.count; r = g.rec[r].next) {
				pos++;
				list_sum += checksum_item(g.rec[r].read, g.rec[r].off, g.rec[r].ori, pos);
				if (oi) { u32 *q = oi + 3 * ii; q[0] = g.rec[r].read; q[1] = g.rec[r].off; q[2] = g.rec[r].ori; }
				ii++;
			}
			if (pos != x.count) { free(oe); free(oi); return 2; }
			checksum_edge(x.src, x.dst, x.off, x.orient, x.count, list_sum, res->ck_xor, res->ck_sum);
			ei++;
		}
	if (out_edges) *out_edges = oe;
	if (out_items) *out_items = oi;
	return 0;
}

extern "C" void cseq_free(void *p) { free(p); }
