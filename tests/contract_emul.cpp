// TEST INFRASTRUCTURE: runs the per-thread bodies of metagenomics_b200/csrc/ogb_contract.cuh -- the very functions the CUDA
// kernels call -- thread by thread on the CPU, in the launch order of ogb_graph_simplify (csrc/ogb_device.cu), so that the logic of
// the device simplification is checked against the reference's fixtures where no GPU is present. The product never links or
// calls this file; the -m gpu tests check the kernels themselves.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -I include -o tests/_build/libcontract_emul.so tests/contract_emul.cpp
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/ogb.h"
#include "../metagenomics_b200/csrc/ogb_contract.cuh"

extern "C" int emul_simplify(const ogb_edge *fin, uint64_t ne, const uint16_t *lens, uint32_t n, int reverse_threads, ogb_cedge **out_edges, uint64_t *n_edges,
                             ogb_clist_item **out_items, uint64_t *n_items, uint64_t *stats /* merges, dead ends, iterations, rounds, jumps */)
{
	std::vector<CEntry> E(ne ? ne : 1);
	std::vector<cu32> rowptr((size_t)n + 2, 0), info(2 * ((size_t)n + 1), 0), cp(2 * ((size_t)n + 1), 0), blocker((size_t)n + 1, 0);
	std::vector<CRec> rec(2 * ((size_t)n + 1));
	std::vector<uint8_t> state((size_t)n + 1, 0), flag((size_t)n + 2, 0);
	std::vector<cu64> meta(n ? n : 1);
	for (uint32_t i = 0; i < n; i++) meta[i] = lens[i];
	CGraph G;
	G.E = E.data(); G.rowptr = rowptr.data(); G.n = n; G.n_entries = ne; G.rec = rec.data(); G.rec_info = info.data(); G.state = state.data(); G.cp = cp.data();
	G.meta = meta.data(); G.uniform_len = 0; G.blocker = blocker.data();
	const cu64 nrec = 2 * ((cu64)n + 1);
	for (cu64 r = 0; r < nrec; r++) { rec[r].next = OGB_C_NIL; rec[r].hops = 0; }
	for (cu64 i = 0; i < ne; i++) cb_rowptr(i, fin, ne, n, rowptr.data());
	for (cu64 i = 0; i < ne; i++) cb_init_entry(i, fin, G);
	for (cu64 i = 0; i < ne; i++) if (!cb_twin(i, fin, G)) return 1;
	memset(stats, 0, 5 * sizeof(uint64_t));
	std::vector<cu32> list, next, ready;
	// a launch = the same body for every index; the order the threads run in must not matter, so the harness can run them backwards
	auto order = [&](size_t k, size_t cnt) { return reverse_threads ? cnt - 1 - k : k; };
	while (ne) {
		stats[2]++;
		uint64_t merges = 0, dead = 0;
		list.clear();
		for (cu32 k = 0; k < n; k++) { const cu32 x = (cu32)order(k, n) + 1; if (cb_candidate(x, G)) list.push_back(x); }
		while (!list.empty()) {
			ready.assign(list.size(), 0);
			for (size_t k = 0; k < list.size(); k++) { const size_t j = order(k, list.size()); ready[j] = cb_ready(list[j], G); }
			next.clear();
			for (size_t k = 0; k < list.size(); k++) {
				const size_t j = order(k, list.size());
				if (ready[j]) merges += cb_turn(list[j], G); else next.push_back(list[j]);
			}
			if (next.size() >= list.size()) return 2;
			list.swap(next);
			stats[3]++;
		}
		for (cu32 k = 0; k < n; k++) { const cu32 x = (cu32)order(k, n) + 1; flag[x] = cb_dead_end(x, G); dead += flag[x]; }
		for (cu32 k = 0; k < n; k++) { const cu32 x = (cu32)order(k, n) + 1; if (flag[x]) cb_dead_remove(x, G); }
		stats[0] += merges; stats[1] += dead;
		if (merges + dead == 0) break;
	}
	std::vector<cu32> keep(ne ? ne : 1), items(ne ? ne : 1);
	std::vector<cu64> epos(ne + 1, 0), lpos(ne + 1, 0);
	for (cu64 i = 0; i < ne; i++) cb_survivor(order(i, ne), G, keep.data(), items.data());
	cu64 te = 0, tl = 0;
	for (cu64 i = 0; i < ne; i++) { epos[i] = te; lpos[i] = tl; te += keep[i]; tl += items[i]; }
	for (bool more = tl != 0; more;) {
		more = false;
		for (cu64 r = 0; r < nrec; r++) more |= cb_jump(order(r, nrec), G);
		if (++stats[4] > 64) return 3;
	}
	ogb_cedge *oe = (ogb_cedge *)malloc((te ? te : 1) * sizeof(ogb_cedge));
	ogb_clist_item *oi = (ogb_clist_item *)malloc((tl ? tl : 1) * sizeof(ogb_clist_item));
	memset(oi, 0xFF, (tl ? tl : 1) * sizeof(ogb_clist_item));
	for (cu32 s = 0; s < n; s++) for (cu32 q = rowptr[s + 1]; q < rowptr[s + 2]; q++) if (keep[q]) cb_emit_edge(q, s + 1, G, epos.data(), lpos.data(), oe);
	if (tl) for (cu64 r = 0; r < nrec; r++) if (!cb_emit_item(r, G, lpos.data(), oi)) return 4;      // no items: no rope was ranked
	*out_edges = oe; *n_edges = te; *out_items = oi; *n_items = tl;
	return 0;
}

extern "C" void emul_free(void *p) { free(p); }
