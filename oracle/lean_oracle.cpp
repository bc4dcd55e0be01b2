// lean_oracle.cpp -- TEST INFRASTRUCTURE (oracle/). Memory-lean, multi-threaded CPU restatement of the
// reference hot path (abiswas-odu/metagenomics, MetaGenomics/*.cpp) in the three-phase form of
// SURVEY.md Appendix A, for the configurations omega_oracle.cpp cannot hold in this container's RAM
// (BASELINE.json configs[3]: 50 M x 150 bp; the weak-scaled bench workloads). Used ONLY as the checker
// by tests/ and by tests/golden/make_full_size.py; the product (metagenomics_b200/) never links, loads
// or calls it.
//
// What is lean about it: reads are 2-bit packed rows instead of std::string, the hash table is a sorted
// array of (key, id<<2|o) entries with a directory on the key's leading bits instead of
// unordered_map<string_view, vector>, adjacency lists are 8-byte words in per-block vectors. The
// algorithm is the reference's, statement by statement where a statement exists:
//   load()            Dataset::readDataset filter + canonical strand (Dataset.cpp:155-167, 398-413),
//                     sortReads (:197-202, compareReads :16-19), removeDupicateReads (:316-345): ID = rank + 1
//   build_index()     HashTable::insertDataset / hashRead (HashTable.cpp:50-104): 4 keys per read; bucket
//                     identity = exact key, content ordered by (id, o) = insertion order (:163-195)
//   mark_contained()  OverlapGraph::markContainedReads + checkOverlapForContainedRead (OverlapGraph.cpp:225-340)
//   scan()            insertAllEdgesOfRead + checkOverlap + the orientation switch (:354-383, :529-565),
//                     every read scanning for its own out-edges (App. A.3)
//   mark()            markTransitiveEdges on the full pre-reduction graph (:574-615, App. A.4)
//   survivors()       removeTransitiveEdges as the union of flags with twins (:623-661)
//
// Pinning: tests/test_oracle_golden.py::test_lean_oracle_* compare it tuple for tuple with
// omega_oracle.cpp (itself pinned to the unmodified reference, oracle/_ref/ref_overlap) on every seeded
// data set of the parity suite and with the committed reference dumps under tests/golden/.
// All file:line citations are relative to /root/reference/MetaGenomics/.

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

template <class F> void run_threads(int threads, F f)
{
	std::vector<std::thread> ts;
	for (int t = 1; t < threads; t++) ts.emplace_back(f, t);
	f(0);
	for (auto &t : ts) t.join();
}

// Parallel sort: threads sort equal pieces, then pairwise std::inplace_merge rounds.
template <class T, class Less> void parallel_sort(std::vector<T> &v, Less less, int threads)
{
	const size_t n = v.size();
	if (threads < 2 || n < (1u << 16)) { std::sort(v.begin(), v.end(), less); return; }
	int pieces = 1;
	while (pieces * 2 <= threads) pieces *= 2;
	std::vector<size_t> cut(pieces + 1);
	for (int i = 0; i <= pieces; i++) cut[i] = n * (size_t)i / pieces;
	{
		std::vector<std::thread> ts;
		for (int i = 0; i < pieces; i++) ts.emplace_back([&, i]() { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], less); });
		for (auto &t : ts) t.join();
	}
	for (int width = 1; width < pieces; width *= 2) {
		std::vector<std::thread> ts;
		for (int i = 0; i + width < pieces; i += 2 * width)
			ts.emplace_back([&, i]() { std::inplace_merge(v.begin() + cut[i], v.begin() + cut[i + width], v.begin() + cut[std::min(pieces, i + 2 * width)], less); });
		for (auto &t : ts) t.join();
	}
}

inline u8 twin_orientation(u8 o) { return o == 0 ? 3 : (o == 3 ? 0 : o); }	// OverlapGraph.cpp:841-855
inline bool compatible(u32 t1, u32 t2)											// OverlapGraph.cpp:593-596
{
	return ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) || ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3));
}
inline u64 mix_tuple(u64 src, u64 dst, u64 offset, u64 orient)
{
	u64 x = src * 0x9E3779B97F4A7C15ULL ^ dst * 0xC2B2AE3D27D4EB4FULL ^ offset * 0x165667B19E3779F9ULL ^ orient * 0x27D4EB2F165667C5ULL;
	x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
	return x;
}

// adjacency entry: offset<<48 | dst<<16 | orient<<14 -- integer order = (offset, dst, orient), the
// deterministic form of the reference's sort key
inline u64 make_edge(u32 offset, u32 dst, u32 orient) { return ((u64)(offset & 0xFFFF) << 48) | ((u64)dst << 16) | ((u64)orient << 14); }
inline u32 e_dst(u64 e) { return (u32)(e >> 16); }
inline u32 e_orient(u64 e) { return (u32)(e >> 14) & 3; }
inline u32 e_offset(u64 e) { return (u32)(e >> 48); }

struct Ent { u64 k0, k1; u32 val; };	// key of up to 64 bases, val = id<<2 | o
inline bool ent_less(const Ent &a, const Ent &b)
{
	if (a.k0 != b.k0) return a.k0 < b.k0;
	if (a.k1 != b.k1) return a.k1 < b.k1;
	return a.val < b.val;
}

enum { BLOCK_SHIFT = 12, BLOCK = 1 << BLOCK_SHIFT };

struct Lean {
	u32 m = 0, h = 0, W = 1;
	int threads = 1;
	u64 n_good = 0, shortest = ~0ULL, longest = 0;
	u64 n = 0;
	std::vector<u64> fwd, rc;		// n rows of W words (+ slack), base k in bits 63-2(k%32).. of word k/32, A0 C1 G2 T3
	std::vector<u16> len;			// index = id-1
	std::vector<u32> freq;
	std::vector<u64> sup;			// index = id
	std::vector<Ent> ents;
	std::vector<u64> dir;			// first entry whose k0 >> (64-dir_bits) >= b
	int dir_bits = 20;
	// adjacency in blocks of BLOCK consecutive ids
	std::vector<std::vector<u64>> blk_e;
	std::vector<std::vector<u8>> blk_flag;
	std::vector<u32> deg, loc;		// index = id
	// results
	u64 E_pre = 0, E_final = 0, nodes = 0, P_c = 0, P_e = 0, C_c = 0, T = 0, active = 0, max_degree = 0, contained = 0, asym = 0;
	u64 ck_xor = 0, ck_sum = 0;
	std::vector<u32> fin;			// optional export: 4 u32 per surviving edge, canonical order
	bool keep_edges = false;

	const u64 *row_f(u64 id) const { return fwd.data() + (id - 1) * W; }
	const u64 *row_r(u64 id) const { return rc.data() + (id - 1) * W; }
	const u64 *adj(u64 id) const { return blk_e[(id - 1) >> BLOCK_SHIFT].data() + loc[id]; }
	const u8 *adj_flag(u64 id) const { return blk_flag[(id - 1) >> BLOCK_SHIFT].data() + loc[id]; }

	// 32 bases starting at base p of a row (high bits first); rows are followed by slack words
	static inline u64 bases32(const u64 *w, u32 p)
	{
		const u32 wi = p >> 5, sh = (p & 31) << 1;
		return sh ? (w[wi] << sh) | (w[wi + 1] >> (64 - sh)) : w[wi];
	}
	static inline bool region_equal(const u64 *s, u32 a, const u64 *t, u32 b, u32 cnt)
	{
		u32 k = 0;
		for (; k + 32 <= cnt; k += 32) if (bases32(s, a + k) != bases32(t, b + k)) return false;
		const u32 rem = cnt - k;
		if (rem && ((bases32(s, a + k) ^ bases32(t, b + k)) >> (64 - 2 * rem))) return false;
		return true;
	}
	inline void key_at(const u64 *w, u32 p, u64 &k0, u64 &k1) const
	{
		k0 = bases32(w, p);
		if (h <= 32) { if (h < 32) k0 &= ~0ULL << (64 - 2 * h); k1 = 0; return; }
		k1 = bases32(w, p + 32);
		if (h < 64) k1 &= ~0ULL << (128 - 2 * h);
	}

	// ---- Dataset stage
	int load(const char *bases, const u64 *offs, u64 n_raw, u32 min_overlap)
	{
		m = min_overlap;
		std::vector<u8> good(n_raw, 0);
		std::atomic<u64> next(0);
		std::vector<u64> t_long(threads, 0), t_short(threads, ~0ULL);
		run_threads(threads, [&](int t) {
			for (;;) {
				const u64 lo = next.fetch_add(65536);
				if (lo >= n_raw) break;
				const u64 hi = std::min(n_raw, lo + 65536);
				for (u64 i = lo; i < hi; i++) {
					const u64 L = offs[i + 1] - offs[i];
					if (!(L > m) || L > 65535) continue;									// Dataset.cpp:158 (strict)
					u64 cnt[4] = {0, 0, 0, 0};
					bool ok = true;
					for (u64 k = 0; k < L && ok; k++) {
						const char c = (char)(bases[offs[i] + k] & ~0x20);				// toupper for letters (:155-156)
						if (c != 'A' && c != 'C' && c != 'G' && c != 'T') ok = false;
						else cnt[(c >> 1) & 3]++;										// :407
					}
					const u64 thr = (u64)(L * .8);										// :409
					if (!ok || cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) continue;
					good[i] = 1;
					t_long[t] = std::max(t_long[t], L); t_short[t] = std::min(t_short[t], L);
				}
			}
		});
		for (int t = 0; t < threads; t++) { longest = std::max(longest, t_long[t]); shortest = std::min(shortest, t_short[t]); }
		std::vector<u64> pos(n_raw + 1, 0);
		for (u64 i = 0; i < n_raw; i++) pos[i + 1] = pos[i] + good[i];
		n_good = pos[n_raw];
		W = (u32)std::max<u64>(1, (longest + 31) / 32);
		std::vector<u64> rows((n_good + 1) * W + 4, 0);
		std::vector<u16> rlen(n_good);
		next = 0;
		run_threads(threads, [&](int) {
			std::vector<u64> a(W), b(W);
			for (;;) {
				const u64 lo = next.fetch_add(65536);
				if (lo >= n_raw) break;
				const u64 hi = std::min(n_raw, lo + 65536);
				for (u64 i = lo; i < hi; i++) {
					if (!good[i]) continue;
					const u32 L = (u32)(offs[i + 1] - offs[i]);
					std::fill(a.begin(), a.end(), 0); std::fill(b.begin(), b.end(), 0);
					for (u32 k = 0; k < L; k++) {
						u32 c = ((u32)bases[offs[i] + k] >> 1) & 3; c ^= c >> 1;		// A C G T -> 0 1 2 3 (order preserving)
						a[k >> 5] |= (u64)c << (62 - 2 * (k & 31));
						const u32 kr = L - 1 - k;										// Read.cpp:115-127: reversed, complemented
						b[kr >> 5] |= (u64)(3 - c) << (62 - 2 * (kr & 31));
					}
					const bool fw_first = std::lexicographical_compare(a.begin(), a.end(), b.begin(), b.end());	// :161-164, equal lengths
					memcpy(rows.data() + pos[i] * W, fw_first ? a.data() : b.data(), W * sizeof(u64));
					rlen[pos[i]] = (u16)L;
				}
			}
		});
		std::vector<u8>().swap(good);
		std::vector<u64>().swap(pos);
		// sortReads: std::string operator< on ACGT strings = row order (zero padding = 'A', the smallest base), a proper prefix first
		std::vector<u32> perm(n_good);
		for (u64 i = 0; i < n_good; i++) perm[i] = (u32)i;
		const u64 *R = rows.data();
		const u32 Wl = W;
		auto less = [R, Wl, &rlen](u32 x, u32 y) {
			const u64 *p = R + (u64)x * Wl, *q = R + (u64)y * Wl;
			for (u32 k = 0; k < Wl; k++) if (p[k] != q[k]) return p[k] < q[k];
			return rlen[x] < rlen[y];
		};
		parallel_sort(perm, less, threads);
		// removeDupicateReads: adjacent equal strings merge, frequency counts (:316-345)
		std::vector<u32> head;
		head.reserve(n_good);
		freq.clear();
		for (u64 i = 0; i < n_good; i++) {
			bool same = i > 0 && rlen[perm[i]] == rlen[perm[i - 1]] && memcmp(R + (u64)perm[i] * W, R + (u64)perm[i - 1] * W, W * sizeof(u64)) == 0;
			if (same) freq.back()++;
			else { head.push_back(perm[i]); freq.push_back(1); }
		}
		std::vector<u32>().swap(perm);
		n = head.size();
		fwd.assign((n + 1) * W + 4, 0); rc.assign((n + 1) * W + 4, 0); len.resize(n);
		next = 0;
		run_threads(threads, [&](int) {
			for (;;) {
				const u64 lo = next.fetch_add(65536);
				if (lo >= n) break;
				const u64 hi = std::min(n, lo + 65536);
				for (u64 i = lo; i < hi; i++) {
					const u64 *src = R + (u64)head[i] * W;
					u64 *f = fwd.data() + i * W, *r = rc.data() + i * W;
					memcpy(f, src, W * sizeof(u64));
					const u32 L = rlen[head[i]];
					len[i] = (u16)L;
					for (u32 k = 0; k < L; k++) {
						const u32 c = (u32)(f[k >> 5] >> (62 - 2 * (k & 31))) & 3, kr = L - 1 - k;
						r[kr >> 5] |= (u64)(3 - c) << (62 - 2 * (kr & 31));
					}
				}
			}
		});
		sup.assign(n + 1, 0);
		return 0;
	}

	// ---- HashTable
	int build_index()
	{
		h = m - 1;																		// HashTable.cpp:54
		if (h > 64 || h < 1) return -1;
		ents.resize(n * 4);
		std::atomic<u64> next(0);
		run_threads(threads, [&](int) {
			for (;;) {
				const u64 lo = next.fetch_add(65536);
				if (lo >= n) break;
				const u64 hi = std::min(n, lo + 65536);
				for (u64 i = lo; i < hi; i++) {
					const u64 id = i + 1;
					const u32 L = len[i];
					for (u32 o = 0; o < 4; o++) {										// :93-101: prefix/suffix of forward, prefix/suffix of reverse
						Ent &e = ents[i * 4 + o];
						key_at(o < 2 ? row_f(id) : row_r(id), (o & 1) ? L - h : 0, e.k0, e.k1);
						e.val = (u32)(id << 2) | o;
					}
				}
			}
		});
		parallel_sort(ents, ent_less, threads);
		dir_bits = n > (1u << 20) ? 26 : 16;
		dir.assign((1ull << dir_bits) + 1, 0);
		{
			u64 at = 0;
			for (u64 b = 0; b <= (1ull << dir_bits); b++) {
				while (at < ents.size() && (ents[at].k0 >> (64 - dir_bits)) < b) at++;
				dir[b] = at;
			}
		}
		return 0;
	}
	// getListOfReads (HashTable.cpp:202-221): the entries whose key equals (k0,k1): [first, last)
	inline void lookup(u64 k0, u64 k1, const Ent *&first, const Ent *&last) const
	{
		const u64 b = k0 >> (64 - dir_bits);
		const Ent *lo = ents.data() + dir[b], *hi = ents.data() + dir[b + 1];
		while (lo < hi) {
			const Ent *mid = lo + (hi - lo) / 2;
			if (mid->k0 < k0 || (mid->k0 == k0 && mid->k1 < k1)) lo = mid + 1; else hi = mid;
		}
		first = lo;
		const Ent *end = ents.data() + ents.size();
		while (lo < end && lo->k0 == k0 && lo->k1 == k1) lo++;
		last = lo;
	}

	// ---- markContainedReads (OverlapGraph.cpp:225-290)
	void mark_contained()
	{
		std::fill(sup.begin(), sup.end(), 0);
		P_c = C_c = 0; contained = 0;
		if (longest == shortest || n == 0) return;										// :228
		// best[r] = (L_i << 32 | ~i) maximal = the longest containing read, smallest id among those: what the
		// sequential loop leaves behind (first hit sets, a strictly longer one replaces, :259-268)
		std::vector<std::atomic<u64>> best(n + 1);
		for (u64 i = 0; i <= n; i++) best[i].store(0, std::memory_order_relaxed);
		std::atomic<u64> next(1), pc(0), cc(0);
		run_threads(threads, [&](int) {
			u64 lp = 0, lc = 0;
			for (;;) {
				const u64 lo = next.fetch_add(1024);
				if (lo > n) break;
				const u64 hi = std::min(n, lo + 1023);
				for (u64 i = lo; i <= hi; i++) {
					const u64 *s = row_f(i);
					const u32 L1 = len[i - 1];
					for (u32 j = 1; j < L1 - h; j++) {									// :240
						lp++;
						u64 k0, k1;
						key_at(s, j, k0, k1);
						const Ent *a, *b;
						lookup(k0, k1, a, b);
						for (; a < b; a++) {
							const u64 r = a->val >> 2; const u32 o = a->val & 3, L2 = len[r - 1];
							if (!(L1 > L2)) continue;									// :256
							const u64 *t = o < 2 ? row_f(r) : row_r(r);
							bool hit;
							if ((o & 1) == 0) hit = L1 - j - h >= L2 - h && region_equal(s, j + h, t, h, L2 - h);			// :316-321
							else hit = j >= L2 - h && region_equal(s, j - (L2 - h), t, 0, L2 - h);					// :331-336
							if (!hit) continue;
							lc++;
							const u64 cand = ((u64)L1 << 32) | (u64)(0xFFFFFFFFu - (u32)i);
							u64 cur = best[r].load(std::memory_order_relaxed);
							while (cand > cur && !best[r].compare_exchange_weak(cur, cand, std::memory_order_relaxed)) {}
						}
					}
				}
			}
			pc += lp; cc += lc;
		});
		P_c = pc; C_c = cc;
		for (u64 r = 1; r <= n; r++) {
			const u64 v = best[r].load(std::memory_order_relaxed);
			sup[r] = v ? (u64)(0xFFFFFFFFu - (u32)v) : 0;
			contained += v != 0;
		}
	}

	// ---- Phase A: insertAllEdgesOfRead for every read (OverlapGraph.cpp:529-565, checkOverlap :354-383)
	void scan()
	{
		const u64 nblk = (n + BLOCK - 1) >> BLOCK_SHIFT;
		blk_e.assign(nblk, std::vector<u64>());
		blk_flag.assign(nblk, std::vector<u8>());
		deg.assign(n + 2, 0); loc.assign(n + 2, 0);
		std::atomic<u64> next(0), pe(0), md(0);
		run_threads(threads, [&](int) {
			u64 lp = 0, lmax = 0;
			std::vector<u64> own;
			for (;;) {
				const u64 bi = next.fetch_add(1);
				if (bi >= nblk) break;
				std::vector<u64> &out = blk_e[bi];
				const u64 lo = (bi << BLOCK_SHIFT) + 1, hi = std::min(n, lo + BLOCK - 1);
				for (u64 i = lo; i <= hi; i++) {
					loc[i] = (u32)out.size();
					if (sup[i] != 0) continue;											// :548: contained reads own no edges
					const u64 *s = row_f(i);
					const u32 L1 = len[i - 1];
					own.clear();
					for (u32 j = 1; j < L1 - h; j++) {									// :534
						lp++;
						u64 k0, k1;
						key_at(s, j, k0, k1);
						const Ent *a, *b;
						lookup(k0, k1, a, b);
						for (; a < b; a++) {
							const u64 r = a->val >> 2; const u32 o = a->val & 3, L2 = len[r - 1];
							if (sup[r] != 0) continue;									// :548
							const u64 *t = o < 2 ? row_f(r) : row_r(r);
							u32 orientation, overlap;
							if ((o & 1) == 0) {											// :359-370
								if (L1 - j - h >= L2 - h) continue;
								if (!region_equal(s, j + h, t, h, L1 - j - h)) continue;
								orientation = o == 0 ? 3 : 2; overlap = L1 - j;			// :552,:554
							} else {													// :371-382
								if (L2 - h < j) continue;
								if (!region_equal(s, 0, t, L2 - h - j, j)) continue;
								orientation = o == 1 ? 0 : 1; overlap = h + j;			// :553,:555
							}
							const u32 offset = (u16)(L1 - overlap);						// :557
							own.push_back(make_edge(offset, (u32)r, orientation));
							// a self-overlap puts the edge and its twin object into the same list (:409-417)
							if (r == i) own.push_back(make_edge((u16)(L2 + offset - L1), (u32)r, twin_orientation((u8)orientation)));
						}
					}
					std::sort(own.begin(), own.end());									// :563 with the deterministic tie-break
					deg[i] = (u32)own.size();
					lmax = std::max<u64>(lmax, own.size());
					out.insert(out.end(), own.begin(), own.end());
				}
				blk_flag[bi].assign(out.size(), 0);
			}
			pe += lp;
			u64 cur = md.load();
			while (lmax > cur && !md.compare_exchange_weak(cur, lmax)) {}
		});
		P_e = pe; max_degree = md;
		E_pre = 0;
		for (u64 i = 1; i <= n; i++) E_pre += deg[i];
	}

	// ---- Phase B: markTransitiveEdges (OverlapGraph.cpp:574-615) on the complete pre-reduction lists
	void mark()
	{
		const u64 nblk = blk_e.size();
		std::atomic<u64> next(0), tt(0), ap(0);
		run_threads(threads, [&](int) {
			u64 lt = 0, la = 0;
			std::vector<u32> keys; std::vector<u8> st;
			for (;;) {
				const u64 bi = next.fetch_add(1);
				if (bi >= nblk) break;
				const u64 lo = (bi << BLOCK_SHIFT) + 1, hi = std::min(n, lo + BLOCK - 1);
				for (u64 u = lo; u <= hi; u++) {
					const u32 d = deg[u];
					if (d == 0) continue;
					const u64 *g = adj(u);
					u32 cap = 64;
					while (cap < 2 * d) cap <<= 1;
					keys.assign(cap, 0); st.assign(cap, 0);
					auto slot = [&](u32 key) { u32 s = (key * 2654435761u) & (cap - 1); while (keys[s] != 0 && keys[s] != key) s = (s + 1) & (cap - 1); return s; };
					for (u32 i = 0; i < d; i++) { const u32 s = slot(e_dst(g[i])); keys[s] = e_dst(g[i]); st[s] = 1; }		// :577-578 INPLAY
					for (u32 i = 0; i < d; i++) {										// :579
						const u32 v = e_dst(g[i]);
						if (st[slot(v)] != 1) continue;									// :583
						la++;
						const u64 *g2 = adj(v);
						const u32 d2 = deg[v];
						lt += d2;
						for (u32 k = 0; k < d2; k++) {									// :588
							const u32 s = slot(e_dst(g2[k]));
							if (keys[s] != 0 && st[s] == 1 && compatible(e_orient(g[i]), e_orient(g2[k]))) st[s] = 2;		// :591-596 ELIMINATED
						}
					}
					u8 *fl = blk_flag[bi].data() + loc[u];
					for (u32 i = 0; i < d; i++) fl[i] = st[slot(e_dst(g[i]))] == 2;		// :601-607 (the twin half is applied in survivors())
				}
			}
			tt += lt; ap += la;
		});
		T = tt; active = ap;
	}

	// ---- Phase C: removeTransitiveEdges (OverlapGraph.cpp:623-661) = union of flags with twins
	void survivors()
	{
		const u64 nblk = blk_e.size();
		std::vector<std::vector<u32>> outs(keep_edges ? nblk : 0);
		std::atomic<u64> next(0), ef(0), nn(0), cx(0), cs(0), as(0);
		run_threads(threads, [&](int) {
			u64 lef = 0, lnn = 0, lx = 0, ls = 0, las = 0;
			for (;;) {
				const u64 bi = next.fetch_add(1);
				if (bi >= nblk) break;
				const u64 lo = (bi << BLOCK_SHIFT) + 1, hi = std::min(n, lo + BLOCK - 1);
				for (u64 u = lo; u <= hi; u++) {
					const u32 d = deg[u];
					const u64 *g = adj(u);
					const u8 *fl = adj_flag(u);
					bool any = false;
					for (u32 i = 0; i < d; i++) {
						if (fl[i]) continue;
						const u32 w = e_dst(g[i]);
						const u64 *gw = adj(w);
						const u8 *fw = adj_flag(w);
						bool found = false, twin_flag = false;
						for (u32 k = 0; k < deg[w]; k++) if (e_dst(gw[k]) == u) { found = true; twin_flag = fw[k]; break; }
						if (!found) las++;
						if (twin_flag) continue;
						any = true; lef++;
						const u64 x = mix_tuple(u, w, e_offset(g[i]), e_orient(g[i]));
						lx ^= x; ls += x;
						if (keep_edges) { std::vector<u32> &o = outs[bi]; o.push_back((u32)u); o.push_back(w); o.push_back(e_offset(g[i])); o.push_back(e_orient(g[i])); }
					}
					lnn += any;
				}
			}
			ef += lef; nn += lnn; cx ^= lx; cs += ls; as += las;
		});
		E_final = ef; nodes = nn; ck_xor = cx; ck_sum = cs; asym = as;
		fin.clear();
		if (keep_edges) for (u64 b = 0; b < nblk; b++) fin.insert(fin.end(), outs[b].begin(), outs[b].end());
	}
};

}  // namespace

extern "C" {

void *lean_create(int threads) { Lean *l = new Lean(); l->threads = threads < 1 ? 1 : threads; return l; }
void lean_destroy(void *p) { delete (Lean *)p; }
int lean_load_reads(void *p, const char *bases, const uint64_t *offs, uint64_t n_raw, uint32_t min_overlap) { return ((Lean *)p)->load(bases, offs, n_raw, min_overlap); }
uint64_t lean_n_unique(void *p) { return ((Lean *)p)->n; }
uint64_t lean_n_good(void *p) { return ((Lean *)p)->n_good; }
// runs index -> containment -> scan -> mark -> survivors; keep_edges: also keep the final tuples for lean_get_edges
int lean_run(void *p, int keep_edges)
{
	Lean *l = (Lean *)p;
	l->keep_edges = keep_edges != 0;
	if (l->build_index() != 0) return -1;
	l->mark_contained();
	l->scan();
	l->mark();
	l->survivors();
	return 0;
}
// arrays of n entries, index = id-1; fnv = FNV-1a of the forward ASCII string (same as oracle_read_info)
void lean_read_info(void *p, uint64_t *sup, uint32_t *len, uint32_t *freq, uint64_t *fnv)
{
	Lean *l = (Lean *)p;
	for (u64 i = 0; i < l->n; i++) {
		if (sup) sup[i] = l->sup[i + 1];
		if (len) len[i] = l->len[i];
		if (freq) freq[i] = l->freq[i];
		if (fnv) {
			u64 hsh = 1469598103934665603ULL;
			const u64 *f = l->row_f(i + 1);
			for (u32 k = 0; k < l->len[i]; k++) { hsh ^= (unsigned char)"ACGT"[(f[k >> 5] >> (62 - 2 * (k & 31))) & 3]; hsh *= 1099511628211ULL; }
			fnv[i] = hsh;
		}
	}
}
// out[14]: n_unique, E_pre, E_final, nodes, contained, max_degree, P_e, T, active_pivots, P_c, C_c, checksum xor, checksum sum, entries without a twin
void lean_counters(void *p, uint64_t *out)
{
	Lean *l = (Lean *)p;
	const u64 v[14] = {l->n, l->E_pre, l->E_final, l->nodes, l->contained, l->max_degree, l->P_e, l->T, l->active, l->P_c, l->C_c, l->ck_xor, l->ck_sum, l->asym};
	memcpy(out, v, sizeof v);
}
uint64_t lean_n_edges(void *p) { return ((Lean *)p)->fin.size() / 4; }
void lean_get_edges(void *p, uint32_t *out) { Lean *l = (Lean *)p; if (!l->fin.empty()) memcpy(out, l->fin.data(), l->fin.size() * sizeof(u32)); }
// pre-reduction degree of every read (index = id-1)
void lean_degrees(void *p, uint32_t *out) { Lean *l = (Lean *)p; for (u64 i = 0; i < l->n; i++) out[i] = l->deg[i + 1]; }

}  // extern "C"
