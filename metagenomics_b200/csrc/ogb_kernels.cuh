// ogb_kernels.cuh -- hand-written sm_100a kernels of the overlap-graph build.
//
//   K0  k_pack_ascii / k_pack_words   Read::setRead + reverseComplement          (Read.cpp:75-127)
//   K1  k_hash_insert                 HashTable::hashRead + insertIntoTable       (HashTable.cpp:88-195)
//   K2  k_scan<MODE_CONTAIN>          markContainedReads + checkOverlapForContainedRead (OverlapGraph.cpp:225-340)
//   K3  k_scan<MODE_OVERLAP>          insertAllEdgesOfRead + checkOverlap + per-node sort (OverlapGraph.cpp:354-383,529-565)
//   K5  k_mark                        markTransitiveEdges                         (OverlapGraph.cpp:574-615)
//   K6  k_twin_keep / k_compact       removeTransitiveEdges                       (OverlapGraph.cpp:623-661)
//       k_lookup_*                    HashTable::getListOfReads                   (HashTable.cpp:202-221)
//
// Everything is integer / bit work bounded by HBM (random 32-byte sector gathers into the index
// and the packed read store); there is no dense contraction, hence no tensor-core code.
//
// Data layout in HBM
//   packed reads  u64 words, base k of a strand in bits 63-2(k%32)..62-2(k%32) of word k/32,
//                 A0 C1 G2 T3 (complement = 3-x). Read idx (= id-1): forward strand at word
//                 offset off, reverse complement at off + pw, pw = 2*ceil(L/64) words (16-byte
//                 aligned strands; 100 bp -> 32 B = one sector). Uniform-length data sets use
//                 off = idx*2*pw (no metadata load); mixed lengths use meta[idx] = off<<16 | L.
//   index         nb buckets x 8 slots x u64 (one 64-byte DRAM burst per bucket: measured on B200, a
//                 random 32-byte sector read costs a 64-byte HBM fetch anyway, profiles/exp_r1_l2fetch.txt).
//                 slot = fp32<<32 | id<<2 | o, 0 = empty. One slot per (key,value); a key's entries
//                 sit in its home bucket and, when that is full, in the following buckets (linear
//                 probing by bucket). Only a 32-bit fingerprint of the key is stored: every consumer
//                 verifies the complete overlap (window included) against the packed reads, which
//                 makes the result exact and independent of hash values.
//   edges         u64 = offset<<48 | dst<<16 | orient<<8, so integer order = (offset,dst,orient).
//   nodes         u64 = start<<24 | degree  (adjacency of a node is contiguous and sorted).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define OGB_SLOTS 8             // slots per bucket (64 bytes = one HBM burst)
#define OGB_WARPS 8             // warps per block in the scan / mark kernels
#ifndef OGB_SCAN_BLOCKS
#define OGB_SCAN_BLOCKS 4       // resident blocks per SM the scan kernel is compiled for (register cap)
#endif
#define OGB_HQ 288              // per-warp candidate queue (<= 31 left over + 32 lanes x 8 slots)
#define OGB_STAGE_WORDS 32      // query reads up to 1024 bp are staged in shared memory (longer: slow path)
#define OGB_EC 256              // per-warp edge buffer (reads with more edges take the slow path)
#define OGB_SETCAP 512          // per-warp neighbour set slots in shared memory (degree <= 256)
#define OGB_DEG_BITS 24
#define OGB_DEG_MASK 0xFFFFFFull
#define OGB_NODE_OVERFLOW OGB_DEG_MASK   // degree field value marking "take the slow path"

enum { MODE_OVERLAP = 0, MODE_CONTAIN = 1 };

// device counters (u64 each)
enum {
	CTR_EDGE_CURSOR = 0, CTR_OVERFLOW, CTR_PROBES, CTR_SECTORS, CTR_CANDIDATES, CTR_CONTAIN_HITS,
	CTR_PIVOT_ENTRIES, CTR_ACTIVE_PIVOTS, CTR_MAX_DEGREE, CTR_N_CONTAINED, CTR_NODES_FINAL,
	CTR_ASYMMETRIC, CTR_SCRATCH_CURSOR, CTR_SCRATCH_FAIL, CTR_EDGES_DROPPED, CTR_PAD, CTR_COUNT
};

struct ReadStore {
	const u64 *words;
	const u64 *meta;     // null when uniform
	u32 n;
	u32 uniform_len;     // 0 when lengths differ
	u32 uniform_pw;      // padded words per strand when uniform
};

struct Table {
	u64 *slots;
	u32 nb;              // buckets
	u32 h;               // hashStringLength = minOverlap-1 (HashTable.cpp:54)
};

__device__ __forceinline__ u32 padded_words(u32 L) { return ((L + 63) >> 6) << 1; }

__device__ __forceinline__ void read_geom(const ReadStore &R, u32 idx, u64 &off, u32 &L)
{
	if (R.uniform_len) { L = R.uniform_len; off = (u64)idx * (2 * R.uniform_pw); }
	else { u64 m = __ldg(R.meta + idx); L = (u32)(m & 0xFFFF); off = m >> 16; }
}

// Load flavours. Random gathers (index buckets, partner reads) bypass L1 allocation so that they do
// not evict anything useful; the query read itself is staged in shared memory.
__device__ __forceinline__ u64 ld_na(const u64 *p)
{
	u64 v;
	asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
struct LdShared { static __device__ __forceinline__ u64 ld(const u64 *p) { return *p; } };
struct LdGlobal { static __device__ __forceinline__ u64 ld(const u64 *p) { return __ldg(p); } };
struct LdStream { static __device__ __forceinline__ u64 ld(const u64 *p) { return ld_na(p); } };

__device__ __forceinline__ u64 funnel(u64 a, u64 b, u32 sh) { return sh ? ((a << sh) | (b >> (64 - sh))) : a; }

// 32 bases starting at base p of a packed strand (high bits first). The words after a strand are
// always allocated, so loads past its end are safe; callers mask what they do not need.
template <class LD> __device__ __forceinline__ u64 extract32(const u64 *__restrict__ w, u32 p)
{
	u32 wi = p >> 5, sh = (p & 31) << 1;
	return funnel(LD::ld(w + wi), LD::ld(w + wi + 1), sh);
}

__device__ __forceinline__ u64 mix64(u64 acc, u64 x)
{
	acc ^= x;
	acc *= 0xff51afd7ed558ccdULL;
	acc ^= acc >> 33;
	return acc;
}

// Multilinear hash of a key of up to 64 bases held in (k0,k1): sum of 32-bit digits times odd 64-bit
// constants, mod 2^64. The high half (strongly universal) picks the bucket, the low half is the
// stored fingerprint. (The reference's polynomial-mod hash, HashTable.cpp:135-155, is not
// reproduced: its value is unobservable in the result.)
__device__ __forceinline__ u64 hash2(u64 k0, u64 k1)
{
	u64 acc = 0x9E3779B97F4A7C15ULL;
	acc += (u64)(u32)(k0 >> 32) * 0xD6E8FEB86659FD93ULL;
	acc += (u64)(u32)k0 * 0xA0761D6478BD642FULL;
	acc += (u64)(u32)(k1 >> 32) * 0xE7037ED1A0B428DBULL;
	acc += (u64)(u32)k1 * 0x8EBC6AF09C88C6E3ULL;
	return acc;
}

// Hash of the h bases starting at base p of a packed strand.
template <class LD> __device__ __forceinline__ u64 key_hash(const u64 *__restrict__ w, u32 p, u32 h)
{
	const u32 wi = p >> 5, sh = (p & 31) << 1;
	const u64 a = LD::ld(w + wi), b = LD::ld(w + wi + 1);
	u64 k0 = funnel(a, b, sh);
	if (h <= 32) return hash2(k0 & (~0ULL << (64 - 2 * h)), 0);
	const u64 c = LD::ld(w + wi + 2);
	u64 k1 = funnel(b, c, sh);
	if (h <= 64) return hash2(k0, k1 & (~0ULL << (128 - 2 * h)));
	u64 acc = hash2(k0, k1);
	u32 rem = h - 64;
	p += 64;
	for (; rem > 32; rem -= 32, p += 32) acc = mix64(acc, extract32<LD>(w, p));
	return mix64(acc, extract32<LD>(w, p) & (~0ULL << (64 - 2 * rem)));
}

__device__ __forceinline__ u32 hash_fp(u64 hash) { u32 f = (u32)hash; return f ? f : 1u; }   // 0 is reserved for "empty"
__device__ __forceinline__ u32 bucket_of(u64 hash, u32 nb) { return __umulhi((u32)(hash >> 32), nb); }

// s[a..a+len) == t[b..b+len) on packed strands, streaming one new word per side and 32 bases. No
// early exit: (almost) every candidate verifies, and independent iterations keep the sector
// requests of the partner read in flight together.
template <class LDS, class LDT>
__device__ __forceinline__ bool region_equal(const u64 *__restrict__ s, u32 a, const u64 *__restrict__ t, u32 b, u32 len)
{
	const u32 sa = (a & 31) << 1, sb = (b & 31) << 1;
	const u64 *ws = s + (a >> 5), *wt = t + (b >> 5);
	u64 s0 = LDS::ld(ws), t0 = LDT::ld(wt), diff = 0;
	u32 k = 0;
	for (; k + 32 <= len; k += 32) {
		u64 s1 = LDS::ld(++ws), t1 = LDT::ld(++wt);
		diff |= funnel(s0, s1, sa) ^ funnel(t0, t1, sb);
		s0 = s1; t0 = t1;
	}
	u32 rem = len - k;
	if (rem) diff |= (funnel(s0, LDS::ld(ws + 1), sa) ^ funnel(t0, LDT::ld(wt + 1), sb)) & (~0ULL << (64 - 2 * rem));
	return diff == 0;
}

__device__ __forceinline__ u64 make_edge(u32 offset, u32 dst, u32 orient) { return ((u64)offset << 48) | ((u64)dst << 16) | ((u64)orient << 8); }
__device__ __forceinline__ u32 edge_dst(u64 e) { return (u32)(e >> 16); }
__device__ __forceinline__ u32 edge_orient(u64 e) { return (u32)(e >> 8) & 3; }
__device__ __forceinline__ u32 edge_offset(u64 e) { return (u32)(e >> 48); }
// OverlapGraph.cpp:593-596: the pivot is entered and left on the same strand.
__device__ __forceinline__ bool compatible(u32 t1, u32 t2) { return ((t1 & 1) == ((t2 >> 1) & 1)); }
// OverlapGraph.cpp:841-855
__device__ __forceinline__ u32 twin_orient(u32 o) { return o == 0 ? 3 : (o == 3 ? 0 : o); }

// ------------------------------------------------------------------------------------------------
// K0: pack + reverse complement.
// ------------------------------------------------------------------------------------------------

// One thread per (read, output word). ASCII input, already validated upper-case ACGT.
__global__ void k_pack_ascii(const char *__restrict__ bases, const u64 *__restrict__ offsets, u64 *__restrict__ words,
                             const u64 *__restrict__ meta, u32 n, u32 uniform_len, u32 uniform_pw, u32 max_pw)
{
	u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u32 idx = (u32)(tid / max_pw), k = (u32)(tid % max_pw);
	if (idx >= n) return;
	u64 off; u32 L;
	if (uniform_len) { L = uniform_len; off = (u64)idx * (2 * uniform_pw); }
	else { u64 m = meta[idx]; L = (u32)(m & 0xFFFF); off = m >> 16; }
	u32 pw = padded_words(L);
	if (k >= pw) return;
	const char *s = bases + offsets[idx];
	u64 fw = 0, rc = 0;
	for (u32 i = 0; i < 32; i++) {
		u32 p = k * 32 + i;
		if (p < L) {
			// A=0x41 C=0x43 G=0x47 T=0x54: (c>>1)&3 = 0,1,3,2 (the reference's code, HashTable.cpp:149);
			// x ^ (x>>1) turns that into the order-preserving 0,1,2,3.
			u32 c = ((u32)s[p] >> 1) & 3; c ^= c >> 1;
			fw |= (u64)c << (62 - 2 * i);
			u32 d = ((u32)s[L - 1 - p] >> 1) & 3; d ^= d >> 1;
			rc |= (u64)(3 - d) << (62 - 2 * i);
		}
	}
	words[off + k] = fw;
	words[off + pw + k] = rc;
}

__device__ __forceinline__ u64 reverse_groups(u64 x)
{
	x = __brevll(x);
	return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);
}

// One thread per (read, output word). Input: tightly packed forward words (host Dataset layout).
__global__ void k_pack_words(const u64 *__restrict__ in_words, const u64 *__restrict__ in_offsets, const unsigned short *__restrict__ lens,
                             u64 *__restrict__ words, const u64 *__restrict__ meta, u32 n, u32 uniform_len, u32 uniform_pw, u32 max_pw)
{
	u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u32 idx = (u32)(tid / max_pw), k = (u32)(tid % max_pw);
	if (idx >= n) return;
	u32 L = lens[idx];
	u64 off;
	if (uniform_len) off = (u64)idx * (2 * uniform_pw); else off = meta[idx] >> 16;
	u32 pw = padded_words(L), nw = (L + 31) >> 5;
	if (k >= pw) return;
	const u64 *src = in_words + in_offsets[idx];
	u64 fw = k < nw ? src[k] : 0;
	// rc bases [32k, 32k+32) = complement of forward bases (L-32k-32 .. L-32k], reversed
	u64 rc = 0;
	if (k < nw) {
		int hi = (int)L - 32 * (int)k;       // exclusive end in forward coordinates, >= 1
		int lo = hi - 32;
		u32 cnt = 32;
		if (lo < 0) { cnt = (u32)hi; lo = 0; }
		u32 wi = (u32)lo >> 5, sh = ((u32)lo & 31) << 1;
		u64 a = src[wi], b = (wi + 1 < nw) ? src[wi + 1] : 0;
		u64 x = sh ? ((a << sh) | (b >> (64 - sh))) : a;       // forward bases lo..lo+31 at the top
		if (cnt < 32) x &= ~0ULL << (64 - 2 * cnt);            // keep forward bases lo..hi-1
		u64 r = ~reverse_groups(x);                            // reversed + complemented; valid groups are the LOW cnt
		if (cnt < 32) r <<= (64 - 2 * cnt);
		rc = cnt < 32 ? (r & (~0ULL << (64 - 2 * cnt))) : r;
	}
	words[off + k] = fw;
	words[off + pw + k] = rc;
}

// ------------------------------------------------------------------------------------------------
// K1: hash insert. One thread per (read, orientation): o=0 prefix(fwd), 1 suffix(fwd),
// 2 prefix(rc), 3 suffix(rc) (HashTable.cpp:93-101). Claims the first empty slot along the probe
// sequence with a 64-bit CAS. Slot order inside a bucket is arbitrary (unobservable).
// ------------------------------------------------------------------------------------------------
__global__ void k_hash_insert(ReadStore R, Table T)
{
	u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (tid >= (u64)R.n * 4) return;
	u32 idx = (u32)(tid >> 2), o = (u32)(tid & 3);
	u64 off; u32 L;
	read_geom(R, idx, off, L);
	const u64 *w = R.words + off + (o >> 1) * padded_words(L);
	u32 p = (o & 1) ? L - T.h : 0;
	u64 hash = key_hash<LdGlobal>(w, p, T.h);
	u64 val = ((u64)hash_fp(hash) << 32) | ((u64)(idx + 1) << 2) | o;
	u32 b = bucket_of(hash, T.nb);
	for (;;) {
		u64 *slot = T.slots + (u64)b * OGB_SLOTS;
		// one L2-coherent look at the whole bucket, then a CAS on its first empty slot; buckets fill
		// front to back, so a lost race just moves on to the next slot
		u64 cur[OGB_SLOTS];
		#pragma unroll
		for (int q = 0; q < OGB_SLOTS; q += 2)
			asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(cur[q]), "=l"(cur[q + 1]) : "l"(slot + q) : "memory");
		int s = 0;
		#pragma unroll
		for (int q = OGB_SLOTS - 1; q >= 0; q--) s = cur[q] == 0 ? q : s;
		bool done = false;
		if (cur[OGB_SLOTS - 1] == 0)
			for (; s < OGB_SLOTS && !done; s++) done = atomicCAS(slot + s, 0ull, val) == 0;
		if (done) break;
		b = (b + 1 == T.nb) ? 0 : b + 1;
	}
}

// One bucket = 64 bytes = two 256-bit non-allocating loads.
__device__ __forceinline__ void load_bucket(const u64 *__restrict__ slots, u32 b, u64 (&s)[OGB_SLOTS])
{
	const u64 *p = slots + (u64)b * OGB_SLOTS;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s[0]), "=l"(s[1]), "=l"(s[2]), "=l"(s[3]) : "l"(p));
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s[4]), "=l"(s[5]), "=l"(s[6]), "=l"(s[7]) : "l"(p + 4));
}

// ------------------------------------------------------------------------------------------------
// K2 / K3: sliding-window scan, one warp per query read.
//
// Phase A  lane l probes windows j = 1+l, 33+l, ... (j = 1 .. L-h-1, OverlapGraph.cpp:534): key hash
//          -> bucket sector -> fingerprint compare; matches go to a per-warp candidate queue in
//          shared memory (ballot + popc compaction, warp-synchronous).
// Phase B  whenever >= 32 candidates are queued (and at the end) each lane verifies one candidate
//          against the partner's packed strand (one random sector for 100 bp) -- checkOverlap /
//          checkOverlapForContainedRead restated on packed words.
// Phase C  (overlap mode) verified edges are sorted by (offset,dst,orient) in shared memory
//          (bitonic), a contiguous range of the global edge array is claimed with one atomicAdd and
//          the node record start<<24|deg is written: the adjacency comes out sorted (:563) with no
//          global sort. Reads with more than OGB_EC edges are queued for k_scan_big.
// ------------------------------------------------------------------------------------------------

struct ScanArgs {
	ReadStore R;
	Table T;
	u32 lo, hi;                 // query read indices [lo, hi) (this rank's shard)
	const u32 *contained;       // bitmap by read index (null when no read is contained)
	u64 *sup;                   // MODE_CONTAIN: per read idx, max over hits of (L_super<<32 | ~super_idx)
	u64 *edges;                 // MODE_OVERLAP outputs
	u64 edge_cap;
	u64 *nodes;
	u32 *overflow_list;
	u32 overflow_cap;
	u64 *ctr;
};

// Verifies candidate (j, val) of query read qi (strand words s, length L1). Returns the number of
// edges produced (0, 1, or 2 for a self-overlap) in e0/e1; in MODE_CONTAIN performs the atomicMax.
template <int MODE, class LDS>
__device__ __forceinline__ int verify_candidate(const ScanArgs &A, const u64 *__restrict__ s, u32 qi, u32 L1, u32 j, u32 val, u64 &e0, u64 &e1)
{
	const u32 h = A.T.h;
	u32 ri = (val >> 2) - 1, o = val & 3;
	u64 roff; u32 L2;
	read_geom(A.R, ri, roff, L2);
	const u64 *t = A.R.words + roff + (o >> 1) * padded_words(L2);
	if (MODE == MODE_CONTAIN) {
		// OverlapGraph.cpp:256: read1 must be longer; :302-340 restated on the whole of read2.
		if (L1 <= L2) return 0;
		u32 a;
		if ((o & 1) == 0) { if (L1 - j < L2) return 0; a = j; }              // :316-321
		else { if (j < L2 - h) return 0; a = j - (L2 - h); }                // :331-336
		if (!region_equal<LDS, LdStream>(s, a, t, 0, L2)) return 0;
		atomicMax(A.sup + ri, ((u64)L1 << 32) | (u64)(0xFFFFFFFFu - qi));   // :259-268
		return 1;
	} else {
		if (A.contained && ((__ldg(A.contained + (ri >> 5)) >> (ri & 31)) & 1)) return 0;   // :548 superReadID == 0
		u32 a, b, len, orient, offset;
		if ((o & 1) == 0) {               // key = prefix of t: s[j..L1) must equal t[0..L1-j)      (:359-370)
			if (L1 - j >= L2) return 0;
			a = j; b = 0; len = L1 - j;
			orient = o == 0 ? 3 : 2;      // :552,:554
			offset = j;                   // L1 - overlap, overlap = L1 - j
		} else {                          // key = suffix of t: s[0..j+h) must equal t[L2-h-j..L2)  (:371-382)
			if (L2 - h < j) return 0;
			a = 0; b = L2 - h - j; len = h + j;
			orient = o == 1 ? 0 : 1;      // :553,:555
			offset = L1 - h - j;          // L1 - overlap, overlap = h + j
		}
		if (!region_equal<LDS, LdStream>(s, a, t, b, len)) return 0;
		e0 = make_edge(offset & 0xFFFF, ri + 1, orient);
		if (ri != qi) return 1;
		// Self-overlap: the reference inserts the edge and its twin object into the same list
		// (OverlapGraph.cpp:409-417); twin offset = (UINT16)(L2 + offset - L1) = offset.
		e1 = make_edge(offset & 0xFFFF, ri + 1, twin_orient(orient));
		return 2;
	}
}

// In-place ascending bitonic sort of n u64 keys in shared memory by one warp (n <= cap, cap a power
// of two; the tail is padded with ~0).
__device__ __forceinline__ void warp_sort(u64 *buf, u32 n, u32 lane)
{
	u32 m = 32;
	while (m < n) m <<= 1;
	for (u32 i = n + lane; i < m; i += 32) buf[i] = ~0ull;
	__syncwarp();
	for (u32 k = 2; k <= m; k <<= 1) {
		for (u32 jj = k >> 1; jj > 0; jj >>= 1) {
			for (u32 t = lane; t < (m >> 1); t += 32) {
				u32 i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
				u32 l = i | jj;
				bool up = (i & k) == 0;
				u64 x = buf[i], y = buf[l];
				if ((x > y) == up) { buf[i] = y; buf[l] = x; }
			}
			__syncwarp();
		}
	}
}

// Ascending bitonic sort of one u64 per lane with shuffles (pad unused lanes with ~0).
__device__ __forceinline__ u64 warp_sort32(u64 v, u32 lane)
{
	#pragma unroll
	for (u32 k = 2; k <= 32; k <<= 1) {
		#pragma unroll
		for (u32 jj = k >> 1; jj > 0; jj >>= 1) {
			u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, jj);
			bool keep_min = ((lane & jj) == 0) == ((lane & k) == 0);
			v = (keep_min == (v < o)) ? v : o;
		}
	}
	return v;
}

template <int MODE>
__global__ void __launch_bounds__(OGB_WARPS * 32, OGB_SCAN_BLOCKS) k_scan(ScanArgs A)
{
	__shared__ u64 s_hq[OGB_WARPS][OGB_HQ];
	__shared__ u64 s_edges[MODE == MODE_OVERLAP ? OGB_WARPS : 1][MODE == MODE_OVERLAP ? OGB_EC : 1];
	__shared__ u64 s_read[OGB_WARPS][OGB_STAGE_WORDS + 4];
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gw = blockIdx.x * OGB_WARPS + wib, nwarps = gridDim.x * OGB_WARPS;
	u64 *hq = s_hq[wib];
	u64 *eb = s_edges[MODE == MODE_OVERLAP ? wib : 0];
	const u32 h = A.T.h;
	u64 c_probes = 0, c_sectors = 0, c_cand = 0, c_hits = 0;
	u32 c_maxdeg = 0;

	for (u32 qi = A.lo + gw; qi < A.hi; qi += nwarps) {
		if (MODE == MODE_OVERLAP && A.contained && ((__ldg(A.contained + (qi >> 5)) >> (qi & 31)) & 1)) {
			if (lane == 0) A.nodes[qi] = 0;                                  // contained reads have no edges (:548)
			continue;
		}
		u64 off; u32 L1;
		read_geom(A.R, qi, off, L1);
		if (L1 > OGB_STAGE_WORDS * 32) {                                     // very long read: slow path
			if (lane == 0) {
				u64 pos = atomicAdd(A.ctr + CTR_OVERFLOW, 1ull);
				if (pos < A.overflow_cap) A.overflow_list[pos] = qi;
				if (MODE == MODE_OVERLAP) A.nodes[qi] = OGB_NODE_OVERFLOW;
			}
			continue;
		}
		// stage the query strand (+2 words of slack for the window extraction) in shared memory
		u64 *s = s_read[wib];
		{
			const u32 nw = (L1 + 31) >> 5;
			__syncwarp();
			for (u32 i = lane; i < nw + 3; i += 32) s[i] = i < nw ? __ldg(A.R.words + off + i) : 0;
			__syncwarp();
		}
		const u32 nwin = L1 - h - 1;                                         // j = 1 .. L1-h-1 (:534)
		u32 qn = 0;        // queued candidates (warp-uniform)
		u32 en = 0;        // buffered edges (warp-uniform); keeps counting past OGB_EC
		c_probes += nwin;

		// ---- phase B: each lane verifies one queued candidate (taken from the tail of the queue)
		auto verify_batch = [&]() {
			u32 take = qn < 32 ? qn : 32;
			u32 base = qn - take;
			u64 e0 = 0, e1 = 0; int ne = 0;
			if (lane < take) {
				u64 c = hq[base + lane];
				ne = verify_candidate<MODE, LdShared>(A, s, qi, L1, (u32)(c >> 32), (u32)c, e0, e1);
				c_cand++;
			}
			qn = base;
			if (MODE == MODE_OVERLAP) {
				u32 b0 = __ballot_sync(0xFFFFFFFFu, ne >= 1), b1 = __ballot_sync(0xFFFFFFFFu, ne >= 2);
				u32 p0 = en + __popc(b0 & ((1u << lane) - 1));
				if (ne >= 1 && p0 < OGB_EC) eb[p0] = e0;
				en += __popc(b0);
				u32 p1 = en + __popc(b1 & ((1u << lane) - 1));
				if (ne >= 2 && p1 < OGB_EC) eb[p1] = e1;
				en += __popc(b1);
			} else {
				c_hits += ne;
			}
			__syncwarp();
		};

		for (u32 jb = 1; jb <= nwin; jb += 32) {
			// ---- phase A: one window per lane
			u32 j = jb + lane;
			bool active = j <= nwin;
			u64 hash = 0; u32 b = 0;
			if (active) { hash = key_hash<LdShared>(s, j, h); b = bucket_of(hash, A.T.nb); }
			const u32 fp = hash_fp(hash);
			while (__any_sync(0xFFFFFFFFu, active)) {
				u64 sl[OGB_SLOTS];
				u32 mm = 0;                                                  // slots of this lane's bucket whose fingerprint matches
				if (active) {
					load_bucket(A.T.slots, b, sl);
					c_sectors++;
					#pragma unroll
					for (int k = 0; k < OGB_SLOTS; k++) mm |= ((u32)(sl[k] >> 32) == fp) << k;
					// K1 fills a bucket front to back, so "last slot taken" = full = the key may continue in the next bucket
					active = sl[OGB_SLOTS - 1] != 0;
					if (active) b = (b + 1 == A.T.nb) ? 0 : b + 1;
				}
				// queue the matches: one ballot per round; a second match in the same bucket is rare
				u32 bal;
				while ((bal = __ballot_sync(0xFFFFFFFFu, mm != 0)) != 0) {
					if (mm) {
						int k = __ffs(mm) - 1;
						u64 v = sl[0];
						#pragma unroll
						for (int q = 1; q < OGB_SLOTS; q++) v = (k == q) ? sl[q] : v;
						hq[qn + __popc(bal & ((1u << lane) - 1))] = ((u64)j << 32) | (u32)v;
						mm &= mm - 1;
					}
					qn += __popc(bal);
				}
				__syncwarp();
				while (qn >= 32) verify_batch();                             // keeps room for 32 lanes x 8 slots
			}
		}
		while (qn > 0) verify_batch();

		if (MODE == MODE_OVERLAP) {
			// ---- phase C
			if (en > c_maxdeg) c_maxdeg = en;
			if (en == 0) { if (lane == 0) A.nodes[qi] = 0; }
			else if (en <= OGB_EC) {
				u64 mine = ~0ull;
				if (en <= 32) { if (lane < en) mine = eb[lane]; mine = warp_sort32(mine, lane); }
				else warp_sort(eb, en, lane);
				u64 start = 0;
				if (lane == 0) start = atomicAdd(A.ctr + CTR_EDGE_CURSOR, (u64)en);
				start = __shfl_sync(0xFFFFFFFFu, start, 0);
				if (start + en <= A.edge_cap) {
					if (en <= 32) { if (lane < en) A.edges[start + lane] = mine; }
					else for (u32 t = lane; t < en; t += 32) A.edges[start + t] = eb[t];
				} else if (lane == 0) atomicAdd(A.ctr + CTR_EDGES_DROPPED, (u64)en);
				if (lane == 0) A.nodes[qi] = (start << OGB_DEG_BITS) | en;
				__syncwarp();
			} else if (lane == 0) {
				u64 pos = atomicAdd(A.ctr + CTR_OVERFLOW, 1ull);
				if (pos < A.overflow_cap) A.overflow_list[pos] = qi;
				A.nodes[qi] = OGB_NODE_OVERFLOW;
			}
		}
	}
	// per-warp counters -> global (one atomic per warp per counter)
	for (int d = 16; d > 0; d >>= 1) {
		c_sectors += __shfl_down_sync(0xFFFFFFFFu, c_sectors, d);
		c_cand += __shfl_down_sync(0xFFFFFFFFu, c_cand, d);
		c_hits += __shfl_down_sync(0xFFFFFFFFu, c_hits, d);
	}
	if (lane == 0) {
		atomicAdd(A.ctr + CTR_PROBES, c_probes);
		atomicAdd(A.ctr + CTR_SECTORS, c_sectors);
		atomicAdd(A.ctr + CTR_CANDIDATES, c_cand);
		if (MODE == MODE_CONTAIN) atomicAdd(A.ctr + CTR_CONTAIN_HITS, c_hits);
		else atomicMax(A.ctr + CTR_MAX_DEGREE, (u64)c_maxdeg);
	}
}

// Slow path for reads with more than OGB_EC edges (repeats): one block per queued read, two
// passes over its windows. Pass 0 counts, then a global range is claimed, pass 1 stores, and the
// block sorts the range in global memory (bitonic). Simple on purpose: it only ever sees a handful
// of reads.
__global__ void __launch_bounds__(256) k_scan_big(ScanArgs A, u32 n_over)
{
	__shared__ u32 s_count;
	__shared__ u64 s_start;
	const u32 h = A.T.h;
	for (u32 it = blockIdx.x; it < n_over; it += gridDim.x) {
		u32 qi = A.overflow_list[it];
		u64 off; u32 L1;
		read_geom(A.R, qi, off, L1);
		const u64 *s = A.R.words + off;
		const u32 nwin = L1 - h - 1;
		for (int pass = 0; pass < 2; pass++) {
			if (threadIdx.x == 0) s_count = 0;
			__syncthreads();
			for (u32 j = 1 + threadIdx.x; j <= nwin; j += blockDim.x) {
				u64 hash = key_hash<LdGlobal>(s, j, h);
				u32 fp = hash_fp(hash), b = bucket_of(hash, A.T.nb);
				for (;;) {
					u64 sl[OGB_SLOTS];
					load_bucket(A.T.slots, b, sl);
					bool full = true;
					for (int k = 0; k < OGB_SLOTS; k++) {
						if (sl[k] == 0) { full = false; continue; }
						if ((u32)(sl[k] >> 32) != fp) continue;
						u64 e0 = 0, e1 = 0;
						int ne = verify_candidate<MODE_OVERLAP, LdGlobal>(A, s, qi, L1, j, (u32)sl[k], e0, e1);
						if (ne) {
							u32 pos = atomicAdd(&s_count, (u32)ne);
							if (pass == 1 && s_start + pos + ne <= A.edge_cap) {
								A.edges[s_start + pos] = e0;
								if (ne == 2) A.edges[s_start + pos + 1] = e1;
							}
						}
					}
					if (!full) break;
					b = (b + 1 == A.T.nb) ? 0 : b + 1;
				}
			}
			__syncthreads();
			if (pass == 0) {
				if (threadIdx.x == 0) {
					u32 m = 1; while (m < s_count) m <<= 1;                  // padded to a power of two for the sort
					s_start = atomicAdd(A.ctr + CTR_EDGE_CURSOR, (u64)m);
					atomicAdd(A.ctr + CTR_PAD, (u64)(m - s_count));          // padding is not an edge
					atomicMax(A.ctr + CTR_MAX_DEGREE, (u64)s_count);
				}
				__syncthreads();
			}
		}
		u32 n = s_count, m = 1;
		while (m < n) m <<= 1;
		u64 start = s_start;
		if (start + m <= A.edge_cap) {
			u64 *buf = A.edges + start;
			for (u32 i = n + threadIdx.x; i < m; i += blockDim.x) buf[i] = ~0ull;
			__syncthreads();
			for (u32 k = 2; k <= m; k <<= 1)
				for (u32 jj = k >> 1; jj > 0; jj >>= 1) {
					for (u32 t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
						u32 i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1)), l = i | jj;
						bool up = (i & k) == 0;
						u64 x = buf[i], y = buf[l];
						if ((x > y) == up) { buf[i] = y; buf[l] = x; }
					}
					__syncthreads();
				}
		} else if (threadIdx.x == 0) atomicAdd(A.ctr + CTR_EDGES_DROPPED, (u64)n);
		if (threadIdx.x == 0) A.nodes[qi] = (start << OGB_DEG_BITS) | n;
		__syncthreads();
	}
}

// superReadID decode + contained bitmap: one thread per read.
__global__ void k_contained_bitmap(const u64 *__restrict__ sup, u32 n, u32 *__restrict__ bitmap, u64 *ctr)
{
	u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
	bool c = idx < n && sup[idx] != 0;
	u32 bal = __ballot_sync(0xFFFFFFFFu, c);
	if ((threadIdx.x & 31) == 0 && idx < n) {
		bitmap[idx >> 5] = bal;
		if (bal) atomicAdd(ctr + CTR_N_CONTAINED, (u64)__popc(bal));
	}
}

// ------------------------------------------------------------------------------------------------
// K5: transitive-edge marking, one warp per node (OverlapGraph.cpp:574-615).
// The neighbour set (destination node -> INPLAY/ELIMINATED) is an open-addressing set in shared
// memory (degree <= 256) or in a global scratch pool (larger). Pivots are walked sequentially in
// adjacency order; the adjacency of an in-play pivot is scanned by all lanes.
// ------------------------------------------------------------------------------------------------
struct MarkArgs {
	const u64 *nodes;
	const u64 *edges;
	unsigned char *eflag;       // per edge: 1 = eliminated in the marking of its own node
	u32 lo, hi;                 // node indices [lo, hi) handled by this rank
	u32 *scratch_keys;          // global pool for big nodes
	unsigned char *scratch_state;
	u64 scratch_cap;
	u64 *ctr;
};

__device__ __forceinline__ u32 set_hash(u32 key, u32 capmask) { return (key * 2654435761u) >> 7 & capmask; }

__device__ __forceinline__ u32 set_insert(u32 *keys, u32 capmask, u32 key)
{
	u32 s = set_hash(key, capmask);
	for (;;) {
		u32 cur = atomicCAS(keys + s, 0u, key);
		if (cur == 0 || cur == key) return s;
		s = (s + 1) & capmask;
	}
}
__device__ __forceinline__ int set_find(const u32 *keys, u32 capmask, u32 key)
{
	u32 s = set_hash(key, capmask);
	for (;;) {
		u32 cur = keys[s];
		if (cur == key) return (int)s;
		if (cur == 0) return -1;
		s = (s + 1) & capmask;
	}
}

__global__ void __launch_bounds__(OGB_WARPS * 32, 6) k_mark(MarkArgs A)
{
	__shared__ u32 s_keys[OGB_WARPS][OGB_SETCAP];
	__shared__ unsigned char s_state[OGB_WARPS][OGB_SETCAP];
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gw = blockIdx.x * OGB_WARPS + wib, nwarps = gridDim.x * OGB_WARPS;
	u64 c_entries = 0, c_pivots = 0;

	for (u32 u = A.lo + gw; u < A.hi; u += nwarps) {
		u64 nd = __ldg(A.nodes + u);
		u32 deg = (u32)(nd & OGB_DEG_MASK);
		if (deg == 0) continue;
		const u64 start = nd >> OGB_DEG_BITS;
		u32 cap = 64;
		while (cap < 2 * deg) cap <<= 1;
		u32 *keys; unsigned char *st;
		if (cap <= OGB_SETCAP) { keys = s_keys[wib]; st = s_state[wib]; }
		else {
			u64 base = 0;
			if (lane == 0) base = atomicAdd(A.ctr + CTR_SCRATCH_CURSOR, (u64)cap);
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if (base + cap > A.scratch_cap) { if (lane == 0) atomicAdd(A.ctr + CTR_SCRATCH_FAIL, 1ull); continue; }
			keys = A.scratch_keys + base; st = A.scratch_state + base;
		}
		const u32 capmask = cap - 1;
		for (u32 i = lane; i < cap; i += 32) keys[i] = 0;
		__syncwarp();
		// mark all neighbours INPLAY (:577-578); a lane keeps the set slot of its own edge
		u64 e = 0; int sk = -1;
		if (deg <= 32) {
			if (lane < deg) { e = __ldg(A.edges + start + lane); sk = (int)set_insert(keys, capmask, edge_dst(e)); st[sk] = 1; }
		} else {
			for (u32 k = lane; k < deg; k += 32) st[set_insert(keys, capmask, edge_dst(__ldg(A.edges + start + k)))] = 1;
		}
		__syncwarp();
		// Pivots in adjacency (offset) order (:580-600). The reference walks every edge and skips
		// the ones whose destination is no longer INPLAY (:583); states only ever go INPLAY ->
		// ELIMINATED, so "the next pivot" is simply the lowest-index edge after the current one whose
		// destination is INPLAY now: one ballot + ffs per ACTIVE pivot (~2 per node) instead of a
		// dependent shared-memory round trip per edge.
		for (u32 cb = 0; cb < deg; cb += 32) {
			const u32 k = cb + lane;
			if (deg > 32) {
				e = 0; sk = -1;
				if (k < deg) { e = __ldg(A.edges + start + k); sk = set_find(keys, capmask, edge_dst(e)); }
			}
			int cur = -1;
			for (;;) {
				u32 m = __ballot_sync(0xFFFFFFFFu, k < deg && (int)lane > cur && st[sk] == 1);
				if (m == 0) break;
				cur = __ffs(m) - 1;
				u64 ei = __shfl_sync(0xFFFFFFFFu, e, cur);
				u32 v = edge_dst(ei), t1 = edge_orient(ei);
				u64 ndv = __ldg(A.nodes + (v - 1));
				u32 degv = (u32)(ndv & OGB_DEG_MASK);
				u64 startv = ndv >> OGB_DEG_BITS;
				c_pivots++; c_entries += degv;
				for (u32 kk = lane; kk < degv; kk += 32) {
					u64 f = __ldg(A.edges + startv + kk);
					if (compatible(t1, edge_orient(f))) {
						int sw = set_find(keys, capmask, edge_dst(f));
						if (sw >= 0 && st[sw] == 1) st[sw] = 2;              // :588-596
					}
				}
				__syncwarp();
			}
		}
		// flag own edges to eliminated nodes (:601-607; the twin half is applied in k_twin_keep)
		if (deg <= 32) { if (lane < deg) A.eflag[start + lane] = st[sk] == 2; }
		else for (u32 k = lane; k < deg; k += 32) A.eflag[start + k] = st[set_find(keys, capmask, edge_dst(__ldg(A.edges + start + k)))] == 2;
		__syncwarp();
	}
	if (lane == 0) { atomicAdd(A.ctr + CTR_PIVOT_ENTRIES, c_entries); atomicAdd(A.ctr + CTR_ACTIVE_PIVOTS, c_pivots); }
}

// ------------------------------------------------------------------------------------------------
// K6: an edge (u,w) survives iff it was not flagged by u's marking and its twin was not flagged by
// w's marking (:605-606, :623-661). Marks are per destination NODE, so w's verdict on u is read off
// any (w,u) entry of w's adjacency -- no twin pointers are needed. One warp per node: the few edges
// that u itself kept (~2) are checked one after another, each by a warp-wide scan of w's adjacency;
// survivors are written compacted to surv[start .. start+cnt) as final records, so that the copy
// kernel after the scan touches survivors only.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OGB_WARPS * 32) k_twin_keep(const u64 *__restrict__ nodes, const u64 *__restrict__ edges,
                                                              const unsigned char *__restrict__ eflag, ogb_edge *__restrict__ surv,
                                                              u32 *__restrict__ cnt, u32 lo, u32 hi, u64 *ctr)
{
	const u32 lane = threadIdx.x & 31;
	const u32 gw = blockIdx.x * OGB_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * OGB_WARPS;
	u32 c_nodes = 0, c_asym = 0;
	for (u32 u = lo + gw; u < hi; u += nwarps) {
		u64 nd = __ldg(nodes + u);
		u32 deg = (u32)(nd & OGB_DEG_MASK);
		u64 start = nd >> OGB_DEG_BITS;
		u32 total = 0;
		for (u32 kb = 0; kb < deg; kb += 32) {
			u32 k = kb + lane;
			u64 e = 0, ndw = 0;
			bool mine = k < deg && !eflag[start + k];
			if (mine) { e = __ldg(edges + start + k); ndw = __ldg(nodes + (edge_dst(e) - 1)); }   // all twin nodes fetched together
			u32 todo = __ballot_sync(0xFFFFFFFFu, mine);
			while (todo) {
				int src = __ffs(todo) - 1;
				todo &= todo - 1;
				u64 ee = __shfl_sync(0xFFFFFFFFu, e, src), nw = __shfl_sync(0xFFFFFFFFu, ndw, src);
				u32 degw = (u32)(nw & OGB_DEG_MASK);
				u64 startw = nw >> OGB_DEG_BITS;
				int verdict = -1;                                            // -1 not found, 0 keep, 1 twin flagged
				for (u32 xb = 0; xb < degw && verdict < 0; xb += 32) {
					u32 x = xb + lane;
					bool hit = x < degw && edge_dst(__ldg(edges + startw + x)) == u + 1;
					u32 hm = __ballot_sync(0xFFFFFFFFu, hit);
					if (hm) {
						int hl = __ffs(hm) - 1;
						int f = 0;
						if ((int)lane == hl) f = eflag[startw + x];
						verdict = __shfl_sync(0xFFFFFFFFu, f, hl);
					}
				}
				if (verdict < 0) { c_asym++; verdict = 0; }
				if (verdict == 0) {
					if (lane == 0) {
						ogb_edge r;
						r.src = u + 1; r.dst = edge_dst(ee); r.offset = (uint16_t)edge_offset(ee); r.orient = (uint8_t)edge_orient(ee); r.reserved = 0;
						surv[start + total] = r;
					}
					total++;
				}
			}
		}
		if (lane == 0) { cnt[u] = total; c_nodes += total > 0; }
	}
	if (lane == 0 && c_nodes) atomicAdd(ctr + CTR_NODES_FINAL, (u64)c_nodes);
	if (lane == 0 && c_asym) atomicAdd(ctr + CTR_ASYMMETRIC, (u64)c_asym);
}

// Exclusive scan of u32 counts into u64 offsets: (1) per-block sums, (2) one block scans the sums,
// (3) per-block scan + base.
#define OGB_SCAN_ITEMS 2048     // per block of 256 threads (8 per thread)
__global__ void __launch_bounds__(256) k_scan_sums(const u32 *__restrict__ cnt, u32 n, u64 *__restrict__ sums)
{
	__shared__ u64 sh[8];
	u64 base = (u64)blockIdx.x * OGB_SCAN_ITEMS, acc = 0;
	for (u32 i = threadIdx.x; i < OGB_SCAN_ITEMS; i += 256) if (base + i < n) acc += cnt[base + i];
	for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, d);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) { u64 t = 0; for (int i = 0; i < 8; i++) t += sh[i]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_top(u64 *__restrict__ sums, u32 nblocks, u64 *__restrict__ total)
{
	__shared__ u64 sh[1024];
	__shared__ u64 carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (u32 base = 0; base < nblocks; base += 1024) {
		u32 i = base + threadIdx.x;
		u64 v = i < nblocks ? sums[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (u32 d = 1; d < 1024; d <<= 1) {
			u64 x = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
			__syncthreads();
			sh[threadIdx.x] += x;
			__syncthreads();
		}
		if (i < nblocks) sums[i] = carry + sh[threadIdx.x] - v;              // exclusive
		__syncthreads();
		if (threadIdx.x == 1023) carry += sh[1023];
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_apply(const u32 *__restrict__ cnt, u32 n, const u64 *__restrict__ sums, u64 *__restrict__ out)
{
	__shared__ u64 sh[256];
	u64 base = (u64)blockIdx.x * OGB_SCAN_ITEMS;
	u32 v[8]; u64 local = 0;
	for (int i = 0; i < 8; i++) { u64 p = base + threadIdx.x * 8 + i; v[i] = p < n ? cnt[p] : 0; local += v[i]; }
	sh[threadIdx.x] = local;
	__syncthreads();
	for (u32 d = 1; d < 256; d <<= 1) {
		u64 x = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
		__syncthreads();
		sh[threadIdx.x] += x;
		__syncthreads();
	}
	u64 run = sums[blockIdx.x] + sh[threadIdx.x] - local;
	for (int i = 0; i < 8; i++) { u64 p = base + threadIdx.x * 8 + i; if (p < n) out[p] = run; run += v[i]; }
}

// Final edge records, sorted by (src, offset, dst, orient) = node order x adjacency order: one
// thread per node copies its cnt[u] survivors from surv[start..] to out[pos[u]..].
__global__ void __launch_bounds__(256) k_compact(const u64 *__restrict__ nodes, const ogb_edge *__restrict__ surv, const u32 *__restrict__ cnt,
                                                 const u64 *__restrict__ pos, ogb_edge *__restrict__ out, u32 lo, u32 hi)
{
	u32 u = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= hi) return;
	u32 c = cnt[u];
	if (c == 0) return;
	u64 start = __ldg(nodes + u) >> OGB_DEG_BITS, p = pos[u];
	for (u32 i = 0; i < c; i++) out[p + i] = surv[start + i];
}

// Pre-reduction edges as records (tests / keep_pre): one warp per node, position = start.
__global__ void __launch_bounds__(OGB_WARPS * 32) k_export_pre(const u64 *__restrict__ nodes, const u64 *__restrict__ edges,
                                                               const u64 *__restrict__ pos, ogb_edge *__restrict__ out, u32 n)
{
	const u32 lane = threadIdx.x & 31;
	const u32 gw = blockIdx.x * OGB_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * OGB_WARPS;
	for (u32 u = gw; u < n; u += nwarps) {
		u64 nd = __ldg(nodes + u);
		u32 deg = (u32)(nd & OGB_DEG_MASK);
		u64 start = nd >> OGB_DEG_BITS, p = pos[u];
		for (u32 k = lane; k < deg; k += 32) {
			u64 e = __ldg(edges + start + k);
			ogb_edge r;
			r.src = u + 1; r.dst = edge_dst(e); r.offset = (uint16_t)edge_offset(e); r.orient = (uint8_t)edge_orient(e); r.reserved = 0;
			out[p + k] = r;
		}
	}
}
__global__ void k_degrees(const u64 *__restrict__ nodes, u32 n, u32 *__restrict__ cnt)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) cnt[i] = (u32)(nodes[i] & OGB_DEG_MASK);
}

// ------------------------------------------------------------------------------------------------
// HashTable::getListOfReads (HashTable.cpp:202-221) for a batch of packed keys: one thread per key.
// pass 0 counts matches into cnt[k]; pass 1 writes id | o<<62 at out[pos[k]..]. Matches are
// verified exactly against the stored read's prefix/suffix.
// ------------------------------------------------------------------------------------------------
__global__ void k_lookup(ReadStore R, Table T, const u64 *__restrict__ keys, u32 kw, u64 n_keys, u32 *__restrict__ cnt,
                         const u64 *__restrict__ pos, u64 *__restrict__ out, int pass)
{
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_keys) return;
	const u64 *key = keys + k * (kw + 2);
	u64 hash = key_hash<LdGlobal>(key, 0, T.h);
	u32 fp = hash_fp(hash), b = bucket_of(hash, T.nb), c = 0;
	for (;;) {
		u64 sl[OGB_SLOTS];
		load_bucket(T.slots, b, sl);
		bool full = true;
		for (int s = 0; s < OGB_SLOTS; s++) {
			if (sl[s] == 0) { full = false; continue; }
			if ((u32)(sl[s] >> 32) != fp) continue;
			u32 val = (u32)sl[s], ri = (val >> 2) - 1, o = val & 3;
			u64 off; u32 L;
			read_geom(R, ri, off, L);
			const u64 *t = R.words + off + (o >> 1) * padded_words(L);
			if (!region_equal<LdGlobal, LdGlobal>(key, 0, t, (o & 1) ? L - T.h : 0, T.h)) continue;
			if (pass == 1) out[pos[k] + c] = (u64)(ri + 1) | ((u64)o << 62);
			c++;
		}
		if (!full) break;
		b = (b + 1 == T.nb) ? 0 : b + 1;
	}
	if (pass == 0) cnt[k] = c;
}
