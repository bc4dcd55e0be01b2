// ogb_kernels.cuh -- hand-written sm_100a kernels of the overlap-graph build.
//
//   K0  k_pack_ascii / k_pack_words   Read::setRead + reverseComplement          (Read.cpp:75-127)
//   K1  k_hash_insert                 HashTable::hashRead + insertIntoTable       (HashTable.cpp:88-195)
//   K2  k_probe/k_verify<CONTAIN>     markContainedReads + checkOverlapForContainedRead (OverlapGraph.cpp:225-340)
//   K3  k_probe/k_verify<OVERLAP>     insertAllEdgesOfRead + checkOverlap (OverlapGraph.cpp:354-383,529-565)
//   K4  (none)                        the per-node sort (OverlapGraph.cpp:563) is replaced by min-reductions in K5 / ranking in K6
//   K5  k_mark                        markTransitiveEdges                         (OverlapGraph.cpp:574-615)
//   K6  k_keep / k_emit               removeTransitiveEdges                       (OverlapGraph.cpp:623-661)
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
//   index         nb buckets of 64 bytes (one HBM burst: measured on B200, a random 32-byte sector read
//                 costs a 64-byte fetch anyway, profiles/exp_r1_l2fetch.txt). A bucket is 16 u32 words:
//                 words 0..4 hold ten 16-bit key fingerprints (slot i in half i&1 of word i>>1), words
//                 5..14 the ten values id<<2|o (0 = empty), word 15 is unused. One slot per (key,value);
//                 a key's entries sit in its home bucket and, when that is full, in the following
//                 buckets (linear probing by bucket). Only a fingerprint of the key is stored: every
//                 consumer verifies the complete overlap (window included) against the packed reads,
//                 which makes the result exact and independent of hash values.
//                 4N entries in N buckets = load 0.4: the ten-slot compare is five SIMD instructions
//                 and fewer than 1 % of the buckets spill.
//   edges         u64 = offset<<48 | dst<<16 | orient<<14 | twin<<2 | flags, so integer order = (offset,dst,orient);
//                 they live in per-read slot regions (see GraphView) from discovery to the final list.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define OGB_SLOTS 10            // slots per bucket (64 bytes = one HBM burst)
#define OGB_BWORDS 16           // u32 words per bucket
#define OGB_WARPS 8             // warps per block in the scan / mark kernels
// Resident blocks per SM the probe / verify kernels are compiled for (register cap). Both are latency-bound:
// 6 blocks (40 registers, no spills) measured 5-15 % faster than 5 (48); 7-8 (32 registers, spills) slower
// (profiles/r1_notes.md).
#ifndef OGB_PROBE_MINBLOCKS
#define OGB_PROBE_MINBLOCKS 6
#endif
#ifndef OGB_VERIFY_MINBLOCKS
#define OGB_VERIFY_MINBLOCKS 6
#endif
#define OGB_SETCAP 512          // per-warp neighbour set slots in shared memory (degree <= 256)

enum { MODE_OVERLAP = 0, MODE_CONTAIN = 1 };

// device counters (u64 each)
enum {
	CTR_EDGE_CURSOR = 0, CTR_OVERFLOW, CTR_PROBES, CTR_SECTORS, CTR_CANDIDATES, CTR_CONTAIN_HITS,
	CTR_PIVOT_ENTRIES, CTR_ACTIVE_PIVOTS, CTR_MAX_DEGREE, CTR_N_CONTAINED, CTR_NODES_FINAL,
	CTR_ASYMMETRIC, CTR_SCRATCH_CURSOR, CTR_SCRATCH_FAIL, CTR_CAND_MAX, CTR_BIG_NODES, CTR_EXT_CURSOR, CTR_PQ_OVERFLOW,
	CTR_TABLE_FULL, CTR_MORE_CURSOR, CTR_MORE_NEED, CTR_BIGREC_CURSOR, CTR_HROW_CURSOR, CTR_COUNT
};

struct ReadStore {
	const u64 *words;
	const u64 *meta;     // null when uniform
	u32 n;
	u32 uniform_len;     // 0 when lengths differ
	u32 uniform_pw;      // padded words per strand when uniform
};

struct Table {
	u32 *slots;          // nb * OGB_BWORDS words
	u32 *summary;        // nb words: Bloom bits of the fingerprints stored in the bucket | OGB_SPILLED; null = not used
	u32 nb;              // buckets = nparts * part_buckets
	u32 nparts;          // hash partitions = ranks x sub: rank r builds partitions [r*sub, (r+1)*sub), then the slices are allgathered;
	                     // a partition is also the unit the probe of a big index works through at a time (L2-sized)
	u32 part_buckets;    // buckets per partition; linear probing wraps inside a partition
	u32 sub;             // partitions per rank
	u32 my_rank;         // K1: insert only the keys of this rank's partitions
	u32 h;               // hashStringLength = minOverlap-1 (HashTable.cpp:54)
	u64 *ctr;            // device counters (K1 raises CTR_TABLE_FULL)
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
__device__ __forceinline__ u32 ld_na32(const u32 *p)
{
	u32 v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
struct LdShared { static __device__ __forceinline__ u64 ld(const u64 *p) { return *p; } };
struct LdGlobal { static __device__ __forceinline__ u64 ld(const u64 *p) { return __ldg(p); } };
struct LdStream { static __device__ __forceinline__ u64 ld(const u64 *p) { return ld_na(p); } };

// Upper 64 bits of (a:b) << sh for sh in [0,63]: two 32-bit funnel shifts (SHF) after picking the
// three source words, instead of the 64-bit shift / or / select sequence the compiler emits.
__device__ __forceinline__ u64 funnel(u64 a, u64 b, u32 sh)
{
	const u32 a1 = (u32)(a >> 32), a0 = (u32)a, b1 = (u32)(b >> 32), b0 = (u32)b;
	const bool hi = sh >= 32;
	const u32 x2 = hi ? a0 : a1, x1 = hi ? b1 : a0, x0 = hi ? b0 : b1;
	return ((u64)__funnelshift_l(x1, x2, sh) << 32) | __funnelshift_l(x0, x1, sh);
}

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

// Hash of the h bases starting at base p of a packed strand. lead = the key's first 16 bases (fewer when
// h < 16): the hash partition is a function of lead alone, so that the sharded table build can tell
// from one funnel shift whether a key belongs to this rank's partition (partition_of).
template <class LD> __device__ __forceinline__ u64 key_hash(const u64 *__restrict__ w, u32 p, u32 h, u32 &lead)
{
	const u32 wi = p >> 5, sh = (p & 31) << 1;
	const u64 a = LD::ld(w + wi), b = LD::ld(w + wi + 1);
	u64 k0 = funnel(a, b, sh);
	if (h <= 32) { k0 &= ~0ULL << (64 - 2 * h); lead = (u32)(k0 >> 32); return hash2(k0, 0); }
	lead = (u32)(k0 >> 32);
	const u64 c = LD::ld(w + wi + 2);
	u64 k1 = funnel(b, c, sh);
	if (h <= 64) return hash2(k0, k1 & (~0ULL << (128 - 2 * h)));
	u64 acc = hash2(k0, k1);
	u32 rem = h - 64;
	p += 64;
	for (; rem > 32; rem -= 32, p += 32) acc = mix64(acc, extract32<LD>(w, p));
	return mix64(acc, extract32<LD>(w, p) & (~0ULL << (64 - 2 * rem)));
}
// lead of the key alone (two loads, one funnel shift)
template <class LD> __device__ __forceinline__ u32 key_lead(const u64 *__restrict__ w, u32 p, u32 h)
{
	u64 k0 = extract32<LD>(w, p);
	if (h < 32) k0 &= ~0ULL << (64 - 2 * h);
	return (u32)(k0 >> 32);
}

__device__ __forceinline__ u32 hash_fp(u64 hash) { u32 f = (u32)hash & 0xFFFFu; return f ? f : 1u; }   // 16 bits; 0 is reserved for "empty"
// Per-bucket summary word (4 bytes per bucket, L2-resident up to tens of millions of reads): two
// Bloom bits per stored fingerprint in bits 0..30, bit 31 = "an insert found this bucket full and moved
// on". A probe whose two bits are not both set, in a bucket that never spilled, cannot match: it skips
// the 64-byte bucket fetch from HBM (about 57 % of the windows at 30x coverage, more at low coverage).
#define OGB_SPILLED 0x80000000u
__device__ __forceinline__ u32 summary_bits(u32 fp) { return (1u << min(fp & 31u, 30u)) | (1u << min((fp >> 5) & 31u, 30u)); }

// Partition = range reduction of a multiplicative hash of the key's lead; bucket inside the partition =
// range reduction of the high hash half. With one partition this is plain floor(x*nb / 2^32).
__device__ __forceinline__ u32 partition_of(u32 lead, const Table &T) { return T.nparts == 1 ? 0 : __umulhi(lead * 0x9E3779B1u, T.nparts); }
__device__ __forceinline__ u32 bucket_of(u64 hash, u32 lead, const Table &T, u32 &part)
{
	part = partition_of(lead, T);
	return part * T.part_buckets + __umulhi((u32)(hash >> 32), T.part_buckets);
}
// next bucket of the probe sequence: wraps at the end of the partition (pend = one past its last bucket)
__device__ __forceinline__ u32 next_bucket(u32 b, u32 pend, const Table &T) { return b + 1 == pend ? pend - T.part_buckets : b + 1; }

// s[a..a+len) == t[b..b+len) on packed strands, streaming one new word per side and 32 bases. No
// early exit: (almost) every candidate verifies, and independent iterations keep the sector
// requests of the partner read in flight together. Words past the last base of a region are never
// loaded: for a 100 bp partner that would cost a second 64-byte HBM burst per candidate.
template <class LDS, class LDT>
__device__ __forceinline__ bool region_equal(const u64 *__restrict__ s, u32 a, const u64 *__restrict__ t, u32 b, u32 len)
{
	const u32 sa = (a & 31) << 1, sb = (b & 31) << 1;
	const u64 *ws = s + (a >> 5), *wt = t + (b >> 5);
	const u64 *es = s + ((a + len - 1) >> 5), *et = t + ((b + len - 1) >> 5);   // last words that hold compared bases
	u64 s0 = LDS::ld(ws), t0 = LDT::ld(wt), diff = 0;
	u32 k = 0;
	for (; k + 32 <= len; k += 32) {
		++ws; ++wt;
		const u64 s1 = ws <= es ? LDS::ld(ws) : 0, t1 = wt <= et ? LDT::ld(wt) : 0;
		diff |= funnel(s0, s1, sa) ^ funnel(t0, t1, sb);
		s0 = s1; t0 = t1;
	}
	const u32 rem = len - k;
	if (rem) {
		++ws; ++wt;
		const u64 s1 = ws <= es ? LDS::ld(ws) : 0, t1 = wt <= et ? LDT::ld(wt) : 0;
		diff |= (funnel(s0, s1, sa) ^ funnel(t0, t1, sb)) & (~0ULL << (64 - 2 * rem));
	}
	return diff == 0;
}

// p[pa..pa+len) == q[0..len): the common shape of both overlap checks (one side always starts at a
// strand boundary), one funnel shift per 32 bases.
template <class LDP, class LDQ>
__device__ __forceinline__ bool region_equal_aligned(const u64 *__restrict__ p, u32 pa, const u64 *__restrict__ q, u32 len)
{
	const u32 sh = (pa & 31) << 1;
	const u64 *wp = p + (pa >> 5), *ep = p + ((pa + len - 1) >> 5);
	u64 p0 = LDP::ld(wp), diff = 0;
	u32 k = 0;
	for (; k + 32 <= len; k += 32) {
		++wp;
		const u64 p1 = wp <= ep ? LDP::ld(wp) : 0;
		diff |= funnel(p0, p1, sh) ^ LDQ::ld(q + (k >> 5));
		p0 = p1;
	}
	const u32 rem = len - k;
	if (rem) {
		++wp;
		const u64 p1 = wp <= ep ? LDP::ld(wp) : 0;
		diff |= (funnel(p0, p1, sh) ^ LDQ::ld(q + (k >> 5))) & (~0ULL << (64 - 2 * rem));
	}
	return diff == 0;
}

// p[pa..pa+len) == q[0..len) for strands of exactly NW words each where the p region runs to the end of
// its strand (every overlap check on equal-length reads has this shape): straight-line code, no trip
// count that differs between the lanes of a warp. Words past the strand are not loaded.
template <int NW>
__device__ __forceinline__ bool suffix_equals_prefix(const u64 *__restrict__ p, u32 pa, const u64 *__restrict__ q, u32 len)
{
	const u32 wi = pa >> 5, sh = (pa & 31) << 1;
	u64 x[NW + 1];
	#pragma unroll
	for (int i = 0; i <= NW; i++) x[i] = wi + i < NW ? __ldg(p + wi + i) : 0;
	u64 diff = 0;
	#pragma unroll
	for (int i = 0; i < NW; i++) {
		const int r = min(max((int)len - 32 * i, 0), 32);                    // bases of word i inside the region
		const u64 m = r ? ~0ULL << (64 - 2 * r) : 0;
		diff |= (funnel(x[i], x[i + 1], sh) ^ __ldg(q + i)) & m;
	}
	return diff == 0;
}

// Edge word: offset<<48 | dst<<16 | orient<<14 | twin<<2 | flags. Integer order of the upper 50 bits =
// (offset, dst, orient), the reference's sort key plus the deterministic tie-break of SURVEY.md App. B.1.
// twin (12 bits) = 1 + position of a (dst,src) entry in dst's list when K5 has seen it (0 = unknown),
// flags = OGB_ELIM | OGB_KEEP (see GraphView).
__device__ __forceinline__ u64 make_edge(u32 offset, u32 dst, u32 orient) { return ((u64)offset << 48) | ((u64)dst << 16) | ((u64)orient << 14); }
__device__ __forceinline__ u32 edge_dst(u64 e) { return (u32)(e >> 16); }
__device__ __forceinline__ u32 edge_orient(u64 e) { return (u32)(e >> 14) & 3; }
__device__ __forceinline__ u32 edge_offset(u64 e) { return (u32)(e >> 48); }
__device__ __forceinline__ u32 edge_twin(u64 e) { return (u32)(e >> 2) & 0xFFFu; }
__device__ __forceinline__ u64 edge_key(u64 e) { return e & ~0x3FFFull; }     // (offset, dst, orient) without twin / flag bits
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
	// lens / in_offsets are null for a uniform-length, tightly packed input (nothing but the words is uploaded)
	u32 L = lens ? lens[idx] : uniform_len;
	u64 off;
	if (uniform_len) off = (u64)idx * (2 * uniform_pw); else off = meta[idx] >> 16;
	u32 pw = padded_words(L), nw = (L + 31) >> 5;
	if (k >= pw) return;
	const u64 *src = in_words + (in_offsets ? in_offsets[idx] : (u64)idx * nw);
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
// Dataset stage on the device (Dataset.cpp:161-164, 197-202, 316-345):
// canonical strand, lexicographic sort, dedupe with frequencies, IDs = ranks. Reads are rows of W
// zero-padded words; padding is 'A' = 0, the smallest base, so row order followed by length is exactly
// std::string operator< (a proper prefix sorts first). The sort is an LSD radix sort over the words
// (cub::DeviceRadixSort per 64-bit word, driven from ogb_device.cu); these kernels do the rest.
// ------------------------------------------------------------------------------------------------

// One thread per kept read: pack forward and reverse complement word by word, keep the smaller row (:161-164).
__global__ void __launch_bounds__(128) k_ds_canon(const char *__restrict__ raw, const u64 *__restrict__ start, const unsigned short *__restrict__ len,
                                                  u64 *__restrict__ rows, u32 n, u32 W)
{
	const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n) return;
	const char *s = raw + start[g];
	const u32 L = len[g], nw = (L + 31) >> 5;
	u64 *row = rows + (u64)g * W;
	int pick = 0;                                                           // 0 undecided, 1 forward, 2 reverse complement
	for (u32 k = 0; k < nw; k++) {
		u64 fw = 0, rc = 0;
		for (u32 i = 0; i < 32; i++) {
			const u32 p = k * 32 + i;
			if (p < L) {
				u32 c = ((u32)s[p] >> 1) & 3; c ^= c >> 1;                      // A C G T -> 0 1 2 3
				u32 d = ((u32)s[L - 1 - p] >> 1) & 3; d ^= d >> 1;
				fw |= (u64)c << (62 - 2 * i);
				rc |= (u64)(3 - d) << (62 - 2 * i);
			}
		}
		if (pick == 0 && fw != rc) pick = fw < rc ? 1 : 2;
		row[k] = pick == 2 ? rc : fw;                                       // words before the first difference are equal on both strands
	}
	for (u32 k = nw; k < W; k++) row[k] = 0;
}

// Key of one LSD pass: word `word` of the row (word == W: the length), in the current order.
__global__ void __launch_bounds__(256) k_ds_key(const u64 *__restrict__ rows, const unsigned short *__restrict__ len, const u32 *__restrict__ perm,
                                                u64 *__restrict__ key, u32 n, u32 W, u32 word)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u32 g = perm[i];
	key[i] = word == W ? (u64)len[g] : rows[(u64)g * W + word];
}
__global__ void __launch_bounds__(256) k_ds_iota(u32 *__restrict__ perm, u32 n)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) perm[i] = i;
}

// Filter of Dataset::readDataset (Dataset.cpp:155-158, testRead :398-413) on the device: one thread per raw read. A read stays
// iff its length is > minOverlap (and < 65536), it consists of ACGT only (either case) and no base reaches (UINT64)(len * .8)
// occurrences. good[i] = 1/0; shortest / longest length of the good reads in stat[0] / stat[1].
__global__ void __launch_bounds__(128) k_ds_filter(const char *__restrict__ raw, const u64 *__restrict__ offs, u32 n_raw, u32 min_overlap, u32 *__restrict__ good, u64 *stat)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_raw) return;
	const u64 len = offs[i + 1] - offs[i];
	const char *s = raw + offs[i];
	bool ok = len > min_overlap && len < 65536;
	u32 cnt[4] = {0, 0, 0, 0};
	for (u32 k = 0; ok && k < (u32)len; k++) {
		const u32 ch = (u32)(unsigned char)s[k] & ~0x20u;                       // toupper for letters (:155-156)
		if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') ok = false;
		else cnt[(ch >> 1) & 3]++;
	}
	if (ok) {
		const u32 thr = (u32)((u32)len * .8);                                  // same double arithmetic as :409
		if (cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) ok = false;
	}
	good[i] = ok;
	if (ok) { atomicMin(stat, len); atomicMax(stat + 1, len); }
}
// start / length of the good reads, in input order (pos = exclusive scan of good)
__global__ void __launch_bounds__(256) k_ds_compact(const u32 *__restrict__ good, const u64 *__restrict__ pos, const u64 *__restrict__ offs, u32 n_raw,
                                                    u64 *__restrict__ start, unsigned short *__restrict__ len)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_raw || !good[i]) return;
	start[pos[i]] = offs[i];
	len[pos[i]] = (unsigned short)(offs[i + 1] - offs[i]);
}

// ---- LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass, stable; hand-written (round 1 used cub::DeviceRadixSort).
// A pass = k_rs_hist (per-block digit histograms, laid out digit-major so that one exclusive scan over the whole array yields
// every block's output offset per digit), the scan, k_rs_scatter. A block owns a tile of 2048 consecutive pairs and walks it
// in 8 rounds of 256; inside a round a warp ranks its 32 pairs with __match_any_sync, and a 64 x 256 table of (round, warp)
// counts in shared memory, prefix-summed per digit, orders the warps and rounds -- which keeps equal digits in input order.
// k_rs_orand finds the bits in which the keys differ at all, so that passes over constant bytes (zero padding, equal lengths) are skipped.
#define OGB_RS_TILE 2048
__global__ void __launch_bounds__(256) k_rs_orand(const u64 *__restrict__ key, u32 n, u64 *out)   // out[0] |= keys, out[1] &= keys
{
	u64 o = 0, a = ~0ull;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const u64 k = key[i]; o |= k; a &= k; }
	for (int d = 16; d > 0; d >>= 1) { o |= __shfl_down_sync(0xFFFFFFFFu, o, d); a &= __shfl_down_sync(0xFFFFFFFFu, a, d); }
	if ((threadIdx.x & 31) == 0) { atomicOr(out, o); atomicAnd(out + 1, a); }
}
__global__ void __launch_bounds__(256) k_rs_hist(const u64 *__restrict__ key, u32 n, u32 shift, u32 *__restrict__ hist, u32 nblk)
{
	__shared__ u32 h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const u32 base = blockIdx.x * OGB_RS_TILE;
	#pragma unroll
	for (int r = 0; r < 8; r++) { const u32 i = base + r * 256 + threadIdx.x; if (i < n) atomicAdd(&h[(u32)(key[i] >> shift) & 255u], 1u); }
	__syncthreads();
	hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}
__global__ void __launch_bounds__(256) k_rs_scatter(const u64 *__restrict__ key, const u32 *__restrict__ val, u64 *__restrict__ key_out, u32 *__restrict__ val_out,
                                                    u32 n, u32 shift, const u64 *__restrict__ offs, u32 nblk)
{
	__shared__ unsigned short cnt[64][256];                                  // pairs of digit d in (round, warp) p; then their exclusive prefix
	__shared__ u32 s_base[256];
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lt = (1u << lane) - 1;
	for (u32 i = threadIdx.x; i < 64 * 256 / 2; i += 256) ((u32 *)cnt)[i] = 0;
	s_base[threadIdx.x] = (u32)offs[threadIdx.x * nblk + blockIdx.x];
	__syncthreads();
	const u32 base = blockIdx.x * OGB_RS_TILE;
	u64 k[8]; u32 v[8]; unsigned char dg[8], rk[8];
	#pragma unroll
	for (int r = 0; r < 8; r++) {
		const u32 i = base + r * 256 + threadIdx.x;
		const bool valid = i < n;
		k[r] = valid ? key[i] : 0; v[r] = valid ? val[i] : 0;
		dg[r] = (unsigned char)((k[r] >> shift) & 255u); rk[r] = 0;
		const u32 act = __ballot_sync(0xFFFFFFFFu, valid);
		if (valid) {
			const u32 m = __match_any_sync(act, (u32)dg[r]);
			rk[r] = (unsigned char)__popc(m & lt);
			if ((m & lt) == 0) cnt[r * 8 + wib][dg[r]] = (unsigned short)__popc(m);   // the first lane of the group
		}
	}
	__syncthreads();
	{
		u32 run = 0;                                                          // thread d: exclusive prefix of digit d over the 64 (round, warp) pairs
		for (int p = 0; p < 64; p++) { const u32 c = cnt[p][threadIdx.x]; cnt[p][threadIdx.x] = (unsigned short)run; run += c; }
	}
	__syncthreads();
	#pragma unroll
	for (int r = 0; r < 8; r++) {
		const u32 i = base + r * 256 + threadIdx.x;
		if (i < n) { const u32 at = s_base[dg[r]] + cnt[r * 8 + wib][dg[r]] + rk[r]; key_out[at] = k[r]; val_out[at] = v[r]; }
	}
}

// head[i] = 1 when the read at sorted position i differs from its predecessor (:316-345).
__global__ void __launch_bounds__(256) k_ds_heads(const u64 *__restrict__ rows, const unsigned short *__restrict__ len, const u32 *__restrict__ perm,
                                                  u32 *__restrict__ head, u32 n, u32 W)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u32 h = 1;
	if (i > 0) {
		const u32 a = perm[i - 1], b = perm[i];
		h = len[a] != len[b];
		for (u32 k = 0; k < W && !h; k++) h = rows[(u64)a * W + k] != rows[(u64)b * W + k];
	}
	head[i] = h;
}
// Unique read u = number of heads before its first sorted position: source row, length, words needed, run start.
__global__ void __launch_bounds__(256) k_ds_unique(const u32 *__restrict__ head, const u64 *__restrict__ pos, const u32 *__restrict__ perm,
                                                   const unsigned short *__restrict__ len, u32 *__restrict__ usrc, unsigned short *__restrict__ ulen,
                                                   u32 *__restrict__ unw, u32 *__restrict__ ustart, u32 n)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !head[i]) return;
	const u64 u = pos[i];
	const u32 g = perm[i];
	usrc[u] = g; ulen[u] = len[g]; unw[u] = ((u32)len[g] + 31) >> 5; ustart[u] = i;
}
// Frequencies = run lengths; tight packed words of the unique reads (the host-side Dataset layout).
__global__ void __launch_bounds__(256) k_ds_emit(const u64 *__restrict__ rows, const u32 *__restrict__ usrc, const u32 *__restrict__ unw, const u32 *__restrict__ ustart,
                                                 const u64 *__restrict__ woff, u32 *__restrict__ freq, u64 *__restrict__ words, u32 nu, u32 n, u32 W)
{
	const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= nu) return;
	freq[u] = (u + 1 < nu ? ustart[u + 1] : n) - ustart[u];
	const u64 *row = rows + (u64)usrc[u] * W;
	u64 *dst = words + woff[u];
	for (u32 k = 0; k < unw[u]; k++) dst[k] = row[k];
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
	if (T.nparts > T.sub && partition_of(key_lead<LdGlobal>(w, p, T.h), T) / T.sub != T.my_rank) return;   // another rank builds that partition
	u32 lead;
	u64 hash = key_hash<LdGlobal>(w, p, T.h, lead);
	const u32 fp = hash_fp(hash), val = ((idx + 1) << 2) | o;
	u32 part;
	u32 b = bucket_of(hash, lead, T, part);
	const u32 pend = (part + 1) * T.part_buckets;
	// The probe sequence stays inside the key's hash partition, and the partition is a function of the key's first 16
	// bases alone: a skewed read set (one primer or repeat in front of most reads) can send more keys to a partition than
	// it has slots. After one lap the insert gives up and raises CTR_TABLE_FULL; ogb_hash_build retries with fewer, larger
	// partitions or a larger table (the reference's probe over the whole table, HashTable.cpp:163-195, cannot fill up).
	for (u32 steps = 0;; steps++) {
		if (steps >= T.part_buckets) { atomicMax(T.ctr + CTR_TABLE_FULL, 1ull); return; }
		u32 *w = T.slots + (u64)b * OGB_BWORDS;
		// one L2-coherent look at the ten values, then a CAS on the first empty one; buckets fill front
		// to back, so a lost race just moves on to the next slot. The fingerprint half-word is ORed in
		// after the value is claimed (readers run in later kernels).
		u32 cur[12];
		#pragma unroll
		for (int q = 0; q < 12; q += 4)
			asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(cur[q]), "=r"(cur[q + 1]), "=r"(cur[q + 2]), "=r"(cur[q + 3]) : "l"(w + 4 + q) : "memory");
		int s = OGB_SLOTS;                                                    // cur[1 + i] = value of slot i
		#pragma unroll
		for (int q = OGB_SLOTS - 1; q >= 0; q--) s = cur[1 + q] == 0 ? q : s;
		bool done = false;
		for (; s < OGB_SLOTS && !done; s++) done = atomicCAS(w + 5 + s, 0u, val) == 0;
		if (done) { atomicOr(w + ((s - 1) >> 1), fp << (16 * ((s - 1) & 1))); if (T.summary) atomicOr(T.summary + b, summary_bits(fp)); break; }
		if (T.summary) atomicOr(T.summary + b, OGB_SPILLED);
		b = next_bucket(b, pend, T);
	}
}

// K1 through partition queues (index beyond L2): random read-modify-writes into hundreds of MB run at the DRAM row /
// translation rate, so the keys are first hashed and SCATTERED into one queue per hash partition (k_key_part: a block
// counts its tile's keys per partition in shared memory, reserves queue space with one atomic per partition and tile,
// then writes bucket / fingerprint / value records), and the queues are inserted one partition at a time
// (k_insert_parts): while a partition is being filled its <= 48 MB of buckets are L2-resident.
struct KeyQueue {
	u32 *b, *f, *v;             // nparts regions of cap records: home bucket, fingerprint, value id<<2|o
	u64 *cursor;                // records appended per partition (may exceed cap: the build falls back to k_hash_insert)
	u64 cap;
	u32 nparts;                 // partitions of the index; this rank fills [first, first + count)
	u32 first, count;
};
#define OGB_KPT 4               // keys per thread and tile in k_key_part
__global__ void __launch_bounds__(256, 6) k_key_part(ReadStore R, Table T, KeyQueue Q)
{
	__shared__ u32 s_cnt[64];
	__shared__ u64 s_base[64];
	const u64 total = (u64)R.n * 4;
	const u32 tile_keys = 256 * OGB_KPT;
	const u64 tiles = (total + tile_keys - 1) / tile_keys;
	for (u64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		if (threadIdx.x < Q.nparts) s_cnt[threadIdx.x] = 0;
		__syncthreads();
		u32 eb[OGB_KPT], ef[OGB_KPT], ev[OGB_KPT], ep[OGB_KPT];
		#pragma unroll
		for (int r = 0; r < OGB_KPT; r++) {
			const u64 x = tile * tile_keys + r * 256 + threadIdx.x;
			ep[r] = 0xFFFFFFFFu; eb[r] = ef[r] = ev[r] = 0;
			if (x >= total) continue;
			const u32 idx = (u32)(x >> 2), o = (u32)(x & 3);
			u64 off; u32 L;
			read_geom(R, idx, off, L);
			const u64 *w = R.words + off + (o >> 1) * padded_words(L);
			const u32 p = (o & 1) ? L - T.h : 0;
			// several ranks: 1 - 1/G of the keys belong to partitions another rank builds -- they leave after one funnel shift
			if (T.nparts > T.sub && partition_of(key_lead<LdGlobal>(w, p, T.h), T) / T.sub != T.my_rank) continue;
			u32 lead;
			const u64 hash = key_hash<LdGlobal>(w, p, T.h, lead);
			u32 part;
			eb[r] = bucket_of(hash, lead, T, part);
			ef[r] = hash_fp(hash); ev[r] = ((idx + 1) << 2) | o;
			ep[r] = (part << 16) | atomicAdd(&s_cnt[part], 1u);
		}
		__syncthreads();
		if (threadIdx.x < Q.nparts) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(Q.cursor + threadIdx.x, (u64)s_cnt[threadIdx.x]) : 0;
		__syncthreads();
		#pragma unroll
		for (int r = 0; r < OGB_KPT; r++)
			if (ep[r] != 0xFFFFFFFFu) {
				const u32 part = ep[r] >> 16;
				const u64 at = s_base[part] + (ep[r] & 0xFFFFu);
				if (at < Q.cap) { const u64 i = (part - Q.first) * Q.cap + at; Q.b[i] = eb[r]; Q.f[i] = ef[r]; Q.v[i] = ev[r]; }
			}
		__syncthreads();
	}
}
// One insert: claims the first empty slot along the probe sequence (see k_hash_insert)
__device__ __forceinline__ void insert_record(const Table &T, u32 b, u32 fp, u32 val)
{
	const u32 part = b / T.part_buckets, pend = (part + 1) * T.part_buckets;
	for (u32 steps = 0;; steps++) {
		if (steps >= T.part_buckets) { atomicMax(T.ctr + CTR_TABLE_FULL, 1ull); return; }
		u32 *w = T.slots + (u64)b * OGB_BWORDS;
		u32 cur[12];
		#pragma unroll
		for (int q = 0; q < 12; q += 4)
			asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(cur[q]), "=r"(cur[q + 1]), "=r"(cur[q + 2]), "=r"(cur[q + 3]) : "l"(w + 4 + q) : "memory");
		int s = OGB_SLOTS;
		#pragma unroll
		for (int q = OGB_SLOTS - 1; q >= 0; q--) s = cur[1 + q] == 0 ? q : s;
		bool done = false;
		for (; s < OGB_SLOTS && !done; s++) done = atomicCAS(w + 5 + s, 0u, val) == 0;
		if (done) { atomicOr(w + ((s - 1) >> 1), fp << (16 * ((s - 1) & 1))); if (T.summary) atomicOr(T.summary + b, summary_bits(fp)); return; }
		if (T.summary) atomicOr(T.summary + b, OGB_SPILLED);
		b = next_bucket(b, pend, T);
	}
}
__global__ void __launch_bounds__(256, 6) k_insert_parts(Table T, KeyQueue Q)
{
	const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (u64)gridDim.x * blockDim.x;
	for (u32 part = Q.first; part < Q.first + Q.count; part++) {
		const u64 n = min(Q.cursor[part], Q.cap), base = (part - Q.first) * Q.cap;
		if (Q.cursor[part] > Q.cap && g == 0) atomicMax(T.ctr + CTR_TABLE_FULL, 2ull);   // a queue dropped keys (skewed leads): the host rebuilds with the direct kernel
		for (u64 i = g; i < n; i += nthreads) insert_record(T, Q.b[base + i], Q.f[base + i], Q.v[base + i]);
	}
}

// One bucket = 64 bytes = two 256-bit non-allocating loads. w[0..4] fingerprint pairs, w[5..14] values.
__device__ __forceinline__ void load_bucket(const u32 *__restrict__ slots, u32 b, u32 (&w)[OGB_BWORDS])
{
	const u32 *p = slots + (u64)b * OGB_BWORDS;
	u64 x[8];
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[0]), "=l"(x[1]), "=l"(x[2]), "=l"(x[3]) : "l"(p));
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[4]), "=l"(x[5]), "=l"(x[6]), "=l"(x[7]) : "l"(p + 8));
	#pragma unroll
	for (int i = 0; i < 8; i++) { w[2 * i] = (u32)x[i]; w[2 * i + 1] = (u32)(x[i] >> 32); }
}

// Ten-bit mask of the slots whose 16-bit fingerprint equals fp, by the zero-halfword trick on
// w ^ splat: z has bit 15 / 31 set where the low / high halfword is zero. The high bit can also come
// out set when the low half matched and the high half differs only in its lowest bit -- a spurious
// candidate (about 1 in 10^5 matches) that verification rejects like any fingerprint collision;
// consumers skip a spurious hit on an empty slot (value 0).
__device__ __forceinline__ u32 match_bucket(const u32 (&w)[OGB_BWORDS], u32 fp)
{
	const u32 splat = fp | (fp << 16);
	u32 z[5], any = 0;
	#pragma unroll
	for (int i = 0; i < 5; i++) { const u32 x = w[i] ^ splat; z[i] = (x - 0x00010001u) & ~x & 0x80008000u; any |= z[i]; }
	if (any == 0) return 0;
	u32 mm = 0;
	#pragma unroll
	for (int i = 0; i < 5; i++) mm |= (((z[i] >> 15) & 1) | ((z[i] >> 30) & 2)) << (2 * i);
	return mm;
}
// value of slot k (0..9)
__device__ __forceinline__ u32 bucket_value(const u32 (&w)[OGB_BWORDS], int k)
{
	u32 v = w[5];
	#pragma unroll
	for (int q = 1; q < OGB_SLOTS; q++) v = (k == q) ? w[5 + q] : v;
	return v;
}

// ------------------------------------------------------------------------------------------------
// K2 / K3: sliding-window scan, decomposed into thin data-parallel kernels (a fused warp-per-read
// kernel measured latency-bound at 31-50 % occupancy, profiles/r1_notes.md):
//
//   k_probe   warp per query read, lanes over its windows j = 1 .. L-h-1 (OverlapGraph.cpp:534):
//             key -> hash -> one 64-byte bucket (non-allocating 256-bit loads) -> fingerprint compare.
//             Matches are appended as candidates (read, j, slot value) to a global queue; a warp
//             reserves queue space in chunks of OGB_QCHUNK entries, so the queue cursor sees one
//             atomic per ~100 candidates, and pads its last chunk with sentinels.
//   k_verify  one THREAD per candidate: fetch the partner strand (one 64-byte burst for 100 bp) and
//             compare the whole overlap on packed words -- checkOverlap / checkOverlapForContainedRead.
//             Overlap mode appends the edge to the source read's slot region (deg[] atomics, spread
//             over N addresses); containment mode does the atomicMax on superRead.
//
// The host runs k_probe + k_verify over chunks of reads sized so that the candidate queue of a chunk
// stays L2-resident.
//
// Adjacency layout: read idx owns slots[(idx-lo)*cap .. +cap); a node with more than cap edges is
// moved, complete, to an extension area (GraphView below).
// ------------------------------------------------------------------------------------------------
#define OGB_QCHUNK 128          // candidate-queue entries a warp reserves at a time
#define OGB_NOCAND 0xFFFFFFFFu  // sentinel read index of a padding entry

struct ScanArgs {
	ReadStore R;
	Table T;
	u32 lo, hi;                 // query read indices [lo, hi) of this launch
	const u32 *contained;       // bitmap by read index (null when no read is contained)
	// candidate queue
	u32 *cand_q;                // read index (OGB_NOCAND = padding)
	u64 *cand_v;                // j<<32 | slot value (id<<2|o)
	u64 cand_cap;
	u64 *cand_cursor;           // entries reserved so far (may exceed cand_cap: entries beyond are dropped and counted)
	// outputs
	u64 *sup;                   // MODE_CONTAIN: per read idx, max over hits of (L_super<<32 | ~super_idx)
	u64 *slots_e;               // MODE_OVERLAP: slot regions, read idx owns [(idx-slot_lo)*cap, +cap)
	u32 slot_lo;
	u32 cap;                    // slots per read
	u32 *deg;                   // edges found per read, by read idx (keeps counting past cap)
	u32 *ov_q;                  // overflow edges of heavy nodes: read idx / edge
	u64 *ov_e;
	u64 ov_cap;
	u64 *ctr;
	u32 prefetch;               // probe: prefetch each candidate's partner strand to L2 (chunks whose candidates stay L2-resident)
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// A warp's position in the candidate queue: it owns [qbase, qbase+qsize) and has used qused of it.
struct QueueCursor { u64 qbase; u32 qused, qsize; };

__device__ __forceinline__ void queue_pad(const ScanArgs &A, const QueueCursor &Q, u32 lane)
{
	for (u32 i = Q.qused + lane; i < Q.qsize; i += 32) if (Q.qbase + i < A.cand_cap) A.cand_q[Q.qbase + i] = OGB_NOCAND;
}

// Appends the matches of all lanes (mm = this lane's 10-bit slot mask, w = its bucket) as candidates.
// Warp-synchronous: one shuffle scan of the per-lane counts, then every lane writes its own matches.
// Each candidate's partner strand (or, for mixed lengths, its geometry word) is prefetched to L2: the
// verify kernel that follows is bound by exactly that fetch.
__device__ __forceinline__ void append_matches(const ScanArgs &A, QueueCursor &Q, u32 lane, u32 mm, const u32 (&w)[OGB_BWORDS], u32 qi, u64 tag)
{
	const u32 cnt = __popc(mm);
	u32 inc = cnt;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (u32)d) inc += t; }
	const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
	if (Q.qused + total > Q.qsize) {                                         // pad what is left, reserve a new piece
		queue_pad(A, Q, lane);
		Q.qsize = total > OGB_QCHUNK ? (total + 31) & ~31u : OGB_QCHUNK;
		if (lane == 0) Q.qbase = atomicAdd(A.cand_cursor, (u64)Q.qsize);
		Q.qbase = __shfl_sync(0xFFFFFFFFu, Q.qbase, 0);
		Q.qused = 0;
	}
	u64 at = Q.qbase + Q.qused + (inc - cnt);
	for (; mm; mm &= mm - 1, at++) {
		const u32 v = bucket_value(w, __ffs(mm) - 1);
		if (at < A.cand_cap) { A.cand_q[at] = qi; A.cand_v[at] = tag | v; }
		if (A.prefetch) {
			const u32 ri = (v >> 2) - 1;
			if (A.R.uniform_len) prefetch_l2(A.R.words + (u64)ri * (2 * A.R.uniform_pw) + ((v >> 1) & 1) * A.R.uniform_pw);
			else prefetch_l2(A.R.meta + ri);
		}
	}
	Q.qused += total;
}

// Windows that passed the per-bucket summary wait in a small per-warp queue in shared memory and are
// taken out 32 at a time: the bucket fetch, the fingerprint compare and the candidate append -- two
// thirds of the kernel's instructions -- then run with every lane busy instead of with the ~45 % of the
// lanes whose window can have an entry at all (30x coverage; fewer at low coverage).
#define OGB_PENDQ 64
struct PendQueue {
	u32 b[OGB_WARPS][OGB_PENDQ];      // home bucket
	u32 f[OGB_WARPS][OGB_PENDQ];      // fingerprint | j << 16
	u32 q[OGB_WARPS][OGB_PENDQ];      // query read index
};

// Bucket fetch + fingerprint compare + candidate append for one window per lane (active lanes only).
__device__ __forceinline__ void probe_buckets(const ScanArgs &A, QueueCursor &Q, u32 lane, bool active, u32 b, u32 fj, u32 qi, u64 &c_sectors)
{
	const u32 fp = fj & 0xFFFFu;
	const u64 tag = (u64)(fj >> 16) << 32;
	u32 pend = 0, steps = 0;
	while (__any_sync(0xFFFFFFFFu, active)) {
		u32 w[OGB_BWORDS];
		u32 mm = 0;                                                          // slots of this lane's bucket whose fingerprint matches
		if (active) {
			load_bucket(A.T.slots, b, w);
			c_sectors++;
			mm = match_bucket(w, fp);
			// K1 fills a bucket front to back, so "last slot taken" = full = the key may continue in the next bucket
			// (at most one lap: a completely full partition never gets past ogb_hash_build)
			active = w[5 + OGB_SLOTS - 1] != 0 && ++steps < A.T.part_buckets;
			if (active) {
				if (pend == 0) pend = (b / A.T.part_buckets + 1) * A.T.part_buckets;
				b = next_bucket(b, pend, A.T);
			}
		}
		if (__any_sync(0xFFFFFFFFu, mm != 0)) append_matches(A, Q, lane, mm, w, qi, tag);
	}
}

// Pushes this lane's window if it passed the filter; when 32 or more are pending, the warp probes 32.
__device__ __forceinline__ void pend_push(const ScanArgs &A, PendQueue &P, u32 wib, u32 &plen, QueueCursor &Q, u32 lane, bool pass, u32 b, u32 fj, u32 qi, u64 &c_sectors)
{
	const u32 bal = __ballot_sync(0xFFFFFFFFu, pass);
	if (pass) {
		const u32 at = plen + __popc(bal & ((1u << lane) - 1));
		P.b[wib][at] = b; P.f[wib][at] = fj; P.q[wib][at] = qi;
	}
	plen += __popc(bal);
	__syncwarp();
	if (plen >= 32) {
		plen -= 32;
		const u32 pb = P.b[wib][plen + lane], pf = P.f[wib][plen + lane], pq = P.q[wib][plen + lane];
		__syncwarp();
		probe_buckets(A, Q, lane, true, pb, pf, pq, c_sectors);
	}
}
__device__ __forceinline__ void pend_drain(const ScanArgs &A, PendQueue &P, u32 wib, u32 plen, QueueCursor &Q, u32 lane, u64 &c_sectors)
{
	if (plen == 0) return;
	const bool act = lane < plen;
	const u32 pb = act ? P.b[wib][lane] : 0, pf = act ? P.f[wib][lane] : 0, pq = act ? P.q[wib][lane] : 0;
	probe_buckets(A, Q, lane, act, pb, pf, pq, c_sectors);
}

// Key of window j of strand s -> home bucket, fingerprint | j<<16, and the summary verdict.
__device__ __forceinline__ bool window_key(const ScanArgs &A, const u64 *__restrict__ s, u32 j, u32 &b, u32 &fj, u32 &part)
{
	u32 lead;
	const u64 hash = key_hash<LdGlobal>(s, j, A.T.h, lead);
	b = bucket_of(hash, lead, A.T, part);
	const u32 fp = hash_fp(hash);
	fj = fp | (j << 16);
	if (!A.T.summary) return true;
	const u32 sm = ld_na32(A.T.summary + b), need = summary_bits(fp);
	return (sm & need) == need || (sm & OGB_SPILLED);                        // else: no entry with this fingerprint, no fetch
}

template <int MODE>
__global__ void __launch_bounds__(256, OGB_PROBE_MINBLOCKS) k_probe(ScanArgs A)
{
	__shared__ PendQueue P;
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const u32 h = A.T.h;
	u64 c_probes = 0, c_sectors = 0;
	u32 plen = 0;
	QueueCursor Q = {0, 0, 0};                                               // this warp's piece of the candidate queue (warp-uniform)

	for (u32 qi = A.lo + gw; qi < A.hi; qi += nwarps) {
		if (MODE == MODE_OVERLAP && A.contained && ((__ldg(A.contained + (qi >> 5)) >> (qi & 31)) & 1)) continue;   // (:548)
		u64 off; u32 L1;
		read_geom(A.R, qi, off, L1);
		const u64 *s = A.R.words + off;
		const u32 nwin = L1 - h - 1;
		c_probes += nwin;
		for (u32 jb = 1; jb <= nwin; jb += 32) {
			const u32 j = jb + lane;
			u32 b = 0, fj = 0, part;
			const bool pass = j <= nwin && window_key(A, s, j, b, fj, part);
			pend_push(A, P, wib, plen, Q, lane, pass, b, fj, qi, c_sectors);
		}
	}
	pend_drain(A, P, wib, plen, Q, lane, c_sectors);
	queue_pad(A, Q, lane);
	for (int d = 16; d > 0; d >>= 1) c_sectors += __shfl_down_sync(0xFFFFFFFFu, c_sectors, d);
	if (lane == 0) { atomicAdd(A.ctr + CTR_PROBES, c_probes); atomicAdd(A.ctr + CTR_SECTORS, c_sectors); }
}

// k_probe for data sets with one read length (no contained reads, no per-read geometry): the windows
// of all reads of the launch are flattened over the threads, x -> (read, j) by an exact multiply-
// shift division, so every lane of every warp owns a window.
template <int MODE>
__global__ void __launch_bounds__(256, OGB_PROBE_MINBLOCKS) k_probe_uniform(ScanArgs A, u32 nwin, u64 div_magic)
{
	__shared__ PendQueue P;
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 total = (A.hi - A.lo) * nwin;                                  // < 2^32: a launch covers at most 2^16 reads
	const u32 rounds = (total + 31) >> 5;
	const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const u32 stride = 2 * A.R.uniform_pw;
	u64 c_sectors = 0;
	u32 plen = 0;
	QueueCursor Q = {0, 0, 0};

	for (u32 rd = gw; rd < rounds; rd += nwarps) {
		const u32 x = rd * 32 + lane;
		const u32 qr = nwin == 1 ? x : (u32)__umul64hi(div_magic, (u64)x);  // x / nwin (Lemire: exact for 32-bit x, nwin > 1)
		const u32 j = x - qr * nwin + 1, qi = A.lo + qr;
		u32 b = 0, fj = 0, part;
		const bool pass = x < total && window_key(A, A.R.words + (u64)qi * stride, j, b, fj, part);
		pend_push(A, P, wib, plen, Q, lane, pass, b, fj, qi, c_sectors);
	}
	pend_drain(A, P, wib, plen, Q, lane, c_sectors);
	queue_pad(A, Q, lane);
	for (int d = 16; d > 0; d >>= 1) c_sectors += __shfl_down_sync(0xFFFFFFFFu, c_sectors, d);
	if (lane == 0) atomicAdd(A.ctr + CTR_SECTORS, c_sectors);
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(A.ctr + CTR_PROBES, (u64)total);
}

// ------------------------------------------------------------------------------------------------
// Probe of an index that does not fit L2 (uniform read length): random 64-byte bucket fetches over
// hundreds of MB run at the DRAM row / translation rate (~25 G/s measured, profiles/gather_bench_b200.txt)
// instead of the L2 rate. So the windows of a chunk are first hashed, filtered by the summary and
// SCATTERED into one queue per hash partition (k_window_part: a block counts its tile's windows per
// partition in shared memory, reserves queue space with one atomic per partition and tile, then
// writes), and the queues are probed one partition at a time (k_probe_parts): while a partition is
// being probed its 32-64 MB of buckets are L2-resident.
// ------------------------------------------------------------------------------------------------
#define OGB_MAXPART 64
#ifndef OGB_WPT
#define OGB_WPT 4               // windows per thread and tile in k_window_part
#endif
struct PartQueue {
	u32 *b, *f, *q;             // nparts regions of cap records: bucket, fingerprint | j<<16, query read index
	u64 *cursor;                // records appended per partition (may exceed cap: the excess is dropped and the chunk retried)
	u64 cap;
	u32 nparts;
};

template <int MODE, bool UNIFORM>
__global__ void __launch_bounds__(256, 6) k_window_part(ScanArgs A, u32 nwin, u64 div_magic, PartQueue PQ)
{
	// nwin = windows per read (one read length) or of the longest read (mixed lengths: window x -> read x / nwin,
	// j = x % nwin + 1; the threads whose j lies past their read's last window, or whose read is contained, drop out)
	__shared__ u32 s_cnt[OGB_MAXPART];
	__shared__ u64 s_base[OGB_MAXPART];
	__shared__ u32 s_valid;
	const u32 total = (A.hi - A.lo) * nwin;
	const u32 stride = 2 * A.R.uniform_pw;
	const u32 tile_windows = 256 * OGB_WPT;
	const u32 tiles = (total + tile_windows - 1) / tile_windows;
	const u32 h = A.T.h;
	u32 c_valid = 0;
	if (threadIdx.x == 0) s_valid = 0;
	for (u32 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		if (threadIdx.x < PQ.nparts) s_cnt[threadIdx.x] = 0;
		__syncthreads();
		u32 eb[OGB_WPT], ef[OGB_WPT], eq[OGB_WPT], ep[OGB_WPT];            // ep = partition<<16 | rank inside the tile's share of it
		#pragma unroll
		for (int r = 0; r < OGB_WPT; r++) {
			const u32 x = tile * tile_windows + r * 256 + threadIdx.x;
			const u32 qr = nwin == 1 ? x : (u32)__umul64hi(div_magic, (u64)x);
			const u32 j = x - qr * nwin + 1;
			eq[r] = A.lo + qr; eb[r] = 0; ef[r] = 0;
			u32 part = 0;
			bool valid = x < total;
			const u64 *s = nullptr;
			if (UNIFORM) s = A.R.words + (u64)eq[r] * stride;                   // one read length: no geometry load, no contained reads
			else if (valid) {
				u64 off; u32 L1;
				read_geom(A.R, eq[r], off, L1);
				s = A.R.words + off;
				valid = j <= L1 - h - 1;
				if (MODE == MODE_OVERLAP && valid && A.contained) valid = !((__ldg(A.contained + (eq[r] >> 5)) >> (eq[r] & 31)) & 1);   // (:548)
				c_valid += valid;
			}
			const bool pass = valid && window_key(A, s, j, eb[r], ef[r], part);
			ep[r] = pass ? (part << 16) | atomicAdd(&s_cnt[part], 1u) : 0xFFFFFFFFu;
		}
		__syncthreads();
		if (threadIdx.x < PQ.nparts) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(PQ.cursor + threadIdx.x, (u64)s_cnt[threadIdx.x]) : 0;
		__syncthreads();
		#pragma unroll
		for (int r = 0; r < OGB_WPT; r++)
			if (ep[r] != 0xFFFFFFFFu) {
				const u32 part = ep[r] >> 16;
				const u64 at = s_base[part] + (ep[r] & 0xFFFFu);
				if (at < PQ.cap) { const u64 i = part * PQ.cap + at; PQ.b[i] = eb[r]; PQ.f[i] = ef[r]; PQ.q[i] = eq[r]; }
			}
		__syncthreads();
	}
	if (UNIFORM) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(A.ctr + CTR_PROBES, (u64)total); }
	else {
		c_valid = __reduce_add_sync(0xFFFFFFFFu, c_valid);
		if ((threadIdx.x & 31) == 0 && c_valid) atomicAdd(&s_valid, c_valid);
		__syncthreads();
		if (threadIdx.x == 0 && s_valid) atomicAdd(A.ctr + CTR_PROBES, (u64)s_valid);
	}
}

// Body of the partition probe for the warp `gw` of `nwarps` warps that share the work.
template <int MODE>
__device__ __forceinline__ void probe_parts_body(const ScanArgs &A, const PartQueue &PQ, u64 gw, u64 nwarps, u32 lane)
{
	u64 c_sectors = 0;
	QueueCursor Q = {0, 0, 0};
	for (u32 part = 0; part < PQ.nparts; part++) {
		const u64 filled = PQ.cursor[part];
		if (filled > PQ.cap && gw == 0 && lane == 0) atomicMax(A.ctr + CTR_PQ_OVERFLOW, filled);   // a skewed partition outgrew its queue: the host retries with more slack
		const u64 n = min(filled, PQ.cap), base = part * PQ.cap;
		for (u64 i0 = gw * 32; i0 < n; i0 += nwarps * 32) {
			const u64 i = i0 + lane;
			const bool act = i < n;
			const u32 pb = act ? PQ.b[base + i] : 0, pf = act ? PQ.f[base + i] : 0, pq = act ? PQ.q[base + i] : 0;
			probe_buckets(A, Q, lane, act, pb, pf, pq, c_sectors);
		}
	}
	queue_pad(A, Q, lane);
	for (int d = 16; d > 0; d >>= 1) c_sectors += __shfl_down_sync(0xFFFFFFFFu, c_sectors, d);
	if (lane == 0) atomicAdd(A.ctr + CTR_SECTORS, c_sectors);
}
template <int MODE>
__global__ void __launch_bounds__(256, OGB_PROBE_MINBLOCKS) k_probe_parts(ScanArgs A, PartQueue PQ)
{
	probe_parts_body<MODE>(A, PQ, (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5, (gridDim.x * (u64)blockDim.x) >> 5, threadIdx.x & 31);
}

// One thread per candidate. The loop is warp-uniform (32 consecutive candidates per warp and
// iteration): a read's candidates are adjacent in the queue, so the lanes of a warp mostly append to
// the same one or two nodes -- their deg[] increments are aggregated with __match_any_sync into one
// atomic per distinct node instead of ~26 serialised same-address atomics.
// Body of the verification for the threads `first + lane` of a warp, `stride` threads sharing the work (warp-uniform loop).
template <int MODE>
__device__ __forceinline__ void verify_body(const ScanArgs &A, u64 first, u64 stride, bool announce)
{
	const u64 total = min(*A.cand_cursor, A.cand_cap);
	if (announce) atomicMax(A.ctr + CTR_CAND_MAX, *A.cand_cursor);
	const u32 h = A.T.h, lane = threadIdx.x & 31, lt = (1u << lane) - 1;
	u32 c_cand = 0, c_hits = 0;
	for (u64 c0 = first; c0 < total; c0 += stride) {
		const u64 c = c0 + lane;
		bool ok = false;
		u32 qi = c < total ? A.cand_q[c] : OGB_NOCAND, ri = 0, o = 0, L1 = 0, L2 = 0, orient = 0, offset = 0;
		if (qi != OGB_NOCAND) {
			const u64 cv = A.cand_v[c];
			const u32 j = (u32)(cv >> 32), val = (u32)cv;
			if (val != 0) {                                                  // 0: spurious fingerprint hit on an empty slot
				c_cand++;
				u64 off, roff;
				read_geom(A.R, qi, off, L1);
				const u64 *s = A.R.words + off;
				ri = (val >> 2) - 1; o = val & 3;
				read_geom(A.R, ri, roff, L2);
				const u64 *t = A.R.words + roff + (o >> 1) * padded_words(L2);
				if (MODE == MODE_CONTAIN) {
					// OverlapGraph.cpp:256: read1 must be longer; :302-340 restated on the whole of read2.
					if (L1 > L2) {
						u32 a = 0; bool fits;
						if ((o & 1) == 0) { fits = L1 - j >= L2; a = j; }            // :316-321
						else { fits = j >= L2 - h; a = j - (L2 - h); }             // :331-336
						if (fits && region_equal_aligned<LdGlobal, LdGlobal>(s, a, t, L2)) {
							atomicMax(A.sup + ri, ((u64)L1 << 32) | (u64)(0xFFFFFFFFu - qi));   // :259-268
							c_hits++;
						}
					}
				} else if (!(A.contained && ((__ldg(A.contained + (ri >> 5)) >> (ri & 31)) & 1))) {   // :548 superReadID == 0
					u32 pa, len; bool fits;
					const u64 *pp, *qq;           // compare pp[pa..pa+len) with qq[0..len)
					if ((o & 1) == 0) {           // key = prefix of t: s[j..L1) must equal t[0..L1-j)      (:359-370)
						fits = L1 - j < L2;
						pp = s; pa = j; qq = t; len = L1 - j;
						orient = o == 0 ? 3 : 2;  // :552,:554
						offset = j;               // L1 - overlap, overlap = L1 - j
					} else {                      // key = suffix of t: s[0..j+h) must equal t[L2-h-j..L2)  (:371-382)
						fits = L2 - h >= j;
						pp = t; pa = L2 - h - j; qq = s; len = h + j;
						orient = o == 1 ? 0 : 1;  // :553,:555
						offset = L1 - h - j;      // L1 - overlap, overlap = h + j
					}
					if (!fits) ok = false;
					else if (A.R.uniform_pw == 4) ok = suffix_equals_prefix<4>(pp, pa, qq, len);      // 65..128 bp
					else if (A.R.uniform_pw == 6) ok = suffix_equals_prefix<6>(pp, pa, qq, len);      // 129..192 bp
					else ok = region_equal_aligned<LdGlobal, LdGlobal>(pp, pa, qq, len);
				}
			}
		}
		if (MODE == MODE_OVERLAP) {
			// Self-overlap: the reference inserts the edge and its twin object into the same list
			// (OverlapGraph.cpp:409-417); twin offset = (UINT16)(L2 + offset - L1) = offset -> two entries.
			const u32 ne = ok ? (ri == qi ? 2u : 1u) : 0u;
			const u32 grp = __match_any_sync(0xFFFFFFFFu, ok ? qi : (0x80000000u | lane));   // lanes appending to the same node
			const int leader = __ffs(grp) - 1;
			const u32 b1 = __ballot_sync(0xFFFFFFFFu, ne >= 1), b2 = __ballot_sync(0xFFFFFFFFu, ne == 2);
			const u32 all = __popc(grp & b1) + __popc(grp & b2);             // entries my group appends / those of lanes before me
			const u32 before = __popc(grp & b1 & lt) + __popc(grp & b2 & lt);
			u32 base = 0;
			if (ok && (int)lane == leader) base = atomicAdd(A.deg + qi, all);
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			if (ok) {
				const u64 e0 = make_edge(offset & 0xFFFF, ri + 1, orient), e1 = make_edge(offset & 0xFFFF, ri + 1, twin_orient(orient));
				for (u32 q = 0; q < ne; q++) {
					const u64 e = q ? e1 : e0;
					const u32 pos = base + before + q;
					if (pos < A.cap) A.slots_e[(u64)(qi - A.slot_lo) * A.cap + pos] = e;
					else {                                                   // heavy node: spill, placed by k_heavy_place
						if (pos == A.cap) atomicAdd(A.ctr + CTR_BIG_NODES, 1ull);   // exactly one entry of a heavy node lands here
						const u64 ov = atomicAdd(A.ctr + CTR_OVERFLOW, 1ull);
						if (ov < A.ov_cap) { A.ov_q[ov] = qi; A.ov_e[ov] = e; }
					}
				}
			}
		}
	}
	for (int d = 16; d > 0; d >>= 1) { c_cand += __shfl_down_sync(0xFFFFFFFFu, c_cand, d); c_hits += __shfl_down_sync(0xFFFFFFFFu, c_hits, d); }
	if (lane == 0) {
		if (c_cand) atomicAdd(A.ctr + CTR_CANDIDATES, (u64)c_cand);
		if (MODE == MODE_CONTAIN && c_hits) atomicAdd(A.ctr + CTR_CONTAIN_HITS, (u64)c_hits);
	}
}
template <int MODE>
__global__ void __launch_bounds__(256, OGB_VERIFY_MINBLOCKS) k_verify(ScanArgs A)
{
	verify_body<MODE>(A, (u64)blockIdx.x * blockDim.x + (threadIdx.x - (threadIdx.x & 31)), (u64)gridDim.x * blockDim.x, blockIdx.x == 0 && threadIdx.x == 0);
}
// Probe of one chunk and verification of the previous one in ONE launch, warp-specialised: warps 0-3 of every block probe,
// warps 4-7 verify. Launched one after the other each kernel fills the SMs with its own blocks, so the issue-bound probe and the
// gather-latency-bound verify never actually share an SM (two streams only overlap their tails); inside one block they do.
template <int MODE>
__global__ void __launch_bounds__(256, OGB_PROBE_MINBLOCKS) k_probe_verify(ScanArgs AP, PartQueue PQ, ScanArgs AV)
{
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	if (wib < 4) probe_parts_body<MODE>(AP, PQ, (u64)blockIdx.x * 4 + wib, (u64)gridDim.x * 4, lane);
	else verify_body<MODE>(AV, (u64)blockIdx.x * 128 + (wib - 4) * 32, (u64)gridDim.x * 128, blockIdx.x == 0 && threadIdx.x == 128);
}

// ------------------------------------------------------------------------------------------------
// The adjacency as the verify kernel leaves it -- and as K5 / K6 consume it, without any copy or sort:
//
//   slot regions  read idx of this rank owns slots[(idx - lo)*cap .. +cap) holding deg[idx] 8-byte edge words in
//                 discovery order (offset, dst, orient: what the node's OWN walk and the final records need). A heavy
//                 node (deg > cap) keeps its whole list in ext, at the word offset stored in its first slot
//                 (k_heavy_move / k_heavy_place).
//   rows          what a node looks like to everybody ELSE -- as a pivot (K5) and as a twin (K6) only dst and the
//                 strand it is left on matter: 4 bytes per entry. Read idx v (any rank) owns the 128-byte line
//                 rows[v*32 .. +32): word 0 = degree, word 1 = offset of its overflow entries inside its rank's
//                 segment of `more`, words 2..31 = entries 0..29 as dst<<1 | strand, in slot order; entries 30.. live in
//                 `more`. k_rows_finish derives both from the edge words in one streaming pass after K3.
//                 A pivot scan is therefore ONE aligned 128-byte gather at an address known from the edge itself
//                 (no degree / node-record fetch in front of it), plus a second gather only when deg > 30.
//   ebits         one ELIM bit per entry, same geometry: word v = row entries of read v, the overflow entries' bits
//                 follow from word `nrows` on. K5 writes the bits of its own nodes, K6 reads one bit per candidate.
//
// Several ranks: rows, `more` and ebits are laid out by GLOBAL read index / rank segment, every rank fills its own
// part and the parts are allgathered (C1 after K3, C2 after K5). Reading the peers' slot regions in place (CUDA IPC
// over NVLink) was built and measured in round 1: peer reads collapse once the mapped footprint exceeds the GPU's
// translation reach (profiles/r1_notes.md item 16).
//
// The low 14 bits of an edge word are annotations (make_edge leaves them 0): bit 1 = OGB_KEEP "survives the
// reduction" (K6, nodes with many survivors only), bits 2-13 = 1 + position of the twin entry in the destination's
// list (K5, nodes of degree > 32 only). Both are written by the owning warp.
// ------------------------------------------------------------------------------------------------
#define OGB_KEEP 2ull
#define OGB_ROW_W 32            // u32 words per adjacency row (one 128-byte line)
#define OGB_ROW_E 30            // entries held in the row itself
#define OGB_SURV 4              // candidates / survivors staged per node (a reduced node keeps ~2 edges)

struct GraphView {
	u64 *slots;                      // this rank's slot regions
	const u32 *deg;                  // by global read index, valid on [lo, hi)
	u64 *ext;                        // heavy lists of this rank
	u32 lo, cap;                     // first read index of this rank; slots per read
	u32 *rows;                       // every rank's rows, by global read index
	u32 *more;                       // overflow entries: rank r's segment starts at more + r*more_stride
	u32 *ebits;                      // ELIM bits: word v = row of read v; bit nrows*32 + g = overflow entry g
	u64 more_stride;                 // entries per rank segment of `more`
	u64 nrows;                       // rows allocated (reads per rank x ranks)
	u64 per_magic;                   // rank of read index v = umul64hi(per_magic, v) (0: one rank)
	u32 my_rank;
};

__device__ __forceinline__ u32 row_entry(u64 e) { return (edge_dst(e) << 1) | ((edge_orient(e) >> 1) & 1); }
__device__ __forceinline__ u32 rank_of(const GraphView &G, u32 vidx) { return G.per_magic ? (u32)__umul64hi(G.per_magic, (u64)vidx) : 0u; }
// 1 + bit address of the ELIM bit of entry `pos` of read index vidx; ovf = word 1 of its row (read only when pos >= OGB_ROW_E)
__device__ __forceinline__ u64 entry_bit(const GraphView &G, u32 vidx, u32 pos)
{
	if (pos < OGB_ROW_E) return (u64)vidx * 32 + pos + 1;
	const u32 ovf = __ldg(G.rows + (u64)vidx * OGB_ROW_W + 1);
	return G.nrows * 32 + (u64)rank_of(G, vidx) * G.more_stride + ovf + (pos - OGB_ROW_E) + 1;
}

// Rows of the own nodes [lo, hi) from their edge words: header, entries 0..29, and the entries beyond in this rank's
// overflow segment `more_own` (more_cap entries; the running cursor CTR_MORE_CURSOR hands out space, so the pass can run
// chunk by chunk right behind k_verify, while the chunk's slot regions are still in L2 -- and, on several ranks, the
// finished rows of a chunk travel to the peers while the next chunk is probed). A warp takes 32 nodes at a time: lane i
// looks at node i's degree, the overflow space of all 32 is reserved with one atomic, then the warp copies node after
// node with lane k on entry k (coalesced on both sides). HEAVY = false: nodes with more edges than slots are left for
// the HEAVY = true pass, which runs once their lists have been gathered in ext (k_heavy_move / k_heavy_place) and also
// appends (node, row) records to hrows for the peers. (Writing the row entries from k_verify, next to the edge words, was
// measured first: scattered 4-byte stores cost 2.2 ms at config 3 against 0.7-0.9 ms for this pass.)
#define OGB_HROW_W 34           // u32 words per heavy-row record: node index, pad, 32 row words
template <bool HEAVY>
__global__ void __launch_bounds__(256) k_rows_finish(GraphView G, u32 lo, u32 hi, u32 *__restrict__ more_own, u64 more_cap, u64 *ctr,
                                                     u32 *__restrict__ hrows, u64 hrows_cap)
{
	const u32 lane = threadIdx.x & 31;
	const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (u32 u0 = lo + gw * 32; u0 < hi; u0 += nwarps * 32) {
		const u32 mine = u0 + lane;
		u32 dm = mine < hi ? G.deg[mine] : 0;
		const bool heavy = dm > G.cap;
		if (heavy != HEAVY) dm = HEAVY ? 0 : 0xFFFFFFFFu;                     // not this pass's node (0: nothing to do; ~0: skip, keep the row)
		const u32 need = dm != 0xFFFFFFFFu && dm > OGB_ROW_E ? dm - OGB_ROW_E : 0;
		u32 inc = need;
		#pragma unroll
		for (int s = 1; s < 32; s <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, s); if (lane >= (u32)s) inc += t; }
		const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
		u64 base = 0;
		if (total) {
			if (lane == 31) base = atomicAdd(ctr + CTR_MORE_CURSOR, (u64)total);
			base = __shfl_sync(0xFFFFFFFFu, base, 31);
		}
		const u64 offm = base + inc - need;                                   // this lane's node: start of its overflow entries
		const u32 cnt = min(32u, hi - u0);
		for (u32 j = 0; j < cnt; j++) {
			const u32 u = u0 + j;
			const u32 d = __shfl_sync(0xFFFFFFFFu, dm, j);
			const u64 off = __shfl_sync(0xFFFFFFFFu, offm, j);
			u32 *row = G.rows + (u64)u * OGB_ROW_W;
			if (d == 0xFFFFFFFFu) continue;
			if (d == 0) { if (!HEAVY && lane == 0) row[0] = 0; continue; }
			const u64 *own = G.slots + (u64)(u - G.lo) * G.cap;
			if (HEAVY) own = G.ext + own[0];
			u32 wv = lane == 0 ? d : (u32)off;                                  // header: degree, overflow offset
			if (lane >= 2) wv = lane - 2 < d ? row_entry(own[lane - 2]) : 0;
			row[lane] = wv;
			if (HEAVY && hrows) {
				u64 at = 0;
				if (lane == 0) at = atomicAdd(ctr + CTR_HROW_CURSOR, 1ull);
				at = __shfl_sync(0xFFFFFFFFu, at, 0);
				if (at < hrows_cap) { u32 *r = hrows + at * OGB_HROW_W; if (lane == 0) { r[0] = u; r[1] = 0; } r[2 + lane] = wv; }
			}
			if (d > OGB_ROW_E && off + (d - OGB_ROW_E) <= more_cap)             // beyond the capacity: the host sees the cursor and retries with more
				for (u32 k = OGB_ROW_E + lane; k < d; k += 32) more_own[off + k - OGB_ROW_E] = row_entry(own[k]);
		}
	}
}
// Heavy-row records of the other ranks (allgathered at a common stride) into the local rows
__global__ void __launch_bounds__(256) k_scatter_hrows(const u32 *__restrict__ hrows, u64 stride, const u64 *__restrict__ counts, u32 nranks, u32 my_rank, u32 *__restrict__ rows)
{
	const u32 lane = threadIdx.x & 31;
	const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
	for (u32 r = 0; r < nranks; r++) {
		if (r == my_rank) continue;
		for (u64 i = gw; i < counts[r]; i += nwarps) {
			const u32 *rec = hrows + (r * stride + i) * OGB_HROW_W;
			rows[(u64)rec[0] * OGB_ROW_W + lane] = rec[2 + lane];
		}
	}
}

// Heavy nodes (repeats): the cap edges in the slot region and the spilled ones are gathered in ext.
__global__ void __launch_bounds__(256) k_heavy_move(u64 *__restrict__ slots, const u32 *__restrict__ deg, u32 lo, u32 hi, u32 cap,
                                                   u64 *__restrict__ ext, u64 ext_cap, u32 *__restrict__ fill, u64 *ctr)
{
	const u32 u = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= hi) return;
	const u32 d = deg[u];
	if (d <= cap) return;
	const u64 off = atomicAdd(ctr + CTR_EXT_CURSOR, (u64)d);
	if (off + d > ext_cap) { atomicAdd(ctr + CTR_SCRATCH_FAIL, 1ull); return; }
	u64 *s = slots + (u64)(u - lo) * cap;
	for (u32 k = 0; k < cap; k++) ext[off + k] = s[k];
	fill[u - lo] = cap;
	s[0] = off;
}
__global__ void __launch_bounds__(256) k_heavy_place(const u32 *__restrict__ ov_q, const u64 *__restrict__ ov_e, u64 n_over, const u64 *__restrict__ slots,
                                                    u32 lo, u32 cap, u64 *__restrict__ ext, u32 *__restrict__ fill)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_over) return;
	const u32 u = ov_q[i];
	const u64 off = slots[(u64)(u - lo) * cap];
	ext[off + atomicAdd(fill + (u - lo), 1u)] = ov_e[i];
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
// K5: transitive-edge marking, one warp per node (OverlapGraph.cpp:574-615), straight on the unsorted
// slot regions. The reference walks the node's edges in sorted order (:563) and skips the ones whose
// destination is no longer INPLAY (:583); states only ever go INPLAY -> ELIMINATED, so "the next
// pivot" is the smallest (offset, dst, orient, slot) above the current one whose destination is still
// INPLAY: a warp min-reduction (two REDUX + one ballot) per ACTIVE pivot -- about two per node --
// replaces the per-node sort. The neighbour set (destination node -> INPLAY/ELIMINATED) is an open-
// addressing set in shared memory (degree <= 256) or in a global scratch pool (larger): one word per
// neighbour, dst | state<<31.
//
// The kernel is bound by the chain of dependent gathers own list -> pivot row -> next pivot row. Nodes of
// degree <= 32 (one edge per lane, in registers) fetch the rows of their two likely pivots at once: the
// first edge overall and the first edge leaving the node on the other side -- the latter is almost always
// the second and last active pivot. The fetch is only a prefetch into registers; the walk itself follows
// the reference order.
//
// Output, for K6: the ELIM bits of the own entries, and per node the edges its own marking leaves (its
// active pivots that were not eliminated later: ~2) with the bit address of the twin entry's ELIM bit,
// which the pivot scan met on its way. K6 never has to read a list again.
// ------------------------------------------------------------------------------------------------
struct MarkArgs {
	GraphView G;
	u32 hi;                     // node indices [G.lo, hi) of this rank
	u32 *cntc;                  // K5 -> K6: candidates (edges left by the own marking) per own node, by idx - lo
	u64 *cand;                  // OGB_SURV records of two words per own node: edge word, 1 + bit address of the twin's ELIM bit (0 = unknown)
	u64 *big;                   // nodes with more candidates: records of three words ((idx-lo)<<32 | slot, edge word, twin bit address)
	u64 big_cap;
	u32 *cnt;                   // K6: survivors per own node, by idx - lo
	u32 *scratch_keys;          // global pool for the neighbour sets of big nodes
	u64 scratch_cap;
	u64 *ctr;
};

#define OGB_SET_ELIM 0x80000000u
#define OGB_SET_KEY 0x3FFFFFFFu
__device__ __forceinline__ u32 set_hash(u32 key, u32 capmask) { return (key * 2654435761u) >> 7 & capmask; }

__device__ __forceinline__ u32 set_insert(u32 *keys, u32 capmask, u32 key)
{
	u32 s = set_hash(key, capmask);
	for (;;) {
		u32 cur = atomicCAS(keys + s, 0u, key);
		if (cur == 0 || (cur & OGB_SET_KEY) == key) return s;
		s = (s + 1) & capmask;
	}
}
__device__ __forceinline__ int set_find(const u32 *keys, u32 capmask, u32 key)
{
	u32 s = set_hash(key, capmask);
	for (;;) {
		u32 cur = keys[s];
		if ((cur & OGB_SET_KEY) == key) return (int)s;
		if (cur == 0) return -1;
		s = (s + 1) & capmask;
	}
}
// INPLAY -> ELIMINATED for key, if it is a neighbour (:591-596). Lanes may race on one slot; they all store the same word.
__device__ __forceinline__ void set_eliminate(u32 *keys, u32 capmask, u32 key)
{
	u32 s = set_hash(key, capmask);
	for (;;) {
		const u32 cur = keys[s];
		if ((cur & OGB_SET_KEY) == key) { if (!(cur & OGB_SET_ELIM)) keys[s] = cur | OGB_SET_ELIM; return; }
		if (cur == 0) return;
		s = (s + 1) & capmask;
	}
}

// Lane holding the smallest (w, lane) among the lanes with cand set, -1 if there is none. w < 2^64-2^32. The high
// half (offset | top bits of dst) is almost always decisive: the second reduction runs only on a tie.
__device__ __forceinline__ int pick_min(bool cand, u64 w)
{
	const u32 hi = cand ? (u32)(w >> 32) : 0xFFFFFFFFu;
	const u32 mh = __reduce_min_sync(0xFFFFFFFFu, hi);
	if (mh == 0xFFFFFFFFu) return -1;
	const bool c2 = cand && hi == mh;
	const u32 b2 = __ballot_sync(0xFFFFFFFFu, c2);
	if ((b2 & (b2 - 1)) == 0) return __ffs(b2) - 1;
	const u32 lo = c2 ? (u32)w : 0xFFFFFFFFu;
	const u32 ml = __reduce_min_sync(0xFFFFFFFFu, lo);
	return __ffs(__ballot_sync(0xFFFFFFFFu, c2 && lo == ml)) - 1;
}

// Neighbour set of a node of degree <= 32*S WITHOUT a probe loop: the destinations stay in the lanes' registers (S per
// lane, entry k = lane + 32*slot), two byte tables per warp map hash(dst) -> entry (1024*S and 256 slots; a destination
// that loses its slot in the first goes to the second). A lookup reads the entry number, fetches that entry's destination
// with indexed shuffles and compares: stale or foreign table bytes can never produce a hit, so the tables are never
// cleared. The states are bit masks in registers (bit k = the destination REPRESENTED by entry k was eliminated; every
// entry knows its representative = the entry the tables return for its destination, which also covers multi-edges).
#ifndef OGB_T1_BITS
#define OGB_T1_BITS 10
#endif
#define OGB_T1 (1 << OGB_T1_BITS)
#define OGB_T2 256
#define OGB_FALLBACK 0xFFFFFFFFu   // cntc value: the fast kernel hands the node to k_mark_any
template <int S> __device__ __forceinline__ u32 rs_h1(u32 key) { return (key * 2654435761u) >> (32 - OGB_T1_BITS - (S - 1)); }
__device__ __forceinline__ u32 rs_h2(u32 key) { return (key * 0x85EBCA6Bu) >> 24; }
// destination held by entry j (every lane of the warp must call this)
template <int S> __device__ __forceinline__ u32 rs_fetch(const u32 (&dst)[S], u32 j)
{
	const u32 v0 = __shfl_sync(0xFFFFFFFFu, dst[0], j & 31);
	if (S == 1) return v0;
	const u32 v1 = __shfl_sync(0xFFFFFFFFu, dst[S - 1], j & 31);
	return (j & 32) ? v1 : v0;
}
// adds to hit[] the representative entry of x (this lane's pivot entry) if x is a neighbour and `on` is set
template <int S> __device__ __forceinline__ void rs_lookup(const unsigned char *t1, const unsigned char *t2, bool two, const u32 (&dst)[S], u32 x, bool on, u32 (&hit)[S])
{
	const u32 j1 = t1[rs_h1<S>(x)] & (32 * S - 1);
	const bool m1 = rs_fetch<S>(dst, j1) == x;
	u32 j = j1; bool m = m1;
	if (two) {
		const u32 j2 = t2[rs_h2(x)] & (32 * S - 1);
		const bool m2 = rs_fetch<S>(dst, j2) == x;
		if (!m1) { j = j2; m = m2; }
	}
	if (on && m) hit[S == 1 ? 0 : (j >> 5)] |= 1u << (j & 31);
}
// scan of pivot v's list (row word r of this lane, overflow entries in `more`): returns its degree, ORs the
// representatives of the destinations it eliminates into elim[], twin = 1 + position of an entry (v, self), 0 if none
template <int S> __device__ __forceinline__ u32 scan_pivot_rs(const GraphView &G, u32 v, u32 t1o, u32 self, const unsigned char *t1, const unsigned char *t2, bool two,
                                                              const u32 (&dst)[S], u32 lane, u32 r, u32 &twin, u32 (&elim)[S])
{
	const u32 dv = __shfl_sync(0xFFFFFFFFu, r, 0);
	const u32 want = t1o & 1;                                                // compatible(): the pivot is entered and left on the same strand
	const bool valid = lane >= 2 && lane - 2 < dv;
	u32 tw = valid && (r >> 1) == self ? lane - 1 : 0;
	u32 hit[S];
	#pragma unroll
	for (int q = 0; q < S; q++) hit[q] = 0;
	rs_lookup<S>(t1, t2, two, dst, r >> 1, valid && (r & 1) == want, hit);
	if (dv > OGB_ROW_E) {                                                    // entries 30.. live in the rank's overflow segment
		const u32 ovf = __shfl_sync(0xFFFFFFFFu, r, 1);
		const u32 *m = G.more + (u64)rank_of(G, v - 1) * G.more_stride + ovf;
		for (u32 k0 = 0; k0 < dv - OGB_ROW_E; k0 += 32) {
			const u32 kk = k0 + lane;
			const bool in = kk < dv - OGB_ROW_E;
			const u32 f = in ? __ldg(m + kk) : 0;
			if (in && (f >> 1) == self) tw = OGB_ROW_E + kk + 1;
			rs_lookup<S>(t1, t2, two, dst, f >> 1, in && (f & 1) == want, hit);
		}
	}
	twin = __reduce_max_sync(0xFFFFFFFFu, tw);
	#pragma unroll
	for (int q = 0; q < S; q++) elim[q] |= __reduce_or_sync(0xFFFFFFFFu, hit[q]);
	return dv;
}
// Entry with the smallest (key, entry number) among the candidate entries, -1 if there is none
template <int S> __device__ __forceinline__ int pick_entry(const bool (&cand)[S], const u64 (&w)[S], u32 lane)
{
	if (S == 1) return pick_min(cand[0], w[0]);
	const bool second = cand[S - 1] && (!cand[0] || w[S - 1] < w[0]);          // this lane's better candidate
	const int l = pick_min(cand[0] || cand[S - 1], second ? w[S - 1] : w[0]);
	if (l < 0) return -1;
	return l + (__shfl_sync(0xFFFFFFFFu, (u32)second, l) ? 32 : 0);
}
template <int S> __device__ __forceinline__ u64 entry_word(const u64 (&e)[S], int k)
{
	const u64 v0 = __shfl_sync(0xFFFFFFFFu, e[0], k & 31);
	if (S == 1) return v0;
	const u64 v1 = __shfl_sync(0xFFFFFFFFu, e[S - 1], k & 31);
	return (k & 32) ? v1 : v0;
}
// ELIM bits of the own entries 30.. (`nbits` of them in `bits`, entry 30 first): up to three atomics into the overflow bit area
__device__ __forceinline__ void publish_more_bits(const GraphView &G, u32 ovf, u64 bits, u32 lane)
{
	if (bits == 0) return;
	const u64 a = G.nrows * 32 + (u64)G.my_rank * G.more_stride + ovf;
	const u32 sh = (u32)a & 31;
	u32 *w = G.ebits + (a >> 5);
	if (lane < 3) {
		// the bits occupy positions sh .. sh+63 of a 96-bit field starting at word w
		u32 v;
		if (lane == 0) v = (u32)(bits << sh);
		else if (lane == 1) v = (u32)(sh ? bits >> (32 - sh) : bits >> 32);
		else v = sh ? (u32)(bits >> (64 - sh)) : 0;
		if (v) atomicOr(w + lane, v);
	}
}

#ifndef OGB_MARK_MINBLOCKS
#define OGB_MARK_MINBLOCKS 5
#endif
// K5 for the nodes of degree (32*(S-1), 32*S]: one warp per node, S edges per lane in registers. A warp takes 32
// consecutive nodes, looks at their degrees and walks the ones of its class.
template <int S>
__global__ void __launch_bounds__(OGB_WARPS * 32, S == 1 ? OGB_MARK_MINBLOCKS : 4) k_mark_fast(MarkArgs A)
{
	__shared__ unsigned char s_t1[OGB_WARPS][OGB_T1 * S], s_t2[OGB_WARPS][OGB_T2];
	const GraphView &G = A.G;
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lt = (1u << lane) - 1;
	const u32 gw = blockIdx.x * OGB_WARPS + wib, nwarps = gridDim.x * OGB_WARPS;
	u64 c_entries = 0, c_pivots = 0;
	unsigned char *t1 = s_t1[wib], *t2 = s_t2[wib];

	for (u32 u0 = G.lo + gw * 32; u0 < A.hi; u0 += nwarps * 32) {
		const u32 mine = u0 + lane;
		const u32 dm = mine < A.hi ? G.deg[mine] : 0;
		if (S == 1 && mine < A.hi && dm == 0) A.cntc[mine - G.lo] = 0;         // nodes without edges
		for (u32 todo = __ballot_sync(0xFFFFFFFFu, dm > 32 * (S - 1) && dm <= 32 * S); todo; todo &= todo - 1) {
			const u32 jn = __ffs(todo) - 1, u = u0 + jn, self = u + 1;
			const u32 d = __shfl_sync(0xFFFFFFFFu, dm, jn);
			const u64 *own = G.slots + (u64)(u - G.lo) * G.cap;
			if (d > G.cap) own = G.ext + own[0];
			u64 e[S], w[S];
			u32 dst[S], rep[S], mytw[S], elim[S], done[S];
			bool have[S], lose[S];
			#pragma unroll
			for (int q = 0; q < S; q++) {
				have[q] = lane + 32 * q < d;
				e[q] = have[q] ? own[lane + 32 * q] : 0;
				w[q] = edge_key(e[q]);
				dst[q] = edge_dst(e[q]);                                         // 0 on idle entries: never equals a pivot entry
				mytw[q] = 0; elim[q] = 0; done[q] = 0;
			}
			// ---- tables
			__syncwarp();
			#pragma unroll
			for (int q = 0; q < S; q++) if (have[q]) t1[rs_h1<S>(dst[q])] = (unsigned char)(lane + 32 * q);
			__syncwarp();
			bool any_lose = false;
			#pragma unroll
			for (int q = 0; q < S; q++) {
				rep[q] = t1[rs_h1<S>(dst[q])] & (32 * S - 1);
				lose[q] = rs_fetch<S>(dst, rep[q]) != dst[q] && have[q];
				any_lose |= lose[q];
			}
			const bool two = __any_sync(0xFFFFFFFFu, any_lose);
			bool fail = false;
			if (two) {
				#pragma unroll
				for (int q = 0; q < S; q++) if (lose[q]) t2[rs_h2(dst[q])] = (unsigned char)(lane + 32 * q);
				__syncwarp();
				any_lose = false;
				#pragma unroll
				for (int q = 0; q < S; q++) {
					const u32 j2 = t2[rs_h2(dst[q])] & (32 * S - 1);
					const bool ok2 = rs_fetch<S>(dst, j2) == dst[q];
					if (lose[q]) { rep[q] = j2; any_lose |= !ok2; }
				}
				fail = __any_sync(0xFFFFFFFFu, any_lose);                        // both slots taken by other destinations: rare
			}
			if (fail) { if (lane == 0) A.cntc[u - G.lo] = OGB_FALLBACK; continue; }

			// ---- walk: the first pivot, the first edge on the other side of u (its row is fetched together with the first), then
			// the smallest edge not walked yet whose destination is still INPLAY (everything smaller was walked or eliminated)
			const int a = pick_entry<S>(have, w, lane);
			const u64 ea = entry_word<S>(e, a);
			bool side[S];
			#pragma unroll
			for (int q = 0; q < S; q++) side[q] = have[q] && ((edge_orient(e[q]) ^ edge_orient(ea)) & 1);
			const int b = pick_entry<S>(side, w, lane);
			const u64 eb = entry_word<S>(e, b < 0 ? 0 : b);
			const u32 ra = __ldg(G.rows + (u64)(edge_dst(ea) - 1) * OGB_ROW_W + lane);
			const u32 rb = b >= 0 ? __ldg(G.rows + (u64)(edge_dst(eb) - 1) * OGB_ROW_W + lane) : 0;
			int p = a;
			u64 ep = ea;
			u32 rp = ra;
			for (;;) {
				u32 tw;
				c_pivots++; c_entries += scan_pivot_rs<S>(G, edge_dst(ep), edge_orient(ep), self, t1, t2, two, dst, lane, rp, tw, elim);
				#pragma unroll
				for (int q = 0; q < S; q++) { if ((int)lane + 32 * q == p) mytw[q] = tw; if ((p >> 5) == q) done[q] |= 1u << (p & 31); }
				bool cand[S];
				#pragma unroll
				for (int q = 0; q < S; q++) cand[q] = have[q] && !((done[q] >> lane) & 1) && !((elim[S == 1 ? 0 : (rep[q] >> 5)] >> (rep[q] & 31)) & 1);
				p = pick_entry<S>(cand, w, lane);
				if (p < 0) break;
				ep = entry_word<S>(e, p);
				rp = p == b ? rb : __ldg(G.rows + (u64)(edge_dst(ep) - 1) * OGB_ROW_W + lane);
			}
			// ---- ELIM bits of the own entries (:601-607; the twin half is applied in k_keep), candidates for K6
			bool left[S];
			u32 gm[S], lb[S];
			#pragma unroll
			for (int q = 0; q < S; q++) {
				const bool gone = have[q] && ((elim[S == 1 ? 0 : (rep[q] >> 5)] >> (rep[q] & 31)) & 1);
				left[q] = have[q] && !gone;
				gm[q] = __ballot_sync(0xFFFFFFFFu, gone);
				lb[q] = __ballot_sync(0xFFFFFFFFu, left[q]);
			}
			if (lane == 0) G.ebits[u] = gm[0] & ((1u << OGB_ROW_E) - 1);
			if (d > OGB_ROW_E) {
				u64 bits = gm[0] >> OGB_ROW_E;
				if (S > 1) bits |= (u64)gm[S - 1] << (32 - OGB_ROW_E);
				publish_more_bits(G, __ldg(G.rows + (u64)u * OGB_ROW_W + 1), bits, lane);
			}
			u32 c = 0;
			#pragma unroll
			for (int q = 0; q < S; q++) c += __popc(lb[q]);
			if (lane == 0) A.cntc[u - G.lo] = c;
			u64 base = 0;
			if (c > OGB_SURV) {                                                  // rare: the node's candidates go to the record list
				if (lane == 0) base = atomicAdd(A.ctr + CTR_BIGREC_CURSOR, (u64)c);
				base = __shfl_sync(0xFFFFFFFFu, base, 0);
			}
			u32 before = 0;
			#pragma unroll
			for (int q = 0; q < S; q++) {
				if (left[q]) {
					const u32 rk = before + __popc(lb[q] & lt);
					const u64 addr = mytw[q] ? entry_bit(G, dst[q] - 1, mytw[q] - 1) : 0;
					if (c <= OGB_SURV) { u64 *o = A.cand + ((u64)(u - G.lo) * OGB_SURV + rk) * 2; o[0] = e[q]; o[1] = addr; }
					else if (base + rk < A.big_cap) { u64 *o = A.big + (base + rk) * 3; o[0] = ((u64)(u - G.lo) << 32) | (lane + 32 * q); o[1] = e[q]; o[2] = addr; }
				}
				before += __popc(lb[q]);
			}
		}
	}
	if (lane == 0) { atomicAdd(A.ctr + CTR_PIVOT_ENTRIES, c_entries); atomicAdd(A.ctr + CTR_ACTIVE_PIVOTS, c_pivots); }
}

// Hash-set form of the pivot scan (k_mark_any): adjacency of pivot v (1-based id) against the neighbour set: a neighbour
// reached through v on the strand v was entered on becomes ELIMINATED (:588-596). r = this lane's word of v's row.
// twin = 1 + position of an entry (v, self) in v's list (0 if there is none). Returns v's degree.
__device__ __forceinline__ u32 scan_pivot(const GraphView &G, u32 v, u32 t1, u32 self, u32 *keys, u32 capmask, u32 lane, u32 r, u32 &twin)
{
	const u32 dv = __shfl_sync(0xFFFFFFFFu, r, 0);
	const u32 want = t1 & 1;
	u32 tw = 0;
	if (lane >= 2 && lane - 2 < dv) {
		const u32 x = r >> 1;
		if (x == self) tw = lane - 1;
		if ((r & 1) == want) set_eliminate(keys, capmask, x);
	}
	if (dv > OGB_ROW_E) {
		const u32 ovf = __shfl_sync(0xFFFFFFFFu, r, 1);
		const u32 *m = G.more + (u64)rank_of(G, v - 1) * G.more_stride + ovf;
		for (u32 kk = lane; kk < dv - OGB_ROW_E; kk += 32) {
			const u32 f = __ldg(m + kk);
			if ((f >> 1) == self) tw = OGB_ROW_E + kk + 1;
			if ((f & 1) == want) set_eliminate(keys, capmask, f >> 1);
		}
	}
	twin = __reduce_max_sync(0xFFFFFFFFu, tw);
	__syncwarp();
	return dv;
}
// ELIM bit of the own entry `pos` >= 30: an atomic into the overflow bit area
__device__ __forceinline__ void publish_more_bit(const GraphView &G, u32 ovf, u32 pos)
{
	const u64 a = G.nrows * 32 + (u64)G.my_rank * G.more_stride + ovf + (pos - OGB_ROW_E);
	atomicOr(G.ebits + (a >> 5), 1u << (a & 31));
}

// K5 for everything else (degree > 64, or a node whose destinations did not fit the byte tables): edges stay in memory
// (L1), a lane looks after entries lane, lane+32, ...; neighbour set = hash set in shared memory or the scratch pool.
__global__ void __launch_bounds__(OGB_WARPS * 32, 6) k_mark_any(MarkArgs A)
{
	__shared__ u32 s_keys[OGB_WARPS][OGB_SETCAP];
	const GraphView &G = A.G;
	const u32 lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lt = (1u << lane) - 1;
	const u32 gw = blockIdx.x * OGB_WARPS + wib, nwarps = gridDim.x * OGB_WARPS;
	u64 c_entries = 0, c_pivots = 0;

	for (u32 u0 = G.lo + gw * 32; u0 < A.hi; u0 += nwarps * 32) {
		const u32 mine = u0 + lane;
		const u32 dm = mine < A.hi ? G.deg[mine] : 0;
		for (u32 todo = __ballot_sync(0xFFFFFFFFu, dm > 64 || (dm && A.cntc[mine - G.lo] == OGB_FALLBACK)); todo; todo &= todo - 1) {
			const u32 jn = __ffs(todo) - 1, u = u0 + jn, self = u + 1;
			const u32 d = __shfl_sync(0xFFFFFFFFu, dm, jn);
			u64 *own = G.slots + (u64)(u - G.lo) * G.cap;
			if (d > G.cap) own = G.ext + own[0];
			u32 cap = 128;
			while (cap < 2 * d) cap <<= 1;
			u32 *keys;
			if (cap <= OGB_SETCAP) keys = s_keys[wib];
			else {
				u64 base = 0;
				if (lane == 0) base = atomicAdd(A.ctr + CTR_SCRATCH_CURSOR, (u64)cap);
				base = __shfl_sync(0xFFFFFFFFu, base, 0);
				if (base + cap > A.scratch_cap) { if (lane == 0) { atomicAdd(A.ctr + CTR_SCRATCH_FAIL, 1ull); A.cntc[u - G.lo] = 0; } continue; }
				keys = A.scratch_keys + base;
			}
			const u32 capmask = cap - 1;
			for (u32 i = lane; i < cap; i += 32) keys[i] = 0;
			__syncwarp();
			for (u32 k = lane; k < d; k += 32) set_insert(keys, capmask, edge_dst(own[k]));
			__syncwarp();
			u64 cw = 0; u32 ck = 0; bool first = true;
			for (;;) {
				u64 bw = ~0ull; u32 bk = 0xFFFFFFFFu;                         // this lane's smallest in-play entry above (cw, ck)
				for (u32 k = lane; k < d; k += 32) {
					const u64 x = edge_key(own[k]);
					if ((first || x > cw || (x == cw && k > ck)) && x < bw && !(keys[set_find(keys, capmask, edge_dst(x))] & OGB_SET_ELIM)) { bw = x; bk = k; }
				}
				const u32 hi = (u32)(bw >> 32);
				const u32 mh = __reduce_min_sync(0xFFFFFFFFu, hi);
				if (mh == 0xFFFFFFFFu) break;
				const bool c2 = hi == mh;
				const u32 ml = __reduce_min_sync(0xFFFFFFFFu, c2 ? (u32)bw : 0xFFFFFFFFu);
				const bool c3 = c2 && (u32)bw == ml;
				ck = __reduce_min_sync(0xFFFFFFFFu, c3 ? bk : 0xFFFFFFFFu);
				cw = ((u64)mh << 32) | ml; first = false;
				u32 tw;
				const u32 rp = __ldg(G.rows + (u64)(edge_dst(cw) - 1) * OGB_ROW_W + lane);
				c_pivots++; c_entries += scan_pivot(G, edge_dst(cw), edge_orient(cw), self, keys, capmask, lane, rp, tw);
				if (lane == 0 && tw && tw < 4096) own[ck] |= (u64)tw << 2;      // twin position: 12 spare bits of the pivot's edge word
				__syncwarp();
			}
			// ELIM bits of the own entries, candidates
			const u32 ovf = __ldg(G.rows + (u64)u * OGB_ROW_W + 1);
			u32 rowbits = 0, c = 0;
			for (u32 kb = 0; kb < d; kb += 32) {
				const u32 k = kb + lane;
				const u64 x = k < d ? own[k] : 0;
				const bool elim = k < d && (keys[set_find(keys, capmask, edge_dst(x))] & OGB_SET_ELIM);
				if (elim) { if (k < OGB_ROW_E) rowbits |= 1u << k; else publish_more_bit(G, ovf, k); }
				const bool left = k < d && !elim;
				const u32 lb = __ballot_sync(0xFFFFFFFFu, left);
				if (left) {
					const u32 at = c + __popc(lb & lt), tw = edge_twin(x);
					if (at < OGB_SURV) { u64 *q = A.cand + ((u64)(u - G.lo) * OGB_SURV + at) * 2; q[0] = edge_key(x); q[1] = tw ? entry_bit(G, edge_dst(x) - 1, tw - 1) : 0; }
				}
				c += __popc(lb);
			}
			rowbits = __reduce_or_sync(0xFFFFFFFFu, rowbits);
			if (lane == 0) { G.ebits[u] = rowbits; A.cntc[u - G.lo] = c; }
			if (c > OGB_SURV) {                                                  // rare: all candidates go to the big list instead
				u64 base = 0;
				if (lane == 0) base = atomicAdd(A.ctr + CTR_BIGREC_CURSOR, (u64)c);
				base = __shfl_sync(0xFFFFFFFFu, base, 0);
				u32 done = 0;
				for (u32 kb = 0; kb < d; kb += 32) {
					const u32 k = kb + lane;
					const u64 x = k < d ? own[k] : 0;
					const bool left = k < d && !(keys[set_find(keys, capmask, edge_dst(x))] & OGB_SET_ELIM);
					const u32 lb = __ballot_sync(0xFFFFFFFFu, left);
					if (left) {
						const u64 at = base + done + __popc(lb & lt);
						const u32 tw = edge_twin(x);
						if (at < A.big_cap) { u64 *q = A.big + at * 3; q[0] = ((u64)(u - G.lo) << 32) | k; q[1] = edge_key(x); q[2] = tw ? entry_bit(G, edge_dst(x) - 1, tw - 1) : 0; }
					}
					done += __popc(lb);
				}
			}
		}
	}
	if (lane == 0) { atomicAdd(A.ctr + CTR_PIVOT_ENTRIES, c_entries); atomicAdd(A.ctr + CTR_ACTIVE_PIVOTS, c_pivots); }
}

// ------------------------------------------------------------------------------------------------
// K6: an edge (u,w) survives iff it was not flagged by u's marking and its twin was not flagged by
// w's marking (:605-606, :623-661). Marks are per destination NODE, so w's verdict on u is the ELIM bit
// of any (w,u) entry of w's list, and K5 has left the address of that bit next to every candidate: one
// bit gather per candidate (~2 per node), no list is read. One THREAD per node; the survivors (<= OGB_SURV)
// are staged in surv[] for k_emit_small and counted in cnt[] for the scan that positions them. The rare
// nodes with more candidates go through the record list (k_keep_big) and k_emit.
// ------------------------------------------------------------------------------------------------

// 1 + bit address of a (v, self) entry found by searching v's list, 0 if there is none (only when K5 could not note it).
__device__ __forceinline__ u64 find_twin_bit(const GraphView &G, u32 vidx, u32 self)
{
	const u32 *row = G.rows + (u64)vidx * OGB_ROW_W;
	const u32 dv = __ldg(row);
	for (u32 k = 0; k < dv && k < OGB_ROW_E; k++) if ((__ldg(row + 2 + k) >> 1) == self) return entry_bit(G, vidx, k);
	if (dv > OGB_ROW_E) {
		const u32 *m = G.more + (u64)rank_of(G, vidx) * G.more_stride + __ldg(row + 1);
		for (u32 k = 0; k < dv - OGB_ROW_E; k++) if ((__ldg(m + k) >> 1) == self) return entry_bit(G, vidx, OGB_ROW_E + k);
	}
	return 0;
}
__device__ __forceinline__ bool twin_eliminated(const GraphView &G, u64 addr, u32 vidx, u32 self, bool &found)
{
	if (addr == 0) addr = find_twin_bit(G, vidx, self);
	found = addr != 0;
	if (!found) return false;
	addr--;
	return (ld_na32(G.ebits + (addr >> 5)) >> (addr & 31)) & 1;
}

__global__ void __launch_bounds__(256) k_keep(MarkArgs A, u64 *__restrict__ surv)
{
	const GraphView &G = A.G;
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	u32 ns = 0, asym = 0;
	if (G.lo + i < A.hi) {
		const u32 c = A.cntc[i];
		if (c && c <= OGB_SURV) {
			u64 e[OGB_SURV], a[OGB_SURV];
			#pragma unroll
			for (int q = 0; q < OGB_SURV; q++) if (q < (int)c) { e[q] = A.cand[((u64)i * OGB_SURV + q) * 2]; a[q] = A.cand[((u64)i * OGB_SURV + q) * 2 + 1]; }
			#pragma unroll
			for (int q = 0; q < OGB_SURV; q++) if (q < (int)c) {
				bool found;
				const bool gone = twin_eliminated(G, a[q], edge_dst(e[q]) - 1, G.lo + i + 1, found);
				asym += !found;
				if (!gone) surv[(u64)i * OGB_SURV + ns++] = e[q];
			}
		}
		A.cnt[i] = ns;                                                       // nodes on the big list: k_keep_big adds theirs
	}
	const u32 nodes = __popc(__ballot_sync(0xFFFFFFFFu, ns > 0));
	if ((threadIdx.x & 31) == 0 && nodes) atomicAdd(A.ctr + CTR_NODES_FINAL, (u64)nodes);
	if (asym) atomicAdd(A.ctr + CTR_ASYMMETRIC, (u64)asym);
}
// One thread per record of the big list: survivors get OGB_KEEP in the node's own list (k_emit writes them).
__global__ void __launch_bounds__(256) k_keep_big(MarkArgs A)
{
	const GraphView &G = A.G;
	const u64 n_rec = min(A.ctr[CTR_BIGREC_CURSOR], A.big_cap);             // the host never learns the count: fixed grid, strided
	for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (u64)gridDim.x * blockDim.x) {
	const u64 *q = A.big + r * 3;
	const u32 i = (u32)(q[0] >> 32), k = (u32)q[0];
	bool found;
	const bool gone = twin_eliminated(G, q[2], edge_dst(q[1]) - 1, G.lo + i + 1, found);
	if (!found) atomicAdd(A.ctr + CTR_ASYMMETRIC, 1ull);
	if (gone) continue;
	u64 *own = G.slots + (u64)i * G.cap;
	if (G.deg[G.lo + i] > G.cap) own = G.ext + own[0];
	own[k] |= OGB_KEEP;
	if (atomicAdd(A.cnt + i, 1u) == 0) atomicAdd(A.ctr + CTR_NODES_FINAL, 1ull);
	}
}

// Exclusive scan of u32 counts into u64 offsets: (1) per-block sums, (2) one block scans the sums,
// (3) per-block scan + base.
#define OGB_SCAN_ITEMS 2048     // per block of 256 threads (8 per thread)
// max_out / more_out (degree scan only): largest count, and the number of entries beyond the adjacency rows (count - 30 each).
__global__ void __launch_bounds__(256) k_scan_sums(const u32 *__restrict__ cnt, u32 n, u64 *__restrict__ sums, u64 *max_out, u64 *more_out)
{
	__shared__ u64 sh[8];
	u64 base = (u64)blockIdx.x * OGB_SCAN_ITEMS, acc = 0;
	u32 mx = 0, more = 0;
	for (u32 i = threadIdx.x; i < OGB_SCAN_ITEMS; i += 256) if (base + i < n) { const u32 v = cnt[base + i]; acc += v; mx = max(mx, v); more += v > 30 ? v - 30 : 0; }
	if (max_out) {
		for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, d));
		if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_out, (u64)mx);
	}
	if (more_out) {
		more = __reduce_add_sync(0xFFFFFFFFu, more);
		if ((threadIdx.x & 31) == 0 && more) atomicAdd(more_out, (u64)more);
	}
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

__device__ __forceinline__ ogb_edge edge_record(u32 src, u64 e)
{
	ogb_edge r;
	r.src = src; r.dst = edge_dst(e); r.offset = (uint16_t)edge_offset(e); r.orient = (uint8_t)edge_orient(e); r.reserved = 0;
	return r;
}

// Final records of the nodes with at most OGB_SURV survivors (nearly all: a reduced node keeps ~2
// edges): one thread per node sorts the staged words and writes them at pos[u - lo] (+ add).
__global__ void __launch_bounds__(256) k_emit_small(const u64 *__restrict__ surv, const u32 *__restrict__ cnt, const u32 *__restrict__ cntc, const u64 *__restrict__ pos,
                                                    ogb_edge *__restrict__ out, u32 lo, u32 hi, u64 add)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (lo + i >= hi) return;
	const u32 c = cnt[i];
	if (c == 0 || cntc[i] > OGB_SURV) return;                                // more candidates than staging slots: k_emit writes that node
	u64 w[OGB_SURV];
	#pragma unroll
	for (int q = 0; q < OGB_SURV; q++) w[q] = q < (int)c ? surv[(u64)i * OGB_SURV + q] : ~0ull;
	// staged in slot order, so a stable sort keeps the slot tie-break of equal keys
	#pragma unroll
	for (int x = 1; x < OGB_SURV; x++)
		#pragma unroll
		for (int y = x; y > 0; y--)
			if (edge_key(w[y]) < edge_key(w[y - 1]) && w[y] != ~0ull) { const u64 t = w[y]; w[y] = w[y - 1]; w[y - 1] = t; }
	ogb_edge *dst = out + pos[i] + add;
	#pragma unroll
	for (int q = 0; q < OGB_SURV; q++) if (q < (int)c) dst[q] = edge_record(lo + i + 1, w[q]);
}

// Edge records of the own nodes in the reference order (src, offset, dst, orient): the surviving
// edges (OGB_KEEP) of the nodes with more than min_cnt candidates in cnt[] (ALL = false) or every edge (ALL = true;
// pre-reduction list for tests / keep_pre). One warp per node; pos[u - lo] (+ add) is where the node's
// records start. Entries are ranked by counting the smaller ones -- shuffles up to degree 32, a plain
// double loop beyond.
template <bool ALL>
__global__ void __launch_bounds__(OGB_WARPS * 32) k_emit(const u64 *__restrict__ own_slots, const u64 *__restrict__ own_ext, const u32 *__restrict__ deg,
                                                         const u32 *__restrict__ cnt, const u64 *__restrict__ pos, ogb_edge *__restrict__ out,
                                                         u32 lo, u32 hi, u32 cap, u64 add, u32 min_cnt)
{
	const u32 lane = threadIdx.x & 31;
	const u32 gw = blockIdx.x * OGB_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * OGB_WARPS;
	// a warp looks at 32 nodes at a time and visits only the ones that have records to write here
	for (u32 base = lo + gw * 32; base < hi; base += nwarps * 32)
	for (u32 todo = __ballot_sync(0xFFFFFFFFu, base + lane < hi && (ALL ? deg[base + lane] > 0 : cnt[base + lane - lo] > min_cnt)); todo; todo &= todo - 1) {
		const u32 u = base + __ffs(todo) - 1;
		const u32 d = deg[u];
		const u64 *own = own_slots + (u64)(u - lo) * cap;
		if (d > cap) own = own_ext + own[0];
		ogb_edge *dst = out + pos[u - lo] + add;
		if (d <= 32) {
			const u64 e = lane < d ? own[lane] : 0, w = edge_key(e);
			const bool s = lane < d && (ALL || (e & OGB_KEEP));
			u32 rank = 0;
			for (u32 m = __ballot_sync(0xFFFFFFFFu, s); m; m &= m - 1) {
				const u32 b = __ffs(m) - 1;
				const u64 o = __shfl_sync(0xFFFFFFFFu, w, b);
				rank += (o < w) || (o == w && b < lane);
			}
			if (s) dst[rank] = edge_record(u + 1, e);
		} else {
			for (u32 k = lane; k < d; k += 32) {
				const u64 e = own[k], w = edge_key(e);
				if (!(ALL || (e & OGB_KEEP))) continue;
				u32 rank = 0;
				for (u32 k2 = 0; k2 < d; k2++) {
					const u64 e2 = own[k2], w2 = edge_key(e2);
					if (ALL || (e2 & OGB_KEEP)) rank += (w2 < w) || (w2 == w && k2 < k);
				}
				dst[rank] = edge_record(u + 1, e);
			}
		}
	}
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
	u32 part, lead;
	u64 hash = key_hash<LdGlobal>(key, 0, T.h, lead);
	u32 fp = hash_fp(hash), b = bucket_of(hash, lead, T, part), c = 0;
	const u32 pend = (part + 1) * T.part_buckets;
	for (u32 steps = 0; steps < T.part_buckets; steps++) {
		u32 w[OGB_BWORDS];
		load_bucket(T.slots, b, w);
		for (u32 mm = match_bucket(w, fp); mm; mm &= mm - 1) {
			u32 val = bucket_value(w, __ffs(mm) - 1), ri = (val >> 2) - 1, o = val & 3;
			if (val == 0) continue;
			u64 off; u32 L;
			read_geom(R, ri, off, L);
			const u64 *t = R.words + off + (o >> 1) * padded_words(L);
			if (!region_equal<LdGlobal, LdGlobal>(key, 0, t, (o & 1) ? L - T.h : 0, T.h)) continue;
			if (pass == 1) out[pos[k] + c] = (u64)(ri + 1) | ((u64)o << 62);
			c++;
		}
		if (w[5 + OGB_SLOTS - 1] == 0) break;
		b = next_bucket(b, pend, T);
	}
	if (pass == 0) cnt[k] = c;
}

// ------------------------------------------------------------------------------------------------
// Measurement helper (ogb_gather_ceiling): random, aligned, non-allocating gathers of BYTES from a buffer of
// nblocks * BYTES bytes (nblocks a power of two), four independent ones in flight per thread -- the access
// pattern of the index probes, the partner-strand fetches and the pivot-row fetches, with nothing else around it.
// ------------------------------------------------------------------------------------------------
template <int BYTES>
__global__ void __launch_bounds__(256) k_gather_ceiling(const u64 *__restrict__ buf, u64 nblocks, u32 per_thread, u64 *out)
{
	const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	u64 acc = 0;
	for (u32 it = 0; it < per_thread; it += 4) {
		u64 v[4][BYTES / 8];
		#pragma unroll
		for (int u = 0; u < 4; u++) {
			u64 x = tid * 0x9E3779B97F4A7C15ULL + it + u;
			x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
			const u64 *p = buf + (x & (nblocks - 1)) * (BYTES / 8);
			#pragma unroll
			for (int q = 0; q < BYTES / 32; q++)
				asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[u][4 * q]), "=l"(v[u][4 * q + 1]), "=l"(v[u][4 * q + 2]), "=l"(v[u][4 * q + 3]) : "l"(p + 4 * q));
		}
		#pragma unroll
		for (int u = 0; u < 4; u++)
			#pragma unroll
			for (int q = 0; q < BYTES / 8; q++) acc ^= v[u][q];
	}
	if (acc == 0x1234567) out[0] = acc;
}

// Order-independent checksum of a list of edge records (xor and sum of a 64-bit mix of every tuple): the bench and
// the full-size tests compare it with the oracle's figure without moving GBs of edges to the host.
__global__ void __launch_bounds__(256) k_edge_checksum(const ogb_edge *__restrict__ e, u64 n, u64 *out)
{
	u64 x_or = 0, x_sum = 0;
	for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
		const ogb_edge r = e[i];
		u64 x = (u64)r.src * 0x9E3779B97F4A7C15ULL ^ (u64)r.dst * 0xC2B2AE3D27D4EB4FULL ^ (u64)r.offset * 0x165667B19E3779F9ULL ^ (u64)r.orient * 0x27D4EB2F165667C5ULL;
		x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
		x_or ^= x; x_sum += x;
	}
	for (int d = 16; d > 0; d >>= 1) { x_or ^= __shfl_down_sync(0xFFFFFFFFu, x_or, d); x_sum += __shfl_down_sync(0xFFFFFFFFu, x_sum, d); }
	if ((threadIdx.x & 31) == 0) { atomicXor(out, x_or); atomicAdd(out + 1, x_sum); }
}

// ------------------------------------------------------------------------------------------------
// Mate-pair pass (Dataset::storeMatePairInformation, Dataset.cpp:208-310), batched: one thread per sequence as
// sequenced. Filter like Dataset.cpp:268 (length > minOverlap, ACGT only, testRead :398-413), canonical strand,
// getReadFromString (:421-455) as ONE index lookup -- the read's own prefix key is in the table with o = 0 -- verified
// against the whole stored strand, redirection to the super read (:280-284; sup[] still holds K2's packed maximum),
// and the orientation bit: 1 iff the sequence is a substring of that read's forward strand (:291-292).
// out_id = ID of the read that stands for the sequence (0: filtered out or not in the data set), out_or = the bit.
// ------------------------------------------------------------------------------------------------
#define OGB_MATE_MAXW 32         // sequences up to 1024 bases on this path (longer ones: the caller's host loop)
__global__ void __launch_bounds__(128) k_mate_lookup(ReadStore R, Table T, const u64 *__restrict__ sup, const char *__restrict__ bases, const u64 *__restrict__ offsets,
                                                     u64 n_seqs, u32 min_overlap, u32 *__restrict__ out_id, unsigned char *__restrict__ out_or)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_seqs) return;
	out_id[i] = 0; out_or[i] = 0;
	const char *s = bases + offsets[i];
	const u32 L = (u32)(offsets[i + 1] - offsets[i]);
	if (L <= min_overlap || L > 32 * (OGB_MATE_MAXW - 2)) return;               // :268 (strict); overlong: left to the host
	u64 fw[OGB_MATE_MAXW], rc[OGB_MATE_MAXW];
	const u32 nw = (L + 31) >> 5;
	for (u32 k = 0; k < nw + 2; k++) { fw[k] = 0; rc[k] = 0; }
	u32 cnt[4] = {0, 0, 0, 0};
	for (u32 p = 0; p < L; p++) {
		const u32 ch = (u32)(unsigned char)s[p] & ~0x20u;                       // toupper (:262-266)
		if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') return;           // testRead :403-406
		u32 c = (ch >> 1) & 3;
		cnt[c]++;
		c ^= c >> 1;                                                          // A C G T -> 0 1 2 3
		fw[p >> 5] |= (u64)c << (62 - 2 * (p & 31));
		const u32 q = L - 1 - p;
		rc[q >> 5] |= (u64)(3 - c) << (62 - 2 * (q & 31));
	}
	const u32 thr = (u32)(L * .8);                                            // :409
	if (cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) return;
	int pick = 0;                                                             // canonical strand = min(read, reverse complement) (:161-164)
	for (u32 k = 0; k < nw && pick == 0; k++) if (fw[k] != rc[k]) pick = fw[k] < rc[k] ? 1 : 2;
	const u64 *can = pick == 2 ? rc : fw;
	// getReadFromString: the canonical strand's prefix key, o = 0 entries, whole-strand compare
	u32 lead;
	const u64 hash = key_hash<LdShared>(can, 0, T.h, lead);
	u32 part;
	u32 b = bucket_of(hash, lead, T, part);
	const u32 fp = hash_fp(hash), pend = (part + 1) * T.part_buckets;
	u32 found = 0;
	for (u32 steps = 0; steps < T.part_buckets && !found; steps++) {
		u32 w[OGB_BWORDS];
		load_bucket(T.slots, b, w);
		for (u32 mm = match_bucket(w, fp); mm && !found; mm &= mm - 1) {
			const u32 val = bucket_value(w, __ffs(mm) - 1);
			if (val == 0 || (val & 3) != 0) continue;
			const u32 ri = (val >> 2) - 1;
			u64 off; u32 L2;
			read_geom(R, ri, off, L2);
			if (L2 != L) continue;
			const u64 *t = R.words + off;
			bool same = true;
			for (u32 k = 0; k < nw && same; k++) same = __ldg(t + k) == can[k];  // both zero-padded behind the last base
			if (same) found = ri + 1;
		}
		if (w[5 + OGB_SLOTS - 1] == 0) break;
		b = next_bucket(b, pend, T);
	}
	if (!found) return;
	u32 id = found;
	bool forward = pick != 2;                                                 // the stored strand IS the sequence (or the sequence is its own reverse complement)
	const u64 sv = sup ? sup[found - 1] : 0;
	if (sv) {                                                                 // contained: the super read stands for it (:280-284)
		id = (0xFFFFFFFFu - (u32)(sv & 0xFFFFFFFFull)) + 1;
		u64 off; u32 LR;
		read_geom(R, id - 1, off, LR);
		const u64 *t = R.words + off;                                         // forward strand of the super read
		forward = false;
		for (u32 sh = 0; sh + L <= LR && !forward; sh++) forward = region_equal<LdGlobal, LdShared>(t, sh, fw, 0, L);   // string::find (:291-292)
	}
	out_id[i] = id;
	out_or[i] = forward ? 1 : 0;
}
