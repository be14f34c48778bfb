// povu_b200 -- data-parallel primitives used by the decompose / call pipelines:
//   generic (non-commutative) scans, stable LSD radix sort of (u32 key, u32 value) pairs,
//   block-decomposed sparse-table range-minimum queries, Wyllie list ranking.
// All are HBM-bound integer kernels: coalesced 128-byte warp accesses, 32-byte-per-thread blocked
// loads, warp match/ballot ranking, shared-memory staging of digit offsets; grids are multiples of the
// SM count.  Under PV_EMU the same entry points run host std:: algorithms (debug rehearsal only).
#pragma once
#include <algorithm>
#include <numeric>
#include "pv_common.cuh"

namespace pv {

// ---------------------------------------------------------------------------------------------
// scans
// ---------------------------------------------------------------------------------------------
struct OpAddU32 { PV_HD u32 operator()(u32 a, u32 b) const { return a + b; } };
struct OpAddU64 { PV_HD u64 operator()(u64 a, u64 b) const { return a + b; } };
struct OpXor128 { PV_HD u128 operator()(u128 a, u128 b) const { return u128{a.lo ^ b.lo, a.hi ^ b.hi}; } };
struct OpAddI32 { PV_HD i32 operator()(i32 a, i32 b) const { return a + b; } };

#ifndef PV_EMU
static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T, class Op>
__device__ __forceinline__ T block_exclusive_scan_256(T v, Op op, T identity, T *sm /*[256]*/, T *total) {
	// Hillis-Steele over 256 thread totals in shared memory (order preserving -> valid for
	// non-commutative operators such as function composition).
	int t = threadIdx.x;
	sm[t] = v;
	__syncthreads();
	T acc = v;
#pragma unroll
	for (int d = 1; d < 256; d <<= 1) {
		T other = identity;
		bool has = t >= d;
		if (has) other = sm[t - d];
		__syncthreads();
		if (has) acc = op(other, acc);
		sm[t] = acc;
		__syncthreads();
	}
	if (total) *total = sm[255];
	T excl = t == 0 ? identity : sm[t - 1];
	__syncthreads();
	return excl;
}

template <class T, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T *in, T *block_sums, size_t n, Op op, T identity) {
	__shared__ T sm[SCAN_THREADS];
	size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
	T acc = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + k;
		if (i < n) acc = op(acc, in[i]);
	}
	T total;
	block_exclusive_scan_256(acc, op, identity, sm, &total);
	if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <class T, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const T *in, T *out, const T *block_offsets, size_t n, Op op, T identity) {
	__shared__ T sm[SCAN_THREADS];
	size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
	T vals[SCAN_ITEMS];
	T acc = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + k;
		vals[k] = i < n ? in[i] : identity;
		acc = op(acc, vals[k]);
	}
	T excl = block_exclusive_scan_256(acc, op, identity, sm, (T *)nullptr);
	T run = block_offsets ? op(block_offsets[blockIdx.x], excl) : excl;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		size_t i = base + k;
		if (INCLUSIVE) { run = op(run, vals[k]); if (i < n) out[i] = run; }
		else { if (i < n) out[i] = run; run = op(run, vals[k]); }
	}
}

template <class T, class Op, bool INCLUSIVE>
inline void scan_impl(Ctx &c, const T *in, T *out, size_t n, Op op, T identity) {
	if (n == 0) return;
	size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (nblocks == 1) {
		c.prof_begin("scan_apply", n, n * 2 * sizeof(T));
		scan_apply_kernel<T, Op, INCLUSIVE><<<1, SCAN_THREADS, 0, c.stream>>>(in, out, (const T *)nullptr, n, op, identity);
		c.prof_end();
		c.launches++;
		PV_CUDA_CHECK(cudaGetLastError());
		return;
	}
	Buf<T> sums(c, nblocks);
	c.prof_begin("scan_reduce", n, n * sizeof(T));
	scan_reduce_kernel<T, Op><<<(unsigned)nblocks, SCAN_THREADS, 0, c.stream>>>(in, sums.p, n, op, identity);
	c.prof_end();
	c.launches++;
	scan_impl<T, Op, false>(c, sums.p, sums.p, nblocks, op, identity);
	c.prof_begin("scan_apply", n, n * 2 * sizeof(T));
	scan_apply_kernel<T, Op, INCLUSIVE><<<(unsigned)nblocks, SCAN_THREADS, 0, c.stream>>>(in, out, sums.p, n, op, identity);
	c.prof_end();
	c.launches++;
	PV_CUDA_CHECK(cudaGetLastError());
}
#else
template <class T, class Op, bool INCLUSIVE>
inline void scan_impl(Ctx &c, const T *in, T *out, size_t n, Op op, T identity) {
	c.launches++;
	T run = identity;
	for (size_t i = 0; i < n; i++) {
		T v = in[i];
		if (INCLUSIVE) { run = op(run, v); out[i] = run; }
		else { out[i] = run; run = op(run, v); }
	}
}
#endif

template <class T, class Op> inline void exclusive_scan(Ctx &c, const T *in, T *out, size_t n, Op op, T identity) {
	scan_impl<T, Op, false>(c, in, out, n, op, identity);
}
template <class T, class Op> inline void inclusive_scan(Ctx &c, const T *in, T *out, size_t n, Op op, T identity) {
	scan_impl<T, Op, true>(c, in, out, n, op, identity);
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of (u32 key, u32 value) pairs on key bits [0, bits)
// Warp-private sub-tiles: every warp owns RS_ROUNDS*32 consecutive elements, ranks them with
// __match_any_sync and keeps its 256 running digit offsets in shared memory, so no block barrier is
// needed inside the scatter loop and stability is by construction.
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
static const int RS_WARPS = 8;
static const int RS_MAX_ROUNDS = 64;        // a warp sub-tile holds 32 * rounds elements; rounds is chosen per sort so that
                                            // small inputs still spread over every SM (latency-bound otherwise)

__global__ void __launch_bounds__(RS_WARPS * 32) rs_hist_kernel(const u32 *keys, u32 *counts, size_t n, size_t nsub, int shift, int rounds) {
	__shared__ u32 hist[RS_WARPS][256];
	int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int d = lane; d < 256; d += 32) hist[w][d] = 0;
	__syncwarp();
	size_t sub = (size_t)blockIdx.x * RS_WARPS + w;
	if (sub < nsub) {
		size_t base = sub * (size_t)(32 * rounds);
#pragma unroll 4
		for (int r = 0; r < rounds; r++) {
			size_t i = base + (size_t)r * 32 + lane;
			bool valid = i < n;
			u32 d = valid ? ((keys[i] >> shift) & 255u) : 256u + lane;   // invalid lanes get unique pseudo-digits
			u32 peers = __match_any_sync(0xFFFFFFFFu, d);
			if (valid && (peers & ((1u << lane) - 1u)) == 0) hist[w][d] += __popc(peers);
			__syncwarp();
		}
		for (int d = lane; d < 256; d += 32) counts[(size_t)d * nsub + sub] = hist[w][d];
	}
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_scatter_kernel(const u32 *keys_in, const u32 *vals_in, u32 *keys_out, u32 *vals_out,
								   const u32 *offsets, size_t n, size_t nsub, int shift, int rounds) {
	__shared__ u32 off[RS_WARPS][256];
	int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	size_t sub = (size_t)blockIdx.x * RS_WARPS + w;
	if (sub >= nsub) return;
	for (int d = lane; d < 256; d += 32) off[w][d] = offsets[(size_t)d * nsub + sub];
	__syncwarp();
	size_t base = sub * (size_t)(32 * rounds);
#pragma unroll 4
	for (int r = 0; r < rounds; r++) {
		size_t i = base + (size_t)r * 32 + lane;
		bool valid = i < n;
		u32 k = valid ? keys_in[i] : 0, v = valid ? vals_in[i] : 0;
		u32 d = valid ? ((k >> shift) & 255u) : 256u + lane;
		u32 peers = __match_any_sync(0xFFFFFFFFu, d);
		u32 rank = __popc(peers & ((1u << lane) - 1u));
		u32 o = valid ? off[w][d] : 0;
		__syncwarp();
		if (valid && rank == 0) off[w][d] = o + __popc(peers);
		__syncwarp();
		if (valid) { keys_out[o + rank] = k; vals_out[o + rank] = v; }
	}
}
#endif

// Sorts in place (result ends up back in keys/vals); tmp buffers are allocated internally.
inline void radix_sort_pairs(Ctx &c, u32 *keys, u32 *vals, size_t n, int bits) {
	if (n <= 1 || bits <= 0) return;
#ifndef PV_EMU
	Buf<u32> k2(c, n), v2(c, n);
	// about 32 warps per SM: rounds = n / (32 lanes * 32 warps * #SM), clamped to [4, 64]
	int rounds = (int)(n / ((size_t)1024 * (size_t)c.sm_count));
	rounds = rounds < 4 ? 4 : (rounds > RS_MAX_ROUNDS ? RS_MAX_ROUNDS : rounds);
	const size_t sub_elems = (size_t)32 * rounds;
	size_t nsub = (n + sub_elems - 1) / sub_elems;
	Buf<u32> counts(c, 256 * nsub);
	unsigned blocks = (unsigned)((nsub + RS_WARPS - 1) / RS_WARPS);
	u32 *ki = keys, *vi = vals, *ko = k2.p, *vo = v2.p;
	for (int shift = 0; shift < bits; shift += 8) {
		c.prof_begin("rs_hist", n, n * 4);
		rs_hist_kernel<<<blocks, RS_WARPS * 32, 0, c.stream>>>(ki, counts.p, n, nsub, shift, rounds);
		c.prof_end();
		c.launches++;
		exclusive_scan(c, counts.p, counts.p, 256 * nsub, OpAddU32(), 0u);
		c.prof_begin("rs_scatter", n, n * 16);
		rs_scatter_kernel<<<blocks, RS_WARPS * 32, 0, c.stream>>>(ki, vi, ko, vo, counts.p, n, nsub, shift, rounds);
		c.prof_end();
		c.launches++;
		PV_CUDA_CHECK(cudaGetLastError());
		std::swap(ki, ko); std::swap(vi, vo);
	}
	if (ki != keys) { c.d2d(keys, ki, n * 4); c.d2d(vals, vi, n * 4); }
#else
	c.launches++;
	std::vector<u32> idx(n);
	std::iota(idx.begin(), idx.end(), 0u);
	u32 mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
	std::stable_sort(idx.begin(), idx.end(), [&](u32 a, u32 b) { return (keys[a] & mask) < (keys[b] & mask); });
	std::vector<u32> k(n), v(n);
	for (size_t i = 0; i < n; i++) { k[i] = keys[idx[i]]; v[i] = vals[idx[i]]; }
	memcpy(keys, k.data(), n * 4); memcpy(vals, v.data(), n * 4);
#endif
}

inline int bits_for(u64 max_value) { int b = 1; while (b < 64 && (max_value >> b)) b++; return b; }

// ---------------------------------------------------------------------------------------------
// range-minimum queries over a u32 array: blocks of 32 with in-block prefix/suffix minima and a
// sparse table over the block minima.  query(l, r) is inclusive on both ends.
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
PV_DEV u32 pv_clz(u32 x) { return (u32)__clz((int)x); }
#else
inline u32 pv_clz(u32 x) { return (u32)__builtin_clz(x); }
#endif

struct RmqView {
	const u32 *a; const u32 *pre; const u32 *suf; const u32 *table; u32 n; u32 nb;
	PV_DEV u32 query(u32 l, u32 r) const {
		u32 bl = l >> 5, br = r >> 5;
		if (bl == br) {
			u32 m = a[l];
			for (u32 i = l + 1; i <= r; i++) { u32 x = a[i]; m = x < m ? x : m; }
			return m;
		}
		u32 m = suf[l]; u32 x = pre[r]; m = x < m ? x : m;
		if (br > bl + 1) {
			u32 len = br - bl - 1;
			u32 k = 31u - pv_clz(len);
			const u32 *row = table + (size_t)k * nb;
			x = row[bl + 1]; m = x < m ? x : m;
			x = row[br - (1u << k)]; m = x < m ? x : m;
		}
		return m;
	}
	// largest index j <= r with a[j] < bound, or NONE ("previous smaller value"): scan r's block, then
	// skip candidate-free block spans by descending the sparse table, then scan the hit block.
	PV_DEV u32 prev_smaller(u32 r, u32 bound) const {
		u32 j = r;
		while (true) {
			if (a[j] < bound) return j;
			if ((j & 31u) == 0) break;
			j--;
		}
		u32 pos = j >> 5;   // blocks [pos, block(r)] hold no candidate
		for (u32 k = levels; k-- > 0;) {
			u32 span = 1u << k;
			if (pos >= span && table[(size_t)k * nb + (pos - span)] >= bound) pos -= span;
		}
		if (pos == 0) return NONE;
		u32 b = pos - 1, s = b << 5;
		for (u32 i = s + 31u;; i--) {
			if (a[i] < bound) return i;
			if (i == s) break;
		}
		return NONE;   // unreachable when the table is consistent
	}
	u32 levels;
};

struct Rmq {
	Buf<u32> pre, suf, table;
	u32 n = 0, nb = 0, levels = 0;
	const u32 *a = nullptr;
	RmqView view() const { return RmqView{a, pre.p, suf.p, table.p, n, nb, levels}; }
};

PV_KERNEL(K_rmq_blocks, 12);   // a in, pre + suf out (+ block min)
PV_KERNEL(K_rmq_level, 12);

inline void rmq_build(Ctx &c, Rmq &r, const u32 *a, u32 n) {
	r.a = a; r.n = n; r.nb = (n + 31) / 32;
	r.levels = 1; while ((1u << r.levels) <= r.nb) r.levels++;
	r.pre = Buf<u32>(c, n); r.suf = Buf<u32>(c, n); r.table = Buf<u32>(c, (size_t)r.levels * r.nb);
	u32 *pre = r.pre.p, *suf = r.suf.p, *table = r.table.p; u32 nb = r.nb;
	// one thread per block of 32: 128 contiguous bytes per thread (4 sectors), sequential scan
	parallel_for<K_rmq_blocks>(c, nb, PV_LAMBDA(size_t b) {
		u32 s = (u32)b << 5, e = s + 32 < n ? s + 32 : n;
		u32 m = NONE;
		for (u32 i = s; i < e; i++) { u32 x = a[i]; m = x < m ? x : m; pre[i] = m; }
		table[b] = m;
		m = NONE;
		for (u32 i = e; i-- > s;) { u32 x = a[i]; m = x < m ? x : m; suf[i] = m; }
	});
	for (u32 k = 1; k < r.levels; k++) {
		u32 half = 1u << (k - 1);
		const u32 *prev = table + (size_t)(k - 1) * nb;
		u32 *cur = table + (size_t)k * nb;
		parallel_for<K_rmq_level>(c, nb, PV_LAMBDA(size_t j) {
			u32 x = prev[j];
			u32 y = j + half < nb ? prev[j + half] : NONE;
			cur[j] = x < y ? x : y;
		});
	}
}

// ---------------------------------------------------------------------------------------------
// list ranking (Wyllie pointer jumping).  Input: next[i] (NONE terminates), weight[i].
// Output: rank[i] = sum of weights of the elements strictly AFTER i on its list ... plus nothing;
// i.e. suffix sums excluding i.  Packed as (next << 32 | acc) so every hop is one 8-byte load.
// ---------------------------------------------------------------------------------------------
PV_KERNEL(K_lr_init, 16);
PV_KERNEL(K_lr_jump, 24);      // a[i], a[next] in, b[i] out (8 B each)
PV_KERNEL(K_lr_out, 12);

inline void list_rank_suffix(Ctx &c, const u32 *next, const u32 *weight, u32 *rank_out, size_t n) {
	if (n == 0) return;
	Buf<u64> A(c, n), B(c, n);
	u64 *a = A.p, *b = B.p;
	// acc(i) = weight of i's successor chain heads: start with acc = weight[next[i]] semantics via two-step:
	// we keep acc(i) = sum of weights of elements in (i, jump(i)]  (initially weight[next[i]]).
	parallel_for<K_lr_init>(c, n, PV_LAMBDA(size_t i) {
		u32 nx = next[i];
		u32 w = nx == NONE ? 0u : weight[nx];
		a[i] = ((u64)nx << 32) | w;
	});
	int iters = bits_for(n);
	for (int it = 0; it < iters; it++) {
		parallel_for<K_lr_jump>(c, n, PV_LAMBDA(size_t i) {
			u64 x = a[i];
			u32 nx = (u32)(x >> 32);
			if (nx != NONE) {
				u64 y = a[nx];
				x = (y & 0xFFFFFFFF00000000ull) | (u32)((u32)x + (u32)y);
			}
			b[i] = x;
		});
		std::swap(a, b);
	}
	parallel_for<K_lr_out>(c, n, PV_LAMBDA(size_t i) { rank_out[i] = (u32)a[i]; });
}

// ---------------------------------------------------------------------------------------------
// stream compaction of indices i in [0,n) with flag[i] != 0 -> out (ascending); returns the count
// ---------------------------------------------------------------------------------------------
PV_KERNEL(K_compact_scatter, 12);
inline u32 compact_flags(Ctx &c, const u32 *flags, size_t n, Buf<u32> &out) {
	if (n == 0) { out = Buf<u32>(c, 1); return 0; }
	Buf<u32> pos(c, n + 1);
	exclusive_scan(c, flags, pos.p, n, OpAddU32(), 0u);
	u32 last_pos = pos.read(n - 1);
	u32 last_flag;
	c.d2h(&last_flag, flags + (n - 1), 4);
	u32 cnt = last_pos + (last_flag ? 1u : 0u);
	out = Buf<u32>(c, cnt ? cnt : 1);
	u32 *o = out.p; const u32 *p = pos.p;
	parallel_for<K_compact_scatter>(c, n, PV_LAMBDA(size_t i) { if (flags[i]) o[p[i]] = (u32)i; });
	return cnt;
}

} // namespace pv
