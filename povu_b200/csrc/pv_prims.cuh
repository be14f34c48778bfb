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
static const int SCAN_THREADS = 512;
// items per thread: up to 128 bytes, at most 16 items (u32: 16, u64 / max-plus: 16, u128: 8).  Fat tiles keep the number
// of tiles -- and with it the length of the look-back chain when every tile of a mid-sized input is resident at once -- small.
template <class T> struct ScanCfg { static const int ITEMS = sizeof(T) >= 16 ? 8 : 16; static const int TILE = SCAN_THREADS * ITEMS; };

// word-wise helpers for generic (POD, multiple-of-4-byte) scan element types
template <class T> struct Words { static_assert(sizeof(T) % 4 == 0 && sizeof(T) <= 16, "scan element must be 4..16 bytes"); static const int N = sizeof(T) / 4; };
template <class T> __device__ __forceinline__ T shfl_up_t(T v, int d) {
	union { T t; u32 w[Words<T>::N]; } a, b; a.t = v;
#pragma unroll
	for (int k = 0; k < Words<T>::N; k++) b.w[k] = __shfl_up_sync(0xFFFFFFFFu, a.w[k], d);
	return b.t;
}
template <class T> __device__ __forceinline__ T shfl_down_t(T v, int d) {
	union { T t; u32 w[Words<T>::N]; } a, b; a.t = v;
#pragma unroll
	for (int k = 0; k < Words<T>::N; k++) b.w[k] = __shfl_down_sync(0xFFFFFFFFu, a.w[k], d);
	return b.t;
}
template <class T> __device__ __forceinline__ T shfl_t(T v, int src) {
	union { T t; u32 w[Words<T>::N]; } a, b; a.t = v;
#pragma unroll
	for (int k = 0; k < Words<T>::N; k++) b.w[k] = __shfl_sync(0xFFFFFFFFu, a.w[k], src);
	return b.t;
}
template <class T> __device__ __forceinline__ T ld_cg_t(const T *p) {   // L2 load (the value was published by another CTA)
	union { T t; u32 w[Words<T>::N]; } a;
#pragma unroll
	for (int k = 0; k < Words<T>::N; k++) a.w[k] = __ldcg((const u32 *)p + k);
	return a.t;
}
__device__ __forceinline__ void st_flag(u32 *p, u32 v) { *(volatile u32 *)p = v; }

// Single-pass scan with decoupled look-back: one launch, every element read once and written once.
// Tiles are handed out by an atomic ticket (a tile's predecessors have always started, so the look-back cannot
// dead-lock); a tile publishes its aggregate, warp 0 walks back 32 predecessors at a time until it meets an
// inclusive prefix, then publishes its own inclusive prefix.  Order of combination is preserved throughout, so
// non-commutative operators (function composition) are fine.  Flags carry the launch epoch: no clearing between launches.
template <class T, class Op, bool INCLUSIVE, bool VEC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_lb_kernel(const T *in, T *out, size_t n, Op op, T identity, u32 *flags, T *agg, T *incl,
								 u32 *ticket, u32 epoch, u32 ntiles) {
	constexpr int SCAN_ITEMS = ScanCfg<T>::ITEMS, SCAN_TILE = ScanCfg<T>::TILE;
	__shared__ T warpTot[SCAN_THREADS / 32];
	__shared__ T s_prefix;
	__shared__ u32 s_tile;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	pdl_wait();
	if (t == 0) { u32 k = atomicAdd(ticket, 1u); if (k == ntiles - 1) atomicExch(ticket, 0u); s_tile = k; }
	__syncthreads();
	const u32 tile = s_tile;
	const size_t base = (size_t)tile * SCAN_TILE + (size_t)t * SCAN_ITEMS;
	const bool full = (size_t)(tile + 1) * SCAN_TILE <= n;
	union Items { T v[SCAN_ITEMS]; uint4 q[sizeof(T) * SCAN_ITEMS / 16]; __device__ Items() {} } it;
	if (full) {
		if (VEC) {
			const uint4 *src = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
			for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) it.q[k] = src[k];
		} else {
#pragma unroll
			for (int k = 0; k < SCAN_ITEMS; k++) it.v[k] = in[base + k];
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; k++) it.v[k] = base + k < n ? in[base + k] : identity;
	}
	T acc = it.v[0];
#pragma unroll
	for (int k = 1; k < SCAN_ITEMS; k++) acc = op(acc, it.v[k]);
	T inc = acc;   // inclusive scan of the thread totals inside the warp
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { T y = shfl_up_t(inc, d); if (lane >= d) inc = op(y, inc); }
	if (lane == 31) warpTot[w] = inc;
	__syncthreads();
	if (w == 0) {
		T tot = warpTot[0];
#pragma unroll
		for (int i = 1; i < SCAN_THREADS / 32; i++) tot = op(tot, warpTot[i]);
		T prefix = identity;
		if (tile == 0) {
			if (lane == 0) { incl[0] = tot; __threadfence(); st_flag(flags, (epoch << 2) | 2u); }
		} else {
			if (lane == 0) { agg[tile] = tot; __threadfence(); st_flag(flags + tile, (epoch << 2) | 1u); }
			T run = identity; bool have = false;
			int cur = (int)tile - 1;
			while (true) {
				int idx = cur - lane;
				u32 st = 2u; T v = identity;
				if (idx >= 0) {
					u32 f;
					do { f = *(volatile u32 *)(flags + idx); } while ((f >> 2) != epoch);
					st = f & 3u;
					__threadfence();
					v = ld_cg_t(st == 2u ? incl + idx : agg + idx);
				}
				u32 inclMask = __ballot_sync(0xFFFFFFFFu, st == 2u);
				int lim = inclMask ? (__ffs((int)inclMask) - 1) : 31;   // nearest tile that already knows its inclusive prefix
				if (lane > lim) v = identity;
				// ordered product v[31] op ... op v[0] (higher lanes = earlier tiles)
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { T y = shfl_down_t(v, d); if (lane + d < 32) v = op(y, v); }
				v = shfl_t(v, 0);
				run = have ? op(v, run) : v; have = true;
				if (inclMask) break;
				cur -= 32;
			}
			prefix = run;
			if (lane == 0) { incl[tile] = op(prefix, tot); __threadfence(); st_flag(flags + tile, (epoch << 2) | 2u); }
		}
		if (lane == 0) s_prefix = prefix;
	}
	__syncthreads();
	T run = s_prefix;
	for (int i = 0; i < w; i++) run = op(run, warpTot[i]);
	T prevLane = shfl_up_t(inc, 1);
	if (lane > 0) run = op(run, prevLane);
	if (tile == 0 && w == 0 && lane == 0) run = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		if (INCLUSIVE) { run = op(run, it.v[k]); it.v[k] = run; }
		else { T x = it.v[k]; it.v[k] = run; run = op(run, x); }
	}
	if (full) {
		if (VEC) {
			uint4 *dst = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
			for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) dst[k] = it.q[k];
		} else {
#pragma unroll
			for (int k = 0; k < SCAN_ITEMS; k++) out[base + k] = it.v[k];
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] = it.v[k];
	}
}

// Mid-sized inputs (every tile resident at once: up to a few million elements, the regime of cfg2) are latency-bound in the
// look-back above: tile i walks back through its predecessors 32 at a time, one dependent round trip to L2 per window.  Here a
// tile publishes its aggregate and then fetches the aggregates of ALL its predecessors at once: thread t of the CTA owns a
// contiguous block of ceil(i / 512) predecessors (independent loads, combined in order), then an ordered warp and CTA reduction.
// One round trip instead of i / 32.  4-byte elements travel with their epoch in one 64-bit word (no fence, no second load).
template <class T> struct ScanPack { static const bool value = sizeof(T) == 4; };

// Bulk asynchronous copies (the TMA unit's 1-D path, UBLKCP in SASS): one elected thread moves a whole scan tile between
// global and shared memory; completion of the load is signalled on an mbarrier (transaction bytes), the store is a bulk group.
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, u32 bytes, u64 *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
	asm volatile("{\n\t.reg .pred p;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LAB_DONE;\n\tbra LAB_WAIT;\n\tLAB_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, u32 bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// TMA = true (16-byte aligned input and output): full tiles are staged through shared memory by bulk asynchronous copies --
// thread 0 arms an mbarrier with the tile's byte count and issues ONE cp.async.bulk for the tile, the CTA waits on the barrier
// and reads its items with 128-bit shared loads; the scanned tile goes back the same way (items to shared memory, one bulk
// store).  The partial last tile takes the register path.
template <class T, class Op, bool INCLUSIVE, bool VEC, bool TMA>
__global__ void __launch_bounds__(SCAN_THREADS) scan_flat_kernel(const T *in, T *out, size_t n, Op op, T identity, u32 *flags, T *agg, u64 *pack,
								   u32 *ticket, u32 epoch, u32 ntiles) {
	constexpr int SCAN_ITEMS = ScanCfg<T>::ITEMS, SCAN_TILE = ScanCfg<T>::TILE;
	extern __shared__ __align__(128) unsigned char scan_dsm[];      // TMA: the tile (SCAN_TILE * sizeof(T) bytes)
	__shared__ T warpTot[SCAN_THREADS / 32], warpPre[SCAN_THREADS / 32];
	__shared__ u32 warpHave[SCAN_THREADS / 32];
	__shared__ u32 s_tile;
	__shared__ __align__(8) u64 s_bar;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	if (TMA && t == 0) mbar_init(&s_bar, 1);
	pdl_wait();
	if (t == 0) { u32 k = atomicAdd(ticket, 1u); if (k == ntiles - 1) atomicExch(ticket, 0u); s_tile = k; }
	__syncthreads();
	const u32 tile = s_tile;
	const size_t base = (size_t)tile * SCAN_TILE + (size_t)t * SCAN_ITEMS;
	const bool full = (size_t)(tile + 1) * SCAN_TILE <= n;
	union Items { T v[SCAN_ITEMS]; uint4 q[sizeof(T) * SCAN_ITEMS / 16]; __device__ Items() {} } it;
	if (TMA && full) {
		constexpr u32 TILE_BYTES = (u32)(SCAN_TILE * sizeof(T));
		if (t == 0) { mbar_expect_tx(&s_bar, TILE_BYTES); bulk_g2s(scan_dsm, in + (size_t)tile * SCAN_TILE, TILE_BYTES, &s_bar); }
		mbar_wait(&s_bar, 0);
		const uint4 *src = reinterpret_cast<const uint4 *>(scan_dsm) + (size_t)t * (sizeof(T) * SCAN_ITEMS / 16);
#pragma unroll
		for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) it.q[k] = src[k];
	} else if (full) {
		if (VEC) {
			const uint4 *src = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
			for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) it.q[k] = src[k];
		} else {
#pragma unroll
			for (int k = 0; k < SCAN_ITEMS; k++) it.v[k] = in[base + k];
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; k++) it.v[k] = base + k < n ? in[base + k] : identity;
	}
	T acc = it.v[0];
#pragma unroll
	for (int k = 1; k < SCAN_ITEMS; k++) acc = op(acc, it.v[k]);
	T inc = acc;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { T y = shfl_up_t(inc, d); if (lane >= d) inc = op(y, inc); }
	if (lane == 31) warpTot[w] = inc;
	__syncthreads();
	if (w == 0 && lane == 0 && tile + 1 < ntiles) {   // publish the tile's aggregate first (nobody reads the last tile's)
		T tot = warpTot[0];
#pragma unroll
		for (int i = 1; i < SCAN_THREADS / 32; i++) tot = op(tot, warpTot[i]);
		if (ScanPack<T>::value) { union { T t; u32 u; } cv; cv.u = 0; cv.t = tot; *(volatile u64 *)(pack + tile) = ((u64)epoch << 32) | cv.u; }
		else { agg[tile] = tot; __threadfence(); st_flag(flags + tile, (epoch << 2) | 1u); }
	}
	// predecessors [0, tile): thread t combines [t * k, (t + 1) * k) in order -- the whole CTA fetches, so up to SCAN_THREADS
	// predecessors cost ONE round trip to L2 (the loads of a thread are issued back to back and only then tested and re-polled)
	{
		const u32 k = (tile + (u32)SCAN_THREADS - 1u) / (u32)SCAN_THREADS;
		const u32 j0 = (u32)t * k < tile ? (u32)t * k : tile, j1 = j0 + k < tile ? j0 + k : tile;
		T v = identity; bool have = false;   // (never combine with `identity` on the right: segmented operators only honour it on the left)
		for (u32 j = j0; j < j1; j += 4) {
			T x[4];
			if (ScanPack<T>::value) {
				u64 wd[4];
#pragma unroll
				for (int q = 0; q < 4; q++) if (j + q < j1) wd[q] = *(const volatile u64 *)(pack + j + q);
#pragma unroll
				for (int q = 0; q < 4; q++) if (j + q < j1) {
					while ((u32)(wd[q] >> 32) != epoch) wd[q] = *(const volatile u64 *)(pack + j + q);
					union { T t; u32 u; } cv; cv.u = (u32)wd[q]; x[q] = cv.t;
				}
			} else {
				u32 f[4];
#pragma unroll
				for (int q = 0; q < 4; q++) if (j + q < j1) f[q] = *(const volatile u32 *)(flags + j + q);
#pragma unroll
				for (int q = 0; q < 4; q++) if (j + q < j1) { while ((f[q] >> 2) != epoch) f[q] = *(const volatile u32 *)(flags + j + q); }
				__threadfence();      // one fence for the (up to) four values whose flags were seen
#pragma unroll
				for (int q = 0; q < 4; q++) if (j + q < j1) x[q] = ld_cg_t(agg + j + q);
			}
#pragma unroll
			for (int q = 0; q < 4; q++) if (j + q < j1) { v = have ? op(v, x[q]) : x[q]; have = true; }
		}
		// ordered product over the lanes that hold something (lower lanes = earlier tiles), then over the warps
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			T y = shfl_up_t(v, d); bool hy = __shfl_up_sync(0xFFFFFFFFu, have ? 1 : 0, d) != 0;
			if (lane >= d && hy) { v = have ? op(y, v) : y; have = true; }
		}
		if (lane == 31) { warpPre[w] = v; warpHave[w] = have ? 1u : 0u; }
	}
	__syncthreads();
	T run = identity;
	{
		bool hr = false;
#pragma unroll
		for (int i = 0; i < SCAN_THREADS / 32; i++) if (warpHave[i]) { run = hr ? op(run, warpPre[i]) : warpPre[i]; hr = true; }
	}
	for (int i = 0; i < w; i++) run = op(run, warpTot[i]);
	T prevLane = shfl_up_t(inc, 1);
	if (lane > 0) run = op(run, prevLane);
	if (tile == 0 && w == 0 && lane == 0) run = identity;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; k++) {
		if (INCLUSIVE) { run = op(run, it.v[k]); it.v[k] = run; }
		else { T x = it.v[k]; it.v[k] = run; run = op(run, x); }
	}
	if (TMA && full) {
		constexpr u32 TILE_BYTES = (u32)(SCAN_TILE * sizeof(T));
		uint4 *dsts = reinterpret_cast<uint4 *>(scan_dsm) + (size_t)t * (sizeof(T) * SCAN_ITEMS / 16);
#pragma unroll
		for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) dsts[k] = it.q[k];
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the generic-proxy writes above become visible to the bulk copy
		__syncthreads();
		if (t == 0) bulk_s2g(out + (size_t)tile * SCAN_TILE, scan_dsm, TILE_BYTES);
	} else if (full) {
		if (VEC) {
			uint4 *dst = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
			for (int k = 0; k < (int)(sizeof(T) * SCAN_ITEMS / 16); k++) dst[k] = it.q[k];
		} else {
#pragma unroll
			for (int k = 0; k < SCAN_ITEMS; k++) out[base + k] = it.v[k];
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) out[base + k] = it.v[k];
	}
}

template <class T, class Op, bool INCLUSIVE>
inline void scan_impl(Ctx &c, const T *in, T *out, size_t n, Op op, T identity) {
	if (n == 0) return;
	const size_t SCAN_TILE = ScanCfg<T>::TILE;
	size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (ntiles >= 0x7FFFFFFFull) throw std::runtime_error("scan: input too large");
	c.lb_reserve(ntiles);
	u32 epoch = c.next_epoch();
	T *agg = (T *)c.lb_vals, *incl = (T *)(c.lb_vals + c.lb_cap * 16);
	bool vec = (((uintptr_t)in | (uintptr_t)out) & 15u) == 0;
	static const u32 flat_max = getenv("POVU_B200_SCAN_FLAT_MAX") ? (u32)strtoul(getenv("POVU_B200_SCAN_FLAT_MAX"), nullptr, 10) : 1024u;
	if (ntiles <= flat_max) {   // every predecessor's aggregate is fetched at once (see scan_flat_kernel)
		c.prof_begin("scan_flat", n, n * 2 * sizeof(T));
		static const bool no_tma = getenv("POVU_B200_SCAN_NO_TMA") != nullptr;
		if (vec && !no_tma && ntiles > 1) {
			static bool attr_set = false;      // (per instantiation: this is a function template)
			const size_t smem = SCAN_TILE * sizeof(T);
			if (!attr_set) { PV_CUDA_CHECK(cudaFuncSetAttribute(scan_flat_kernel<T, Op, INCLUSIVE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
			PV_CUDA_CHECK(launch_pdl(scan_flat_kernel<T, Op, INCLUSIVE, true, true>, dim3((unsigned)ntiles), dim3(SCAN_THREADS), smem, c.stream, in, out, n, op, identity, c.lb_flags, agg, c.lb_pack, c.lb_ticket, epoch, (u32)ntiles));
		}
		else if (vec) PV_CUDA_CHECK(launch_pdl(scan_flat_kernel<T, Op, INCLUSIVE, true, false>, dim3((unsigned)ntiles), dim3(SCAN_THREADS), 0, c.stream, in, out, n, op, identity, c.lb_flags, agg, c.lb_pack, c.lb_ticket, epoch, (u32)ntiles));
		else PV_CUDA_CHECK(launch_pdl(scan_flat_kernel<T, Op, INCLUSIVE, false, false>, dim3((unsigned)ntiles), dim3(SCAN_THREADS), 0, c.stream, in, out, n, op, identity, c.lb_flags, agg, c.lb_pack, c.lb_ticket, epoch, (u32)ntiles));
		c.prof_end();
		c.launches++;
		PV_CUDA_CHECK(cudaGetLastError());
		return;
	}
	c.prof_begin("scan_lb", n, n * 2 * sizeof(T));
	if (vec) PV_CUDA_CHECK(launch_pdl(scan_lb_kernel<T, Op, INCLUSIVE, true>, dim3((unsigned)ntiles), dim3(SCAN_THREADS), 0, c.stream, in, out, n, op, identity, c.lb_flags, agg, incl, c.lb_ticket, epoch, (u32)ntiles));
	else PV_CUDA_CHECK(launch_pdl(scan_lb_kernel<T, Op, INCLUSIVE, false>, dim3((unsigned)ntiles), dim3(SCAN_THREADS), 0, c.stream, in, out, n, op, identity, c.lb_flags, agg, incl, c.lb_ticket, epoch, (u32)ntiles));
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
// stable LSD radix sort of (u32 key, u32 value) pairs on key bits [0, bits), 8 bits per pass, "onesweep" style:
//   * ONE histogram kernel reads the keys once and counts the digits of every pass (thread-private run-length
//     accumulation -> shared-memory atomics -> global atomics);
//   * ONE kernel per pass: a CTA ranks its tile (warp-private __match_any_sync ranking, in index order, so the sort
//     is stable by construction), publishes its digit counts, obtains its global offsets by decoupled look-back over
//     the tiles before it (one thread per digit), stages the tile in digit order in shared memory and writes it out
//     in coalesced runs.  Every pair is read once and written once per pass; nothing else touches HBM.
// Tiles are handed out by an atomic ticket; status words carry the launch epoch (no clearing between launches).
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
static const int RS_THREADS = 256;
static const int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_all_kernel(const u32 *keys, size_t n, u32 *ghist, int passes, int vec, int begin) {
	__shared__ u32 hist[4][256];
	for (int i = threadIdx.x; i < 4 * 256; i += RS_THREADS) (&hist[0][0])[i] = 0;
	__syncthreads();
	const size_t chunk = 16;   // consecutive keys per thread per iteration: runs of equal digits collapse into one atomic
	size_t nchunks = (n + chunk - 1) / chunk;
	for (size_t cidx = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; cidx < nchunks; cidx += (size_t)gridDim.x * RS_THREADS) {
		size_t base = cidx * chunk;
		u32 k[16]; int m = 16;
		if (vec && base + 16 <= n) {
			const uint4 *src = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
			for (int j = 0; j < 4; j++) { uint4 q = src[j]; k[4 * j] = q.x; k[4 * j + 1] = q.y; k[4 * j + 2] = q.z; k[4 * j + 3] = q.w; }
		} else {
			m = (int)(n - base < 16 ? n - base : 16);
#pragma unroll
			for (int j = 0; j < 16; j++) k[j] = j < m ? keys[base + j] : 0u;
		}
		for (int p = 0; p < passes; p++) {
			int sh = begin + 8 * p;
			u32 cur = (k[0] >> sh) & 255u, cnt = 1;
#pragma unroll
			for (int j = 1; j < 16; j++) {
				if (j < m) {
					u32 d = (k[j] >> sh) & 255u;
					if (d == cur) cnt++;
					else { atomicAdd(&hist[p][cur], cnt); cur = d; cnt = 1; }
				}
			}
			atomicAdd(&hist[p][cur], cnt);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < passes * 256; i += RS_THREADS) { u32 v = (&hist[0][0])[i]; if (v) atomicAdd(ghist + i, v); }
}

__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32 *warpSum /*[8] shared*/) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	u32 inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += y; }
	if (lane == 31) warpSum[w] = inc;
	__syncthreads();
	u32 base = 0;
#pragma unroll
	for (int i = 0; i < RS_WARPS; i++) if (i < w) base += warpSum[i];
	__syncthreads();
	return base + inc - v;
}

template <int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, 4) rs_onesweep_kernel(const u32 *kin, const u32 *vin, u32 *kout, u32 *vout, size_t n, int shift,
								      const u32 *ghist, u64 *status, u32 *ticket, u32 epoch, u32 ntiles) {
	constexpr int TILE = RS_THREADS * ITEMS;
	__shared__ u32 sKeys[TILE];
	__shared__ u32 sVals[TILE];
	__shared__ u32 warpCnt[RS_WARPS][256];
	__shared__ u32 digStart[256];
	__shared__ u32 gOff[256];
	__shared__ u32 warpSum[RS_WARPS];
	__shared__ u32 s_tile;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	if (t == 0) { u32 k = atomicAdd(ticket, 1u); if (k == ntiles - 1) atomicExch(ticket, 0u); s_tile = k; }
#pragma unroll
	for (int i = 0; i < RS_WARPS; i++) warpCnt[i][t] = 0;
	__syncthreads();
	const u32 tile = s_tile;
	const size_t tileBase = (size_t)tile * TILE;
	const u32 tileCount = (u32)(n - tileBase < (size_t)TILE ? n - tileBase : (size_t)TILE);
	const size_t wbase = tileBase + (size_t)w * (32 * ITEMS) + lane;
	const bool full = tileCount == (u32)TILE;
	u32 key[ITEMS], pos[ITEMS];
	if (full) {
#pragma unroll
		for (int k = 0; k < ITEMS; k++) key[k] = kin[wbase + (size_t)k * 32];
	} else {
#pragma unroll
		for (int k = 0; k < ITEMS; k++) { size_t i = wbase + (size_t)k * 32; key[k] = i < n ? kin[i] : 0u; }
	}
	// warp-private ranking in index order (stable).  All the matches are issued first, then the leaders' counter
	// updates (shared-memory atomics; __syncwarp keeps successive rounds ordered), then the broadcasts: the three
	// latencies overlap across the ITEMS rounds instead of adding up.
	{
		const u32 ltMask = (1u << lane) - 1u;
		u32 peers[ITEMS];
#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			bool valid = full || (wbase + (size_t)k * 32 < n);
			u32 d = valid ? ((key[k] >> shift) & 255u) : 256u + (u32)lane;   // invalid lanes get unique pseudo-digits
			peers[k] = __match_any_sync(0xFFFFFFFFu, d);
		}
#pragma unroll
		for (int k = 0; k < ITEMS; k++) {
			bool valid = full || (wbase + (size_t)k * 32 < n);
			u32 c0 = 0;
			if (valid && (peers[k] & ltMask) == 0) c0 = atomicAdd(&warpCnt[w][(key[k] >> shift) & 255u], (u32)__popc(peers[k]));
			__syncwarp();
			pos[k] = c0;
		}
#pragma unroll
		for (int k = 0; k < ITEMS; k++) pos[k] = __shfl_sync(0xFFFFFFFFu, pos[k], __ffs((int)peers[k]) - 1) + __popc(peers[k] & ltMask);
	}
	__syncthreads();
	// thread t owns digit t: warp counts -> warp offsets, tile count
	u32 cnt = 0;
#pragma unroll
	for (int i = 0; i < RS_WARPS; i++) { u32 x = warpCnt[i][t]; warpCnt[i][t] = cnt; cnt += x; }
	u64 *myStatus = status + (size_t)tile * 256 + t;
	*(volatile u64 *)myStatus = ((u64)epoch << 34) | ((u64)(tile == 0 ? 2u : 1u) << 32) | (u64)cnt;
	u32 ds = block_excl_scan_256(cnt, warpSum);
	u32 gbase = block_excl_scan_256(ghist[t], warpSum);
	u32 excl = 0;
	if (tile > 0) {
		// look-back, four predecessors' status words in flight at a time (the chain is latency-bound: when every tile of a
		// mid-sized input is resident at once, tile i walks ~sqrt(2i) predecessors before it meets an inclusive prefix)
		const u64 done = ((u64)epoch << 34) | ((u64)2u << 32);
		int i = (int)tile - 1;
		bool fin = false;
		while (!fin) {
			u64 sw[4];
#pragma unroll
			for (int q = 0; q < 4; q++) sw[q] = i - q >= 0 ? *(const volatile u64 *)(status + (size_t)(i - q) * 256 + t) : done;
			int used = 0;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				if (!fin && used == q && (u32)(sw[q] >> 34) == epoch) {
					excl += (u32)sw[q]; used = q + 1;
					if (((sw[q] >> 32) & 3u) == 2u) fin = true;
				}
			}
			i -= used;   // the first word that was not ready yet is polled again
		}
		*(volatile u64 *)myStatus = ((u64)epoch << 34) | ((u64)2u << 32) | (u64)(excl + cnt);
	}
	digStart[t] = ds;
	gOff[t] = gbase + excl - ds;
	__syncthreads();
	// stage the tile in digit order, keys first, then the values (loaded only now: fewer live registers -> 4 CTAs / SM)
#pragma unroll
	for (int k = 0; k < ITEMS; k++) {
		bool valid = full || (wbase + (size_t)k * 32 < n);
		if (valid) {
			u32 d = (key[k] >> shift) & 255u;
			pos[k] += digStart[d] + warpCnt[w][d];
			sKeys[pos[k]] = key[k];
		}
	}
	if (full) {
#pragma unroll
		for (int k = 0; k < ITEMS; k++) key[k] = vin[wbase + (size_t)k * 32];
#pragma unroll
		for (int k = 0; k < ITEMS; k++) sVals[pos[k]] = key[k];
	} else {
#pragma unroll
		for (int k = 0; k < ITEMS; k++) { size_t i = wbase + (size_t)k * 32; if (i < n) sVals[pos[k]] = vin[i]; }
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < ITEMS; k++) {
		u32 j = (u32)t + (u32)k * RS_THREADS;
		if (j < tileCount) {
			u32 kk = sKeys[j];
			u32 dst = gOff[(kk >> shift) & 255u] + j;
			kout[dst] = kk; vout[dst] = sVals[j];
		}
	}
}
#endif

// Sorts (keys, vals) by key bits [begin, begin + bits) -- the bits below `begin` ride along as payload; on return keys/vals may
// have been SWAPPED with the internal double buffers (the Buf objects always own the sorted data; raw pointers taken before the
// call are stale).
inline void radix_sort_pairs(Ctx &c, Buf<u32> &keys, Buf<u32> &vals, size_t n, int bits, int begin = 0) {
	if (n <= 1 || bits <= 0) return;
#ifndef PV_EMU
	if (n >= (1ull << 32)) throw std::runtime_error("radix sort: input too large");
	if (begin + bits > 32) bits = 32 - begin;
	const int passes = (bits + 7) / 8;
	Buf<u32> k2(c, keys.n), v2(c, vals.n);
	Buf<u32> ghist(c, 4 * 256);
	ghist.zero();
	bool small = n < (size_t)c.sm_count * 4 * 4096;
	const size_t tile = small ? RS_THREADS * 8 : RS_THREADS * 16;
	size_t ntiles = (n + tile - 1) / tile;
	c.rs_reserve(ntiles * 256);
	{
		size_t nchunks = (n + 15) / 16;
		size_t blocks = (nchunks + RS_THREADS - 1) / RS_THREADS;
		size_t cap = (size_t)c.sm_count * 8;
		if (blocks > cap) blocks = cap;
		c.prof_begin("rs_hist", n, n * 4);
		rs_hist_all_kernel<<<(unsigned)blocks, RS_THREADS, 0, c.stream>>>(keys.p, n, ghist.p, passes, (((uintptr_t)keys.p) & 15u) == 0 ? 1 : 0, begin);
		c.prof_end();
		c.launches++;
	}
	for (int p = 0; p < passes; p++) {
		u32 epoch = c.next_epoch();
		c.prof_begin("rs_onesweep", n, n * 16);
		if (small) rs_onesweep_kernel<8><<<(unsigned)ntiles, RS_THREADS, 0, c.stream>>>(keys.p, vals.p, k2.p, v2.p, n, begin + 8 * p, ghist.p + 256 * p, c.rs_status, c.lb_ticket + 1, epoch, (u32)ntiles);
		else rs_onesweep_kernel<16><<<(unsigned)ntiles, RS_THREADS, 0, c.stream>>>(keys.p, vals.p, k2.p, v2.p, n, begin + 8 * p, ghist.p + 256 * p, c.rs_status, c.lb_ticket + 1, epoch, (u32)ntiles);
		c.prof_end();
		c.launches++;
		PV_CUDA_CHECK(cudaGetLastError());
		std::swap(keys, k2); std::swap(vals, v2);
	}
#else
	c.launches++;
	std::vector<u32> idx(n);
	std::iota(idx.begin(), idx.end(), 0u);
	u32 mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
	u32 *kp = keys.p, *vp = vals.p;
	std::stable_sort(idx.begin(), idx.end(), [&](u32 a, u32 b) { return ((kp[a] >> begin) & mask) < ((kp[b] >> begin) & mask); });
	// like the device path with an odd number of passes: the result lives in fresh buffers, the old ones are poisoned
	Buf<u32> k2(c, keys.n), v2(c, vals.n);
	for (size_t i = 0; i < n; i++) { k2.p[i] = kp[idx[i]]; v2.p[i] = vp[idx[i]]; }
	memset(kp, 0xCD, n * 4); memset(vp, 0xCD, n * 4);
	std::swap(keys, k2); std::swap(vals, v2);
#endif
}

// pointer flavour for cold call sites: sorts a copy and writes the result back to keys/vals
inline void radix_sort_pairs_inplace(Ctx &c, u32 *keys, u32 *vals, size_t n, int bits) {
	if (n <= 1 || bits <= 0) return;
	Buf<u32> k(c, n), v(c, n);
	c.d2d(k.p, keys, n * 4); c.d2d(v.p, vals, n * 4);
	radix_sort_pairs(c, k, v, n, bits);
	c.d2d(keys, k.p, n * 4); c.d2d(vals, v.p, n * 4);
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

// Two levels of blocking so that a build is ONE launch (the sparse table over all block minima took one launch per five levels:
// 12 launches for the four RMQs of a cfg2 decompose, all latency):
//   block      = 32 elements: prefix / suffix minima inside the block (pre, suf)
//   superblock = 32 blocks (1024 elements, one CTA): per block a 6-level table of minima over the next 1, 2, 4, 8, 16, 32 blocks of its
//                superblock (bt), prefix / suffix minima of the block minima inside the superblock (pre2, suf2)
//   a sparse table over the superblock minima (st), built by the last CTA to finish (n / 1024 entries per level).
struct RmqView {
	const u32 *a; const u32 *pre; const u32 *suf; const u32 *bt; const u32 *pre2; const u32 *suf2; const u32 *st; u32 n; u32 nb; u32 ns; u32 levels;
	PV_DEV u32 query(u32 l, u32 r) const {
#ifdef PV_EMU
		if (l > r || r >= n) { fprintf(stderr, "RmqView::query(%u, %u) outside [0, %u)\n", l, r, n); abort(); }
#endif
		u32 bl = l >> 5, br = r >> 5;
		if (bl == br) {
			u32 m = a[l];
			for (u32 i = l + 1; i <= r; i++) { u32 x = a[i]; m = x < m ? x : m; }
			return m;
		}
		u32 m = suf[l]; u32 x = pre[r]; m = x < m ? x : m;
		if (br > bl + 1) {
			const u32 p = bl + 1, q = br - 1;      // whole blocks in between
			const u32 sp = p >> 5, sq = q >> 5;
			if (sp == sq) {
				u32 k = 31u - pv_clz(q - p + 1);
				x = bt[(size_t)k * nb + p]; m = x < m ? x : m;
				x = bt[(size_t)k * nb + (q + 1 - (1u << k))]; m = x < m ? x : m;
			} else {
				x = suf2[p]; m = x < m ? x : m;
				x = pre2[q]; m = x < m ? x : m;
				if (sq > sp + 1) {
					u32 k = 31u - pv_clz(sq - sp - 1);
					const u32 *row = st + (size_t)k * ns;
					x = row[sp + 1]; m = x < m ? x : m;
					x = row[sq - (1u << k)]; m = x < m ? x : m;
				}
			}
		}
		return m;
	}
	PV_DEV u32 scan_block_down(u32 blk, u32 bound) const {   // largest index of block blk with a value below bound (the block holds one)
		u32 s = blk << 5; u32 i = s + 31u < n ? s + 31u : n - 1;
		for (;; i--) { if (a[i] < bound) return i; if (i == s) break; }
		return NONE;   // unreachable when the tables are consistent
	}
	// largest index j <= r with a[j] < bound, or NONE ("previous smaller value"): r's block, then the earlier blocks of its
	// superblock, then whole superblocks by descending the sparse table, then the hit superblock's blocks.
	PV_DEV u32 prev_smaller(u32 r, u32 bound) const {
		u32 j = r;
		while (true) {
			if (a[j] < bound) return j;
			if ((j & 31u) == 0) break;
			j--;
		}
		const u32 b = r >> 5, sb = b >> 5;
		for (u32 blk = b; blk-- > (sb << 5);) if (bt[blk] < bound) return scan_block_down(blk, bound);
		u32 pos = sb;   // superblocks [pos, sb] hold no candidate
		for (u32 k = levels; k-- > 0;) {
			u32 span = 1u << k;
			if (pos >= span && st[(size_t)k * ns + (pos - span)] >= bound) pos -= span;
		}
		if (pos == 0) return NONE;
		const u32 s = pos - 1;
		u32 top = (s << 5) + 31u; if (top >= nb) top = nb - 1;
		for (u32 blk = top + 1; blk-- > (s << 5);) if (bt[blk] < bound) return scan_block_down(blk, bound);
		return NONE;   // unreachable when the tables are consistent
	}
};

struct Rmq {
	Buf<u32> pre, suf, bt, pre2, suf2, st;
	u32 n = 0, nb = 0, ns = 0, levels = 0;
	const u32 *a = nullptr;
	RmqView view() const { return RmqView{a, pre.p, suf.p, bt.p, pre2.p, suf2.p, st.p, n, nb, ns, levels}; }
};

PV_KERNEL(K_rmq_blocks, 12);   // a in, pre + suf out (+ the block tables: 8 words per 32 elements)
PV_KERNEL(K_rmq_level, 8);

#ifndef PV_EMU
__global__ void __launch_bounds__(1024) rmq_build_kernel(const u32 *a, u32 *pre, u32 *suf, u32 *bt, u32 *pre2, u32 *suf2, u32 *st, u32 n, u32 nb, u32 ns, u32 levels, u32 *ticket) {
	__shared__ u32 bm[32];
	__shared__ u32 s_last;
	const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5, s = blockIdx.x;
	const u32 b = (s << 5) + w;
	pdl_wait();
	{
		u32 i = (b << 5) + lane;
		u32 x = i < n ? a[i] : NONE;
		u32 p = x, q = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 y = __shfl_up_sync(0xFFFFFFFFu, p, d); if (lane >= (u32)d && y < p) p = y;
			u32 z = __shfl_down_sync(0xFFFFFFFFu, q, d); if (lane + d < 32u && z < q) q = z;
		}
		if (i < n) { pre[i] = p; suf[i] = q; }
		if (lane == 0) bm[w] = q;      // NONE for blocks past the end
	}
	__syncthreads();
	if (w < 6) {   // warp k: level k of the block table (minimum over the next 2^k blocks of this superblock)
		const u32 k = w, span = 1u << k;
		u32 m = NONE;
		for (u32 i = 0; i < span && lane + i < 32u; i++) { u32 x = bm[lane + i]; m = x < m ? x : m; }
		if ((s << 5) + lane < nb) bt[(size_t)k * nb + (s << 5) + lane] = m;
	} else if (w == 6) {
		u32 x = bm[lane], p = x, q = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			u32 y = __shfl_up_sync(0xFFFFFFFFu, p, d); if (lane >= (u32)d && y < p) p = y;
			u32 z = __shfl_down_sync(0xFFFFFFFFu, q, d); if (lane + d < 32u && z < q) q = z;
		}
		if ((s << 5) + lane < nb) { pre2[(s << 5) + lane] = p; suf2[(s << 5) + lane] = q; }
		if (lane == 0) { st[s] = q; __threadfence(); u32 k = atomicAdd(ticket, 1u); s_last = (k == gridDim.x - 1) ? 1u : 0u; if (s_last) atomicExch(ticket, 0u); }
	}
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	// the last CTA: sparse table over the superblock minima, level by level
	for (u32 k = 1; k < levels; k++) {
		const u32 half = 1u << (k - 1);
		const u32 *src = st + (size_t)(k - 1) * ns; u32 *dst = st + (size_t)k * ns;
		for (u32 i = threadIdx.x; i < ns; i += 1024u) {
			u32 x = __ldcg(src + i), y = i + half < ns ? __ldcg(src + i + half) : NONE;
			dst[i] = x < y ? x : y;
		}
		__syncthreads();
	}
}
#endif

inline void rmq_build(Ctx &c, Rmq &r, const u32 *a, u32 n) {
	r.a = a; r.n = n; r.nb = (n + 31) / 32; r.ns = (r.nb + 31) / 32;
	r.levels = 1; while ((1u << r.levels) <= r.ns) r.levels++;
	r.pre = Buf<u32>(c, n); r.suf = Buf<u32>(c, n); r.bt = Buf<u32>(c, (size_t)6 * r.nb); r.pre2 = Buf<u32>(c, r.nb); r.suf2 = Buf<u32>(c, r.nb);
	r.st = Buf<u32>(c, (size_t)r.levels * r.ns);
	u32 *pre = r.pre.p, *suf = r.suf.p, *bt = r.bt.p, *pre2 = r.pre2.p, *suf2 = r.suf2.p, *st = r.st.p; const u32 nb = r.nb, ns = r.ns, levels = r.levels;
#ifndef PV_EMU
	c.lb_reserve(1);
	c.prof_begin(K_rmq_blocks::name(), n, (u64)n * K_rmq_blocks::bytes);
	PV_CUDA_CHECK(launch_pdl(rmq_build_kernel, dim3(ns), dim3(1024), 0, c.stream, a, pre, suf, bt, pre2, suf2, st, n, nb, ns, levels, c.lb_ticket + 2));
	c.prof_end();
	c.launches++;
	PV_CUDA_CHECK(cudaGetLastError());
#else
	parallel_for<K_rmq_blocks>(c, ns, PV_LAMBDA(size_t s) {
		u32 bm[32];
		for (u32 w = 0; w < 32; w++) {
			u32 b = ((u32)s << 5) + w; u32 lo = b << 5, hi = lo + 32 < n ? lo + 32 : n;
			u32 m = NONE;
			for (u32 i = lo; i < hi && i >= lo; i++) { u32 x = a[i]; m = x < m ? x : m; pre[i] = m; }
			m = NONE;
			if (lo < n) for (u32 i = hi; i-- > lo;) { u32 x = a[i]; m = x < m ? x : m; suf[i] = m; }
			bm[w] = m;
		}
		for (u32 w = 0; w < 32; w++) {
			u32 b = ((u32)s << 5) + w;
			if (b >= nb) break;
			for (u32 k = 0; k < 6; k++) { u32 m = NONE; for (u32 i = 0; i < (1u << k) && w + i < 32; i++) m = bm[w + i] < m ? bm[w + i] : m; bt[(size_t)k * nb + b] = m; }
			u32 p = NONE; for (u32 i = 0; i <= w; i++) p = bm[i] < p ? bm[i] : p;
			u32 q = NONE; for (u32 i = w; i < 32; i++) q = bm[i] < q ? bm[i] : q;
			pre2[b] = p; suf2[b] = q;
			if (w == 0) st[s] = q;
		}
	});
	parallel_for<K_rmq_level>(c, 1, PV_LAMBDA(size_t) {
		for (u32 k = 1; k < levels; k++)
			for (u32 i = 0; i < ns; i++) { u32 x = st[(size_t)(k - 1) * ns + i]; u32 j = i + (1u << (k - 1)); u32 y = j < ns ? st[(size_t)(k - 1) * ns + j] : NONE; st[(size_t)k * ns + i] = x < y ? x : y; }
	});
#endif
}

// ---------------------------------------------------------------------------------------------
// list ranking (Wyllie pointer jumping).  Input: next[i] (NONE terminates), weight[i].
// Output: rank[i] = sum of weights of the elements strictly AFTER i on its list ... plus nothing;
// i.e. suffix sums excluding i.  Packed as (next << 32 | acc) so every hop is one 8-byte load.
// ---------------------------------------------------------------------------------------------
PV_KERNEL(K_lr_init, 16);
PV_KERNEL(K_lr_jump, 24);      // a[i], a[next] in, b[i] out (8 B each)
PV_KERNEL(K_lr_out, 12);

inline void list_rank_wyllie(Ctx &c, const u32 *next, const u32 *weight, u32 *rank_out, size_t n) {
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
	const int iters = bits_for(n);
	// every doubling step of the list in ONE cooperative launch (sweep `it` reads buffer it & 1, writes the other)
	parallel_iter<K_lr_jump>(c, n, iters, PV_LAMBDA(size_t i, int it) {
		const u64 *src = (it & 1) ? b : a; u64 *dst = (it & 1) ? a : b;
		u64 x = ld_cg(src + i);
		u32 nx = (u32)(x >> 32);
		if (nx != NONE) {
			u64 y = ld_cg(src + nx);
			x = (y & 0xFFFFFFFF00000000ull) | (u32)((u32)x + (u32)y);
		}
		dst[i] = x;
	});
	if (iters & 1) std::swap(a, b);
	parallel_for<K_lr_out>(c, n, PV_LAMBDA(size_t i) { rank_out[i] = (u32)a[i]; });
}

inline u32 compact_flags(Ctx &c, const u32 *flags, size_t n, Buf<u32> &out, Buf<u32> *keep_pos = nullptr, MultiRead *mr = nullptr, u32 *cnt_dst = nullptr);

// Block-local list ranking.  The lists ranked here are Euler tours of DFS trees whose vertices are numbered along the chain
// (segment ids follow the backbone), so a list element's successor is almost always a few indices away.  Every CTA takes 2048
// consecutive elements, cuts the list where it leaves the block and ranks the pieces inside shared memory (11 pointer-jumping
// steps, no global traffic); what is left for the global Wyllie pass is one node per piece -- a few thousand instead of four
// million at cfg2.  Without locality every element is its own piece and this degenerates to the plain pass plus one sweep.
PV_KERNEL(K_lr_local, 20);     // next + weight in, tail + partial sum + flag out
PV_KERNEL(K_lr_reduce, 24);
PV_KERNEL(K_lr_expand, 16);
#ifndef PV_EMU
static const int LR_BLOCK = 2048, LR_THREADS = 512;
__global__ void __launch_bounds__(LR_THREADS) lr_local_kernel(const u32 *next, const u32 *weight, u32 n, u32 *tailOf, u32 *sumTo, u32 *isTail) {
	__shared__ u32 P[2][LR_BLOCK];      // next element of the piece inside the block (local index), NONE = none
	__shared__ u32 A[2][LR_BLOCK];      // weights of the elements in (l, P(l)]
	pdl_wait();
	const u32 base = blockIdx.x * LR_BLOCK;
	for (u32 l = threadIdx.x; l < LR_BLOCK; l += LR_THREADS) {
		const u32 i = base + l;
		u32 p = NONE, a = 0;
		if (i < n) { u32 nx = next[i]; if (nx != NONE && nx >= base && nx - base < LR_BLOCK) { p = nx - base; a = weight[nx]; } }
		P[0][l] = p; A[0][l] = a;
	}
	__syncthreads();
	int cur = 0;
#pragma unroll 1
	for (int it = 0; it < 11; it++) {      // 2^11 = LR_BLOCK
		for (u32 l = threadIdx.x; l < LR_BLOCK; l += LR_THREADS) {
			u32 q = P[cur][l], a = A[cur][l];
			if (q != NONE) { u32 qq = P[cur][q]; if (qq != NONE) { a += A[cur][q]; q = qq; } }      // never jump past the piece's last element
			P[cur ^ 1][l] = q; A[cur ^ 1][l] = a;
		}
		__syncthreads();
		cur ^= 1;
	}
	for (u32 l = threadIdx.x; l < LR_BLOCK; l += LR_THREADS) {
		const u32 i = base + l;
		if (i >= n) continue;
		const u32 q = P[cur][l];
		tailOf[i] = q == NONE ? i : base + q; sumTo[i] = q == NONE ? 0u : A[cur][l]; isTail[i] = q == NONE ? 1u : 0u;
	}
}
#endif

inline void list_rank_suffix(Ctx &c, const u32 *next, const u32 *weight, u32 *rank_out, size_t n) {
	if (n == 0) return;
#ifdef PV_EMU
	list_rank_wyllie(c, next, weight, rank_out, n);
#else
	static const bool plain = getenv("POVU_B200_LR_PLAIN") != nullptr;
	if (plain || n < 4 * (size_t)LR_BLOCK || n >= 0xFFFFFFF0ull) { list_rank_wyllie(c, next, weight, rank_out, n); return; }
	Buf<u32> tailOf(c, n), sumTo(c, n), isTail(c, n), tails, tid;
	c.prof_begin(K_lr_local::name(), n, n * K_lr_local::bytes);
	PV_CUDA_CHECK(launch_pdl(lr_local_kernel, dim3((unsigned)((n + LR_BLOCK - 1) / LR_BLOCK)), dim3(LR_THREADS), 0, c.stream, next, weight, (u32)n, tailOf.p, sumTo.p, isTail.p));
	c.prof_end();
	c.launches++;
	const u32 m = compact_flags(c, isTail.p, n, tails, &tid);      // tails ascending; tid[i] = dense id of tail i
	Buf<u32> rnext(c, m), rw(c, m), rrank(c, m);
	rw.zero();
	u32 *rnext_ = rnext.p, *rw_ = rw.p; const u32 *tails_ = tails.p, *tid_ = tid.p, *tailOf_ = tailOf.p, *sumTo_ = sumTo.p;
	parallel_for<K_lr_reduce>(c, m, PV_LAMBDA(size_t k) {
		const u32 t = tails_[k]; const u32 nx = next[t];
		if (nx == NONE) { rnext_[k] = NONE; return; }
		const u32 k2 = tid_[tailOf_[nx]];            // the piece that starts at nx: its weight = every element from nx to its tail
		rnext_[k] = k2; rw_[k2] = weight[nx] + sumTo_[nx];
	});
	list_rank_wyllie(c, rnext.p, rw.p, rrank.p, m);
	const u32 *rrank_ = rrank.p;
	parallel_for<K_lr_expand>(c, n, PV_LAMBDA(size_t i) { rank_out[i] = sumTo_[i] + rrank_[tid_[tailOf_[i]]]; });
#endif
}

// ---------------------------------------------------------------------------------------------
// stream compaction of indices i in [0,n) with flag[i] != 0 -> out (ascending); returns the count
// ---------------------------------------------------------------------------------------------
PV_KERNEL(K_compact_scatter, 12);
// keep_pos (optional) receives the exclusive scan of the flags (n + 1 entries): rank of every index among the kept ones
// mr (optional): the count is not read here but queued on the caller's MultiRead into *cnt_dst (valid after mr->wait(); keep_pos
// must be given and stay alive until then) -- several small results then share one host wait
inline u32 compact_flags(Ctx &c, const u32 *flags, size_t n, Buf<u32> &out, Buf<u32> *keep_pos, MultiRead *mr, u32 *cnt_dst) {
	if (n == 0) { out = Buf<u32>(c, 1); if (keep_pos) { *keep_pos = Buf<u32>(c, 1); keep_pos->zero(); } if (cnt_dst) *cnt_dst = 0; return 0; }
	Buf<u32> pos(c, n + 1);
	exclusive_scan(c, flags, pos.p, n, OpAddU32(), 0u);
	out = Buf<u32>(c, n);     // capacity n: the count is known only after the scatter (one host read instead of two)
	u32 *o = out.p; u32 *p = pos.p;
	parallel_for<K_compact_scatter>(c, n, PV_LAMBDA(size_t i) {
		u32 f = flags[i];
		if (f) o[p[i]] = (u32)i;
		if (i == n - 1) p[n] = p[i] + (f ? 1u : 0u);
	});
	if (mr) {
		if (!keep_pos || !cnt_dst) throw std::runtime_error("compact_flags: a deferred count needs keep_pos and cnt_dst");
		*keep_pos = std::move(pos);
		mr->add(cnt_dst, keep_pos->p + n, 4);
		return 0;
	}
	const u32 cnt = pos.read(n);
	if (keep_pos) *keep_pos = std::move(pos);
	return cnt;
}

// ---------------------------------------------------------------------------------------------
// indices i in [0, n) grouped by key(i) in [0, B), ascending inside a group: off[B + 1] + list[n] (a CSR).
// Almost every group is small in the graphs this serves (steps per vertex, children per PVST vertex), so instead of a
// stable sort of all n pairs: count, scan, drop every index into a free slot of its group (atomic cursor, arbitrary order),
// one thread per group insertion-sorts its short list.  Members of the few large groups (more than `small` entries: hub
// vertices, a PVST root with a million children) take the stable radix sort, restricted to them.
// ---------------------------------------------------------------------------------------------
PV_KERNEL(K_grp_count, 8);
PV_KERNEL(K_grp_slot, 16);
PV_KERNEL(K_grp_sort, 16);
PV_KERNEL(K_grp_big, 12);
// ord (optional): a secondary key -- inside a group the indices come out ascending by (ord(i), i) instead of by i alone.
struct NoOrd { PV_DEV u32 operator()(size_t) const { return 0u; } };
// Two halves, so that a caller can queue several groupings and learn all their "has large groups" flags with ONE host wait:
// group_begin runs count / scan / slot / small-group sort and queues the flag on the caller's MultiRead (or reads it at once);
// group_finish takes the large groups through the radix path.
struct GroupState { Buf<u32> cnt, cur, nBigBuf; size_t n = 0; u32 B = 0; u32 *off = nullptr, *list = nullptr; u32 small = 48; u32 nBig = 0; };
template <class KeyFn, class OrdFn = NoOrd, bool HAS_ORD = false>
inline void group_begin(Ctx &c, KeyFn key, OrdFn ord, size_t n, u32 B, u32 *off, u32 *list, u32 small, GroupState &st, MultiRead *mr = nullptr) {
	if (n >= 0xFFFFFFF0ull) throw std::runtime_error("group_indices: input too large");
	st.n = n; st.B = B; st.off = off; st.list = list; st.small = small; st.nBig = 0;
	st.cnt = Buf<u32>(c, (size_t)B + 1); st.cur = Buf<u32>(c, B ? B : 1); st.nBigBuf = Buf<u32>(c, 1);
	st.cnt.zero(); st.cur.zero(); st.nBigBuf.zero();
	u32 *cnt_ = st.cnt.p, *cur_ = st.cur.p, *nBig_ = st.nBigBuf.p;
	parallel_for<K_grp_count>(c, n, PV_LAMBDA(size_t i) { atomic_add(cnt_ + key(i), 1u); });
	exclusive_scan(c, st.cnt.p, off, (size_t)B + 1, OpAddU32(), 0u);
	parallel_for<K_grp_slot>(c, n, PV_LAMBDA(size_t i) {
		u32 k = key(i);
		if (cnt_[k] > small) return;
		list[off[k] + atomic_add(cur_ + k, 1u)] = (u32)i;
	});
	parallel_for<K_grp_sort>(c, B, PV_LAMBDA(size_t k) {
		u32 m = cnt_[k];
		if (m > small) { nBig_[0] = 1; return; }   // (a flag, not a sum: with many paths every group is large and a counter would serialise)
		u32 *a = list + off[k];
		for (u32 i = 1; i < m; i++) {
			u32 x = a[i]; u32 j = i;
			if (HAS_ORD) { const u64 kx = ((u64)ord(x) << 32) | x; while (j > 0 && (((u64)ord(a[j - 1]) << 32) | a[j - 1]) > kx) { a[j] = a[j - 1]; j--; } }
			else { while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; j--; } }
			a[j] = x;
		}
	});
	if (mr) mr->add(&st.nBig, st.nBigBuf.p, 4);
	else st.nBig = st.nBigBuf.read(0);
}
template <class KeyFn, class OrdFn = NoOrd, bool HAS_ORD = false>
inline void group_finish(Ctx &c, KeyFn key, OrdFn ord, GroupState &st) {
	if (!st.nBig) { st.cnt.release(); st.cur.release(); return; }
	const size_t n = st.n; const u32 B = st.B, small = st.small; u32 *off = st.off, *list = st.list;
	u32 *cnt_ = st.cnt.p, *cur_ = st.cur.p;
	Buf<u32> flag(c, n), idx;
	u32 *flag_ = flag.p;
	parallel_for<K_grp_big>(c, n, PV_LAMBDA(size_t i) { flag_[i] = cnt_[key(i)] > small ? 1u : 0u; });
	const u32 nb = compact_flags(c, flag.p, n, idx);   // ascending index
	flag.release();
	Buf<u32> kk(c, nb), vv(c, nb);
	u32 *kk_ = kk.p, *vv_ = vv.p; const u32 *idx_ = idx.p;
	if (HAS_ORD) {   // stable by the secondary key first (the members are in index order), then by group
		parallel_for<K_grp_big>(c, nb, PV_LAMBDA(size_t j) { u32 i = idx_[j]; kk_[j] = ord(i); vv_[j] = i; });
		radix_sort_pairs(c, kk, vv, nb, 32);
		kk_ = kk.p; vv_ = vv.p;
		parallel_for<K_grp_big>(c, nb, PV_LAMBDA(size_t j) { kk_[j] = key(vv_[j]); });
	} else {
		parallel_for<K_grp_big>(c, nb, PV_LAMBDA(size_t j) { u32 i = idx_[j]; kk_[j] = key(i); vv_[j] = i; });
	}
	radix_sort_pairs(c, kk, vv, nb, bits_for(B));
	kk_ = kk.p; vv_ = vv.p;
	// first sorted position of every large group (cur is free again: large groups never touched it)
	parallel_for<K_grp_big>(c, nb, PV_LAMBDA(size_t j) { if (j == 0 || kk_[j - 1] != kk_[j]) cur_[kk_[j]] = (u32)j; });
	parallel_for<K_grp_big>(c, nb, PV_LAMBDA(size_t j) { u32 k = kk_[j]; list[off[k] + ((u32)j - cur_[k])] = vv_[j]; });
}
template <class KeyFn, class OrdFn = NoOrd, bool HAS_ORD = false>
inline void group_indices_ord(Ctx &c, KeyFn key, OrdFn ord, size_t n, u32 B, u32 *off, u32 *list, u32 small = 48) {
	GroupState st;
	group_begin<KeyFn, OrdFn, HAS_ORD>(c, key, ord, n, B, off, list, small, st);
	group_finish<KeyFn, OrdFn, HAS_ORD>(c, key, ord, st);
}
template <class KeyFn> inline void group_indices(Ctx &c, KeyFn key, size_t n, u32 B, u32 *off, u32 *list, u32 small = 48) {
	group_indices_ord<KeyFn, NoOrd, false>(c, key, NoOrd(), n, B, off, list, small);
}

} // namespace pv
