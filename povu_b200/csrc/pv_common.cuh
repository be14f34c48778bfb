// povu_b200 -- common device/host plumbing for the sm_100a decompose -> call kernels.
//
// Two build modes over the SAME pipeline sources:
//   * default  : nvcc -gencode arch=compute_100a,code=sm_100a  -> the product (libpovu_b200.so)
//   * PV_EMU   : g++ -x c++ -DPV_EMU -> a host-side *debug rehearsal* of the same kernel bodies, one
//                "thread" at a time (optionally in reverse / shuffled order to expose order
//                dependence).  It exists because the development container has no GPU; it is built
//                only by tests/emu/ into a separate libpovu_emu.so, is never loaded by the product
//                path (povu_b200/api.py refuses to run without the CUDA library) and is not a fallback.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <new>
#include <utility>
#include <vector>
#include <mutex>
#include <map>
#include <unordered_map>

#ifndef PV_EMU
#include <cuda_runtime.h>
#define PV_DEV __device__ __forceinline__
#define PV_HD __host__ __device__ __forceinline__
#define PV_LAMBDA [=] __device__
#else
#define PV_DEV inline
#define PV_HD inline
#define PV_LAMBDA [=]
#endif

namespace pv {

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

static const u32 NONE = 0xFFFFFFFFu;

struct u128 { u64 lo, hi; };

#ifndef PV_EMU
#define PV_CUDA_CHECK(x)                                                                         \
	do {                                                                                     \
		cudaError_t e_ = (x);                                                            \
		if (e_ != cudaSuccess)                                                           \
			throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
						 " at " __FILE__ ":" + std::to_string(__LINE__));        \
	} while (0)
#endif

#ifndef PV_EMU
// per-process pool of 1 MiB page-locked staging blocks (GFA text on its way in, VCF text on its way out): page-locking is
// slow (~0.3 ms per MiB), so blocks are allocated once and recycled
struct PinnedPool {
	static const size_t BLK = 1u << 20;
	std::mutex m; std::vector<char *> free_;
	char *get() {
		{ std::lock_guard<std::mutex> l(m); if (!free_.empty()) { char *p = free_.back(); free_.pop_back(); return p; } }
		char *p = nullptr;
		if (cudaHostAlloc((void **)&p, BLK, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		return p;
	}
	void put(char *p) { if (!p) return; std::lock_guard<std::mutex> l(m); free_.push_back(p); }
	static PinnedPool &inst() { static PinnedPool *g = new PinnedPool(); return *g; }   // never destroyed: outlives the CUDA context teardown order
};
#endif

// ---------------------------------------------------------------------------------------------
// execution context: one CUDA stream, stream-ordered allocations, launch counter
// ---------------------------------------------------------------------------------------------
struct Ctx {
#ifndef PV_EMU
	cudaStream_t stream = nullptr;
	bool own_stream = false;
#else
	int emu_order = 0;  // 0 forward, 1 reverse, 2 shuffled
	u64 emu_seed = 1;
#endif
	int device = 0;
	int sm_count = 148;
	u64 launches = 0;       // kernels launched through this context (bench.py: gpu_launches)
	u64 host_syncs = 0;     // times the host waited for the main stream (every one drains the GPU: the launch queue runs dry)
	// optional per-kernel CUDA-event timing (bench.py roofline leg): name -> {launches, items, bytes, ms}
	bool profiling = false;
	struct ProfRec { const char *name; u64 items; u64 bytes;
#ifndef PV_EMU
		cudaEvent_t e0, e1;
#endif
	};
	std::vector<ProfRec> prof;
	struct ProfAgg { std::string name; u64 launches, items, bytes; double ms; };
	std::vector<ProfAgg> prof_agg;
	void prof_begin(const char *name, u64 items, u64 bytes);
	void prof_end();
	void prof_collect();
	u64 bytes_alloc = 0, bytes_peak = 0;
	// Block cache in front of cudaMallocAsync: the pipelines allocate ~200 scratch buffers per step, the big graphs
	// multi-GB ones, and the driver pool's split / merge / remap work showed up as erratic 10-1000 ms stalls.  Freed
	// blocks are kept by size (8 size classes per octave) and handed to the next request of a fitting class; everything
	// runs on the context's stream, so reuse is ordered like the stream.  cache_limit bounds the idle bytes.
#ifndef PV_EMU
	std::multimap<size_t, void *> cache_free; std::unordered_map<void *, size_t> cache_live;
	u64 cache_idle = 0, cache_limit = 0, cache_hits = 0, cache_misses = 0;
	static size_t size_class(size_t bytes);
	void cache_trim(u64 keep);
#endif
	// persistent scratch of the single-pass (decoupled look-back) scan and the onesweep radix sort.  Status words carry
	// an epoch that grows with every use, so the scratch is never cleared between launches; only these two primitives
	// ever write it (memory recycled from the pool could alias a live epoch).
	u32 *lb_flags = nullptr; char *lb_vals = nullptr; size_t lb_cap = 0;     // scan: per tile flag + (aggregate, inclusive) values
	u64 *lb_pack = nullptr;                                                  // flat scan of 4-byte elements: (epoch << 32 | aggregate) per tile, written by that kernel only
	u64 *rs_status = nullptr; size_t rs_cap = 0;                             // sort: per (tile, digit) status
	u32 *lb_ticket = nullptr;                                                // [0] scan tile ticket, [1] sort tile ticket (self-resetting)
	u32 *grid_bar = nullptr;                                                 // grid-wide barrier of the cooperative kernels: arrivals, generation
	u32 lb_epoch = 0;
	u32 next_epoch();
	void lb_reserve(size_t tiles);
	void rs_reserve(size_t words);

	void init(int dev, void *ext_stream);
	void destroy();
	void sync();
	void *raw_alloc(size_t bytes);
	void raw_free(void *p, size_t bytes);
	void memset0(void *p, size_t bytes);
	void memset_ff(void *p, size_t bytes);
	void h2d(void *d, const void *h, size_t bytes);
	void d2h(void *h, const void *d, size_t bytes);   // synchronous on return
	void d2h_async(void *h, const void *d, size_t bytes);   // caller syncs before reading h
	void d2d(void *d, const void *s, size_t bytes);
	// background device -> host copies on a second stream, overlapped with later kernels of the main stream:
	// bg_begin() orders the copy stream after everything issued so far; the SOURCE buffers must stay allocated and
	// the destinations unread until bg_sync().
#ifndef PV_EMU
	cudaStream_t copy_stream = nullptr; cudaEvent_t copy_ev = nullptr;
#endif
	// background host -> device copies on a third stream (inputs that a later phase needs travel while an earlier phase
	// computes): up_begin() orders the upload stream after everything issued so far (the destination buffers were allocated
	// on the main stream), up_join() makes the MAIN STREAM wait for the uploads -- no host synchronisation.
#ifndef PV_EMU
	cudaStream_t up_stream = nullptr; cudaEvent_t up_ev = nullptr;
#endif
	char *h_scratch = nullptr;   // page-locked block for MultiRead
	bool up_pending = false;
	void up_begin();
	void h2d_up(void *d, const void *h, size_t bytes);
	void up_join();
	bool bg_pending = false;
	void bg_begin();
	void d2h_bg(void *h, const void *d, size_t bytes);
	void bg_sync();
};

inline void Ctx::init(int dev, void *ext_stream) {
	device = dev;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaSetDevice(dev));
	cudaDeviceProp p;
	PV_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
	sm_count = p.multiProcessorCount;
	if (ext_stream) { stream = (cudaStream_t)ext_stream; own_stream = false; }
	else { PV_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); own_stream = true; }
	cudaMemPool_t pool;
	PV_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
	u64 thr = ~0ull;   // keep freed blocks cached in the pool between calls
	PV_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
	cache_limit = (u64)(p.totalGlobalMem / 2);   // idle blocks may hold up to half of the device memory
	if (const char *e = getenv("POVU_B200_CACHE_LIMIT_MB")) cache_limit = strtoull(e, nullptr, 10) << 20;
#else
	(void)ext_stream;
	const char *o = getenv("PV_EMU_ORDER");
	emu_order = o ? atoi(o) : 0;
	if (const char *sd = getenv("PV_EMU_SEED")) emu_seed = strtoull(sd, nullptr, 10);
#endif
}
inline u32 Ctx::next_epoch() {
	if (++lb_epoch >= (1u << 30)) {   // wrap: clear the scratch once, start over
		lb_epoch = 1;
		if (lb_flags) { memset0(lb_flags, lb_cap * 4); memset0(lb_pack, lb_cap * 8); }
		if (rs_status) memset0(rs_status, rs_cap * 8);
	}
	return lb_epoch;
}
inline void Ctx::lb_reserve(size_t tiles) {
#ifndef PV_EMU
	if (!lb_ticket) { PV_CUDA_CHECK(cudaMalloc((void **)&lb_ticket, 16)); PV_CUDA_CHECK(cudaMemsetAsync(lb_ticket, 0, 16, stream)); }
	if (tiles <= lb_cap) return;
	sync();
	if (lb_flags) { cudaFree(lb_flags); cudaFree(lb_vals); cudaFree(lb_pack); }
	lb_cap = tiles * 2 > 4096 ? tiles * 2 : 4096;
	PV_CUDA_CHECK(cudaMalloc((void **)&lb_flags, lb_cap * 4));
	PV_CUDA_CHECK(cudaMalloc((void **)&lb_vals, lb_cap * 32));
	PV_CUDA_CHECK(cudaMalloc((void **)&lb_pack, lb_cap * 8));
	PV_CUDA_CHECK(cudaMemsetAsync(lb_flags, 0, lb_cap * 4, stream));
	PV_CUDA_CHECK(cudaMemsetAsync(lb_pack, 0, lb_cap * 8, stream));
#else
	(void)tiles;
#endif
}
inline void Ctx::rs_reserve(size_t words) {
#ifndef PV_EMU
	if (!lb_ticket) { PV_CUDA_CHECK(cudaMalloc((void **)&lb_ticket, 16)); PV_CUDA_CHECK(cudaMemsetAsync(lb_ticket, 0, 16, stream)); }
	if (words <= rs_cap) return;
	sync();
	if (rs_status) cudaFree(rs_status);
	rs_cap = words > (1u << 16) ? words : (1u << 16);
	PV_CUDA_CHECK(cudaMalloc((void **)&rs_status, rs_cap * 8));
	PV_CUDA_CHECK(cudaMemsetAsync(rs_status, 0, rs_cap * 8, stream));
#else
	(void)words;
#endif
}
inline void Ctx::destroy() {
#ifndef PV_EMU
	if (stream) { cache_trim(0); cudaStreamSynchronize(stream); }
	if (lb_flags) { cudaFree(lb_flags); cudaFree(lb_vals); cudaFree(lb_pack); lb_flags = nullptr; lb_vals = nullptr; lb_pack = nullptr; lb_cap = 0; }
	if (rs_status) { cudaFree(rs_status); rs_status = nullptr; rs_cap = 0; }
	if (lb_ticket) { cudaFree(lb_ticket); lb_ticket = nullptr; }
	if (grid_bar) { cudaFree(grid_bar); grid_bar = nullptr; }
	if (copy_stream) { cudaStreamDestroy(copy_stream); cudaEventDestroy(copy_ev); copy_stream = nullptr; copy_ev = nullptr; }
	if (up_stream) { cudaStreamDestroy(up_stream); cudaEventDestroy(up_ev); up_stream = nullptr; up_ev = nullptr; }
	if (h_scratch) { PinnedPool::inst().put(h_scratch); h_scratch = nullptr; }
	if (own_stream && stream) cudaStreamDestroy(stream);
	stream = nullptr;
#endif
}
inline void Ctx::sync() {
	host_syncs++;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaStreamSynchronize(stream));
#endif
}
#ifndef PV_EMU
inline size_t Ctx::size_class(size_t bytes) {
	if (bytes <= 512) return 512;
	int hb = 63 - __builtin_clzll((unsigned long long)bytes);   // highest set bit
	size_t step = (size_t)1 << (hb - 3);                          // 8 classes per octave: at most 12.5 % slack
	return (bytes + step - 1) & ~(step - 1);
}
inline void Ctx::cache_trim(u64 keep) {
	while (cache_idle > keep && !cache_free.empty()) {
		auto it = std::prev(cache_free.end());   // largest first
		cache_idle -= it->first;
		cudaFreeAsync(it->second, stream);
		cache_free.erase(it);
	}
}
#endif
inline void *Ctx::raw_alloc(size_t bytes) {
	if (bytes == 0) bytes = 16;
	bytes_alloc += bytes;
	if (bytes_alloc > bytes_peak) bytes_peak = bytes_alloc;
#ifndef PV_EMU
	const size_t cls = size_class(bytes);
	auto it = cache_free.lower_bound(cls);
	if (it != cache_free.end() && it->first <= cls + cls / 4) {   // a block of this class, or at most two classes up
		void *p = it->second; const size_t sz = it->first;
		cache_free.erase(it); cache_idle -= sz; cache_live[p] = sz; cache_hits++;
		return p;
	}
	void *p = nullptr;
	cudaError_t e = cudaMallocAsync(&p, cls, stream);
	if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); cache_trim(0); cudaStreamSynchronize(stream); e = cudaMallocAsync(&p, cls, stream); }   // give the idle blocks back and retry
	PV_CUDA_CHECK(e);
	cache_live[p] = cls; cache_misses++;
	return p;
#else
	void *p = malloc(bytes);
	if (!p) throw std::runtime_error("emu: out of memory");
	memset(p, 0xCD, bytes);   // poison: uninitialised reads show up
	return p;
#endif
}
inline void Ctx::raw_free(void *p, size_t bytes) {
	if (!p) return;
	bytes_alloc -= bytes ? bytes : 16;
#ifndef PV_EMU
	auto it = cache_live.find(p);
	if (it == cache_live.end()) { cudaFreeAsync(p, stream); return; }
	const size_t sz = it->second;
	cache_live.erase(it);
	cache_free.emplace(sz, p); cache_idle += sz;
	if (cache_limit && cache_idle > cache_limit) cache_trim(cache_limit / 2);
#else
	free(p);
#endif
}
inline void Ctx::memset0(void *p, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream));
#else
	memset(p, 0, bytes);
#endif
}
inline void Ctx::memset_ff(void *p, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemsetAsync(p, 0xFF, bytes, stream));
#else
	memset(p, 0xFF, bytes);
#endif
}
inline void Ctx::h2d(void *d, const void *h, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream));
#else
	memcpy(d, h, bytes);
#endif
}
inline void Ctx::d2h(void *h, const void *d, size_t bytes) {
	if (!bytes) return;
	host_syncs++;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream));
	PV_CUDA_CHECK(cudaStreamSynchronize(stream));
#else
	memcpy(h, d, bytes);
#endif
}
inline void Ctx::d2h_async(void *h, const void *d, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream));
#else
	memcpy(h, d, bytes);
#endif
}
inline void Ctx::up_begin() {
#ifndef PV_EMU
	if (!up_stream) { PV_CUDA_CHECK(cudaStreamCreateWithFlags(&up_stream, cudaStreamNonBlocking)); PV_CUDA_CHECK(cudaEventCreateWithFlags(&up_ev, cudaEventDisableTiming)); }
	PV_CUDA_CHECK(cudaEventRecord(up_ev, stream));
	PV_CUDA_CHECK(cudaStreamWaitEvent(up_stream, up_ev, 0));
#endif
	up_pending = true;
}
inline void Ctx::h2d_up(void *d, const void *h, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, up_stream));
#else
	memcpy(d, h, bytes);
#endif
}
inline void Ctx::up_join() {
	if (!up_pending) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaEventRecord(up_ev, up_stream));
	PV_CUDA_CHECK(cudaStreamWaitEvent(stream, up_ev, 0));
#endif
	up_pending = false;
}
inline void Ctx::bg_begin() {
#ifndef PV_EMU
	if (!copy_stream) { PV_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking)); PV_CUDA_CHECK(cudaEventCreateWithFlags(&copy_ev, cudaEventDisableTiming)); }
	PV_CUDA_CHECK(cudaEventRecord(copy_ev, stream));
	PV_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, copy_ev, 0));
#endif
	bg_pending = true;
}
inline void Ctx::d2h_bg(void *h, const void *d, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, copy_stream));
#else
	memcpy(h, d, bytes);
#endif
}
inline void Ctx::bg_sync() {
	if (!bg_pending) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
#endif
	bg_pending = false;
}
inline void Ctx::d2d(void *d, const void *s, size_t bytes) {
	if (!bytes) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, stream));
#else
	memmove(d, s, bytes);
#endif
}

inline void Ctx::prof_begin(const char *name, u64 items, u64 bytes) {
	if (!profiling) return;
	ProfRec r; r.name = name; r.items = items; r.bytes = bytes;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaEventCreate(&r.e0)); PV_CUDA_CHECK(cudaEventCreate(&r.e1));
	PV_CUDA_CHECK(cudaEventRecord(r.e0, stream));
#endif
	prof.push_back(r);
}
inline void Ctx::prof_end() {
	if (!profiling) return;
#ifndef PV_EMU
	PV_CUDA_CHECK(cudaEventRecord(prof.back().e1, stream));
#endif
}
inline void Ctx::prof_collect() {
	sync();
	for (auto &r : prof) {
		float ms = 0;
#ifndef PV_EMU
		cudaEventElapsedTime(&ms, r.e0, r.e1);
		cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
#endif
		static const bool dump = getenv("PV_PROF_DUMP") != nullptr;
		if (dump) fprintf(stderr, "[prof] %-16s items %10llu bytes %11llu  %8.2f us\n", r.name, (unsigned long long)r.items, (unsigned long long)r.bytes, ms * 1e3);
		bool found = false;
		for (auto &a : prof_agg) if (a.name == r.name) { a.launches++; a.items += r.items; a.bytes += r.bytes; a.ms += ms; found = true; break; }
		if (!found) prof_agg.push_back(ProfAgg{r.name, 1, r.items, r.bytes, ms});
	}
	prof.clear();
}

// kernel tags: a name for launch lists / profiles and the ALGORITHMIC bytes one logical item moves
// (compulsory reads + writes of that kernel, DESIGN.md "per-kernel algorithmic bytes")
#define PV_KERNEL(tag, bytes_per_item)                                   \
	struct tag {                                                     \
		static const char *name() { return #tag; }               \
		static constexpr u64 bytes = (bytes_per_item);           \
	}

// host vectors that receive device results: page-locked storage (D2H at PCIe speed, truly asynchronous) and no
// zero-fill on resize; capacity is kept across calls by the owners, so steady-state steps allocate nothing.
template <class T> struct HostAlloc {
	typedef T value_type;
	HostAlloc() {}
	template <class U> HostAlloc(const HostAlloc<U> &) {}
	T *allocate(size_t n) {
		size_t bytes = n * sizeof(T) + 64;
		char *raw = nullptr; u64 kind = 0;
#ifndef PV_EMU
		if (cudaHostAlloc((void **)&raw, bytes, cudaHostAllocDefault) == cudaSuccess) kind = 1; else { cudaGetLastError(); raw = nullptr; }
#endif
		if (!raw) { raw = (char *)malloc(bytes); kind = 2; }
		if (!raw) throw std::bad_alloc();
		*(u64 *)raw = kind;
		return (T *)(raw + 64);
	}
	void deallocate(T *p, size_t) {
		char *raw = (char *)p - 64;
#ifndef PV_EMU
		if (*(u64 *)raw == 1) { cudaFreeHost(raw); return; }
#endif
		free(raw);
	}
	template <class U> void construct(U *) {}                                   // resize() leaves new elements uninitialised
	template <class U, class A0, class... A> void construct(U *p, A0 &&a0, A &&...a) { ::new ((void *)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
	template <class U> bool operator==(const HostAlloc<U> &) const { return true; }
	template <class U> bool operator!=(const HostAlloc<U> &) const { return false; }
};
template <class T> using hvec = std::vector<T, HostAlloc<T>>;

// coarse phase tracer (PV_TRACE=1): synchronises, so only for finding host-side stalls
struct Trace {
	Ctx &c; const char *scope; bool on; double t0;
	static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
	Trace(Ctx &ctx, const char *s) : c(ctx), scope(s) { static const bool en = getenv("PV_TRACE") != nullptr; on = en; if (on) { c.sync(); t0 = now(); } }
	void mark(const char *what) { if (!on) return; c.sync(); double t = now(); fprintf(stderr, "[trace] %s: %-28s %8.3f ms\n", scope, what, t - t0); t0 = t; }
};

// RAII device buffer (stream-ordered)
template <class T> struct Buf {
	Ctx *c = nullptr;
	T *p = nullptr;
	size_t n = 0;
	Buf() {}
	Buf(Ctx &ctx, size_t n_) : c(&ctx), n(n_) { p = (T *)ctx.raw_alloc(n_ * sizeof(T)); }
	Buf(const Buf &) = delete;
	Buf &operator=(const Buf &) = delete;
	Buf(Buf &&o) noexcept : c(o.c), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
	Buf &operator=(Buf &&o) noexcept {
		if (this != &o) { release(); c = o.c; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
		return *this;
	}
	~Buf() { release(); }
	void release() { if (p && c) c->raw_free(p, n * sizeof(T)); p = nullptr; n = 0; }
	void zero() { c->memset0(p, n * sizeof(T)); }
	void fill_ff() { c->memset_ff(p, n * sizeof(T)); }
	operator T *() const { return p; }
	T *get() const { return p; }
	std::vector<T> to_host() const { std::vector<T> h(n); c->d2h(h.data(), p, n * sizeof(T)); return h; }
	std::vector<T> to_host(size_t m) const { std::vector<T> h(m); c->d2h(h.data(), p, m * sizeof(T)); return h; }   // first m elements
	void to_hvec_bg(hvec<T> &h, size_t m) const { h.resize(m); c->d2h_bg(h.data(), p, m * sizeof(T)); }            // on the copy stream
	// asynchronous copy of the first m elements into a pinned host vector (capacity reused); sync before reading
	void to_hvec(hvec<T> &h, size_t m) const { h.resize(m); c->d2h_async(h.data(), p, m * sizeof(T)); }
	T read(size_t i) const { T v; c->d2h(&v, p + i, sizeof(T)); return v; }
};

// several small device -> host reads with ONE host wait: the values land in a page-locked scratch block of the context
struct MultiRead {
	Ctx &c; size_t used = 0;
	struct Item { void *dst; size_t off, bytes; };
	std::vector<Item> items;
	explicit MultiRead(Ctx &c_) : c(c_) {}
	void add(void *dst, const void *src, size_t bytes) {
#ifndef PV_EMU
		if (!c.h_scratch) { c.h_scratch = PinnedPool::inst().get(); if (!c.h_scratch) throw std::runtime_error("povu_b200: no page-locked scratch"); }
		if (used + bytes > PinnedPool::BLK) throw std::runtime_error("MultiRead: too much data for the scratch block");
		PV_CUDA_CHECK(cudaMemcpyAsync(c.h_scratch + used, src, bytes, cudaMemcpyDeviceToHost, c.stream));
		items.push_back(Item{dst, used, bytes}); used += (bytes + 7) & ~(size_t)7;
#else
		memcpy(dst, src, bytes);
#endif
	}
	void wait() {
		c.sync();
#ifndef PV_EMU
		for (auto &i : items) memcpy(i.dst, c.h_scratch + i.off, i.bytes);
#endif
		items.clear(); used = 0;
	}
};

// ---------------------------------------------------------------------------------------------
// atomics (device: hardware atomics; emu: plain read-modify-write). All return the OLD value.
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
PV_DEV u32 atomic_add(u32 *p, u32 v) { return atomicAdd(p, v); }
PV_DEV u64 atomic_add(u64 *p, u64 v) { return (u64)atomicAdd((unsigned long long *)p, (unsigned long long)v); }
PV_DEV u32 atomic_min(u32 *p, u32 v) { return atomicMin(p, v); }
PV_DEV u32 atomic_max(u32 *p, u32 v) { return atomicMax(p, v); }
PV_DEV u64 atomic_min(u64 *p, u64 v) { return (u64)atomicMin((unsigned long long *)p, (unsigned long long)v); }
PV_DEV u64 atomic_max(u64 *p, u64 v) { return (u64)atomicMax((unsigned long long *)p, (unsigned long long)v); }
PV_DEV u32 atomic_cas(u32 *p, u32 cmp, u32 v) { return atomicCAS(p, cmp, v); }
PV_DEV u32 atomic_or(u32 *p, u32 v) { return atomicOr(p, v); }
PV_DEV u64 atomic_xor(u64 *p, u64 v) { return (u64)atomicXor((unsigned long long *)p, (unsigned long long)v); }
PV_DEV u32 atomic_exch(u32 *p, u32 v) { return atomicExch(p, v); }
PV_DEV u32 ld_volatile(const u32 *p) { return *(const volatile u32 *)p; }
// L2 loads for data that another CTA wrote earlier in the SAME kernel (sweeps of parallel_iter): never served from this SM's L1
PV_DEV u32 ld_cg(const u32 *p) { return __ldcg(p); }
PV_DEV u64 ld_cg(const u64 *p) { return (u64)__ldcg((const unsigned long long *)p); }
#else
template <class T> inline T atomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomic_min(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomic_max(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
inline u32 atomic_cas(u32 *p, u32 cmp, u32 v) { u32 o = *p; if (o == cmp) *p = v; return o; }
inline u32 atomic_or(u32 *p, u32 v) { u32 o = *p; *p = o | v; return o; }
inline u64 atomic_xor(u64 *p, u64 v) { u64 o = *p; *p = o ^ v; return o; }
inline u32 atomic_exch(u32 *p, u32 v) { u32 o = *p; *p = v; return o; }
inline u32 ld_volatile(const u32 *p) { return *p; }
inline u32 ld_cg(const u32 *p) { return *p; }
inline u64 ld_cg(const u64 *p) { return *p; }
#endif

// counter increment where many threads of a warp hit the same counter (e.g. one giant component): lanes with the
// same index elect a leader that adds the group's size once.  Must be called by converged code (no early return
// before it inside the warp's iteration).
#ifndef PV_EMU
PV_DEV void atomic_inc_aggregated(u32 *base, u32 idx) {
	u32 active = __activemask();
	u32 peers = __match_any_sync(active, idx);
	u32 lane = threadIdx.x & 31u;
	if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(base + idx, (u32)__popc(peers));
}
#else
inline void atomic_inc_aggregated(u32 *base, u32 idx) { base[idx]++; }
#endif

// ---------------------------------------------------------------------------------------------
// parallel_for: one logical thread per index, grid-stride, grid sized as a multiple of the SM count.
// The Tag type only names the kernel instantiation (so ncu launch lists read `pf_kernel<K_claim,...>`).
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while its predecessor in the stream is
// still draining (its launch latency hides behind the predecessor's tail); it must not touch memory before pdl_wait(), which
// returns once every earlier grid has completed and flushed.  A step here is a chain of ~140 short dependent kernels, so the
// launch + drain gap between two of them is a visible share of the step.  POVU_B200_NO_PDL=1 launches the plain way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <class... KArgs, class... Args> inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
	static const bool off = getenv("POVU_B200_NO_PDL") != nullptr;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = off ? 0 : 1;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <class Tag, class F> __global__ void __launch_bounds__(256) pf_kernel(size_t n, F f) {
	pdl_wait();
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <class Tag, class F> inline void parallel_for(Ctx &c, size_t n, F f) {
	if (n == 0) return;
	size_t blocks = (n + 255) / 256;
	size_t cap = (size_t)c.sm_count * 8;
	if (blocks > cap) blocks = cap;
	c.prof_begin(Tag::name(), n, n * Tag::bytes);
	cudaError_t lerr = launch_pdl(pf_kernel<Tag, F>, dim3((unsigned)blocks), dim3(256), 0, c.stream, n, f);
	c.prof_end();
	c.launches++;
	// POVU_B200_SYNC_LAUNCHES=1 (debug): synchronise after every launch so that a device fault names its kernel
	static const bool sync_every = getenv("POVU_B200_SYNC_LAUNCHES") != nullptr;
	cudaError_t err = lerr != cudaSuccess ? lerr : (sync_every ? cudaStreamSynchronize(c.stream) : cudaSuccess);
	if (err == cudaSuccess) err = cudaGetLastError();
	if (err != cudaSuccess) throw std::runtime_error(std::string("CUDA error in kernel ") + Tag::name() + " (" + std::to_string(n) + " items): " + cudaGetErrorString(err));
}
#else
template <class Tag, class F> inline void parallel_for(Ctx &c, size_t n, F f) {
	c.launches++;
	static const char *fwd = getenv("PV_EMU_FWD");   // debug: kernels named here run in forward order (bisecting an order dependence)
	if (c.emu_order == 2 && fwd && strstr(fwd, Tag::name())) { c.emu_seed++; for (size_t i = 0; i < n; i++) f(i); }   // (the other kernels keep their permutations)
	else if (c.emu_order == 0) { for (size_t i = 0; i < n; i++) f(i); }
	else if (c.emu_order == 1) { for (size_t i = n; i-- > 0;) f(i); }
	else {
		// deterministic pseudo-random order: i -> (a*i + b) mod n with a coprime to n
		if (n == 0) return;
		u64 a = 0x9E3779B97F4A7C15ull % n; if (a == 0) a = 1;
		auto gcd = [](u64 x, u64 y) { while (y) { u64 t = x % y; x = y; y = t; } return x; };
		while (gcd(a, n) != 1) a++;
		u64 b = (c.emu_seed++ * 0xD1B54A32D192ED03ull) % n;
		for (size_t i = 0; i < n; i++) f((size_t)(((unsigned __int128)a * i + b) % n));
	}
}
#endif

// ---------------------------------------------------------------------------------------------
// parallel_iter: up to `iters` dependent sweeps over [0, n) in ONE cooperative launch, a grid-wide barrier between sweeps
// (pointer doubling: every sweep is one doubling step; as separate launches each step pays launch + drain, ~4 us of a ~13 us
// step at cfg2).  f(i, it) does the work of item i in sweep it; `live` (optional, device, iters words, zeroed by the caller):
// a sweep that did something sets live[it]; the loop ends after the first sweep that leaves its word zero.
// The grid is sized to what is co-resident (occupancy query), as cudaLaunchCooperativeKernel requires.
// ---------------------------------------------------------------------------------------------
#ifndef PV_EMU
__device__ __forceinline__ void grid_barrier(u32 *bar /* [0] arrivals, [1] generation */, u32 nblocks) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const u32 g = *(volatile u32 *)(bar + 1);
		if (atomicAdd(bar, 1u) == nblocks - 1) { atomicExch(bar, 0u); __threadfence(); atomicAdd(bar + 1, 1u); }
		else { while (*(volatile u32 *)(bar + 1) == g) { } }
		__threadfence();
	}
	__syncthreads();
}
template <class Tag, class F> __global__ void __launch_bounds__(256) pi_kernel(size_t n, int iters, F f, const u32 *live, u32 *bar) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (int it = 0; it < iters; it++) {
		for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, it);
		if (it + 1 == iters) break;
		grid_barrier(bar, gridDim.x);
		if (live && *(volatile const u32 *)(live + it) == 0) break;   // (every thread reads the same word after the barrier)
	}
}
template <class Tag, class F> inline void parallel_iter(Ctx &c, size_t n, int iters, F f, const u32 *live = nullptr) {
	if (n == 0 || iters <= 0) return;
	static int per_sm = 0;
	if (!per_sm) { PV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pi_kernel<Tag, F>, 256, 0)); if (per_sm < 1) throw std::runtime_error("parallel_iter: kernel does not fit an SM"); }
	size_t blocks = (n + 255) / 256, cap = (size_t)per_sm * c.sm_count;
	if (blocks > cap) blocks = cap;
	if (!c.grid_bar) { PV_CUDA_CHECK(cudaMalloc((void **)&c.grid_bar, 8)); PV_CUDA_CHECK(cudaMemsetAsync(c.grid_bar, 0, 8, c.stream)); }
	c.prof_begin(Tag::name(), n * (u64)iters, n * (u64)iters * Tag::bytes);
	u32 *bar = c.grid_bar;
	void *args[] = {(void *)&n, (void *)&iters, (void *)&f, (void *)&live, (void *)&bar};
	PV_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)pi_kernel<Tag, F>, dim3((unsigned)blocks), dim3(256), args, 0, c.stream));
	c.prof_end();
	c.launches++;
}
#else
template <class Tag, class F> inline void parallel_iter(Ctx &c, size_t n, int iters, F f, const u32 *live = nullptr) {
	for (int it = 0; it < iters; it++) {
		parallel_for<Tag>(c, n, [=](size_t i) { f(i, it); });
		if (live && it + 1 < iters && live[it] == 0) break;
	}
}
#endif

// splitmix64: label generator for the cycle-equivalence hashing
PV_HD u64 splitmix64(u64 x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

} // namespace pv
