// povu_b200 -- C ABI (include/povu_b200.h) over the device pipelines.
#include "../../include/povu_b200.h"
#include "pv_engine.hpp"
#include "pv_ffi.hpp"
#include "pv_export.hpp"
#include <cerrno>

struct pvgpu_ctx { pv::Ctx c; std::string err; };
struct pvgpu_flubbles { pv::DecompOut o; };
template <class T, class V> static void cp(T *dst, const V &src) { if (dst && !src.empty()) memcpy(dst, src.data(), src.size() * sizeof(T)); }

#ifndef PV_EMU
#define PV_SET_DEVICE(x) PV_CUDA_CHECK(cudaSetDevice((x)->c.device))
#else
#define PV_SET_DEVICE(x) (void)0
#endif

extern "C" {

int pvgpu_build_kind(void) {
#ifndef PV_EMU
	return 1;
#else
	return 2;
#endif
}

int pvgpu_create(int device, void *stream, pvgpu_ctx **out) {
	if (!out) return 1;
	*out = nullptr;
	pvgpu_ctx *x = new pvgpu_ctx();
	try {
#ifndef PV_EMU
		int n = 0;
		if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) throw std::runtime_error("no CUDA device available (povu_b200 has no CPU fallback)");
		if (device < 0 || device >= n) throw std::runtime_error("invalid CUDA device ordinal");
#endif
		x->c.init(device, stream);
	} catch (std::exception &e) {
		fprintf(stderr, "pvgpu_create: %s\n", e.what());
		delete x;
		return 1;
	}
	*out = x;
	return 0;
}
void pvgpu_destroy(pvgpu_ctx *x) { if (x) { x->c.destroy(); delete x; } }
const char *pvgpu_last_error(const pvgpu_ctx *x) { return x ? x->err.c_str() : "null context"; }
uint64_t pvgpu_launch_count(const pvgpu_ctx *x) { return x ? x->c.launches : 0; }
uint64_t pvgpu_host_sync_count(const pvgpu_ctx *x) { return x ? x->c.host_syncs : 0; }
uint64_t pvgpu_peak_device_bytes(const pvgpu_ctx *x) { return x ? x->c.bytes_peak : 0; }
void pvgpu_set_profiling(pvgpu_ctx *x, int on) { if (x) { x->c.profiling = on != 0; if (on) x->c.prof_agg.clear(); } }
size_t pvgpu_profile_json(pvgpu_ctx *x, char *buf, size_t cap) {
	if (!x) return 0;
	try { x->c.prof_collect(); } catch (...) {}
	std::string o = "[";
	for (size_t i = 0; i < x->c.prof_agg.size(); i++) {
		auto &a = x->c.prof_agg[i];
		char tmp[256];
		snprintf(tmp, sizeof tmp, "%s{\"name\":\"%s\",\"launches\":%llu,\"items\":%llu,\"bytes\":%llu,\"ms\":%.6f}", i ? "," : "", a.name.c_str(),
			 (unsigned long long)a.launches, (unsigned long long)a.items, (unsigned long long)a.bytes, a.ms);
		o += tmp;
	}
	o += "]";
	if (buf && cap) { size_t n = o.size() < cap - 1 ? o.size() : cap - 1; memcpy(buf, o.data(), n); buf[n] = 0; }
	return o.size() + 1;
}

static int run_decompose(pvgpu_ctx *x, uint32_t V, uint64_t E, const uint32_t *d1, const uint32_t *d2, uint32_t flags, pvgpu_flubbles **out) {
	pvgpu_flubbles *f = new pvgpu_flubbles();
	pv::DecompOpts opt; opt.want_stack = flags & 1u;
	pv::decompose_device(x->c, V, E, d1, d2, opt, f->o);
	if (f->o.collisions) { delete f; throw std::runtime_error("cycle-equivalence label collision detected by the exact bracket-set check"); }
	*out = f;
	return 0;
}

int pvgpu_decompose_dev(pvgpu_ctx *x, uint32_t V, uint64_t E, const uint32_t *d1, const uint32_t *d2, uint32_t flags, pvgpu_flubbles **out) {
	if (!x || !out) return 1;
	*out = nullptr;
	try {
#ifndef PV_EMU
		PV_CUDA_CHECK(cudaSetDevice(x->c.device));
#endif
		return run_decompose(x, V, E, d1, d2, flags, out);
	} catch (std::exception &e) { x->err = e.what(); return 1; }
}

int pvgpu_decompose(pvgpu_ctx *x, uint32_t V, uint64_t E, const uint32_t *ep1, const uint32_t *ep2, uint32_t flags, pvgpu_flubbles **out) {
	if (!x || !out) return 1;
	*out = nullptr;
	try {
#ifndef PV_EMU
		PV_CUDA_CHECK(cudaSetDevice(x->c.device));
#endif
		pv::Buf<pv::u32> d1(x->c, E ? E : 1), d2(x->c, E ? E : 1);
		x->c.h2d(d1.p, ep1, E * 4); x->c.h2d(d2.p, ep2, E * 4);
		return run_decompose(x, V, E, d1.p, d2.p, flags, out);
	} catch (std::exception &e) { x->err = e.what(); return 1; }
}

uint64_t pvgpu_flubbles_count(const pvgpu_flubbles *f) { return f->o.a_v.size(); }
uint32_t pvgpu_flubbles_component_count(const pvgpu_flubbles *f) { return (uint32_t)f->o.comp_root.size(); }
uint32_t pvgpu_flubbles_dfs_rounds(const pvgpu_flubbles *f) { return f->o.dfs_rounds; }
uint32_t pvgpu_flubbles_hash_collisions(const pvgpu_flubbles *f) { return f->o.collisions; }
uint32_t pvgpu_flubbles_arc_rounds(const pvgpu_flubbles *f) { return f->o.arc_rounds; }
uint32_t pvgpu_flubbles_faithful_classes(const pvgpu_flubbles *f) { return f->o.faithful_classes; }
void pvgpu_flubbles_copy(const pvgpu_flubbles *f, uint32_t *comp, uint32_t *a_v, uint8_t *a_fwd, uint32_t *z_v, uint8_t *z_fwd, uint32_t *parent, uint32_t *depth) {
	cp(comp, f->o.comp); cp(a_v, f->o.a_v); cp(a_fwd, f->o.a_fwd); cp(z_v, f->o.z_v); cp(z_fwd, f->o.z_fwd); cp(parent, f->o.parent); cp(depth, f->o.depth);
}
void pvgpu_flubbles_copy_components(const pvgpu_flubbles *f, uint32_t *root_v, uint32_t *nv, uint32_t *nfl, uint32_t *tips, uint32_t *nonlam) {
	cp(root_v, f->o.comp_root); cp(nv, f->o.comp_nv); cp(nfl, f->o.comp_nfl); cp(tips, f->o.comp_has_tips); cp(nonlam, f->o.comp_nonlaminar);
}
uint64_t pvgpu_flubbles_stack_size(const pvgpu_flubbles *f) { return f->o.stack_v.size(); }
void pvgpu_flubbles_copy_stack(const pvgpu_flubbles *f, uint32_t *v, uint32_t *cls, uint32_t *fwd) { cp(v, f->o.stack_v); cp(cls, f->o.stack_class); cp(fwd, f->o.stack_fwd); }
void pvgpu_flubbles_free(pvgpu_flubbles *f) { delete f; }

/* ---------------------------------------------------------------- primitive self-test (tests/test_prims_gpu.py) */
int pvgpu_selftest_prims(pvgpu_ctx *x, uint64_t n, uint64_t seed, int bits, uint64_t *mismatches) {
	if (!x || !mismatches) return 1;
	*mismatches = 0;
	try {
		PV_SET_DEVICE(x);
		using namespace pv;
		Ctx &c = x->c;
		u64 bad = 0, seen = 0;
		auto section = [&](const char *what) { if (bad != seen) fprintf(stderr, "[selftest n=%llu seed=%llu bits=%d] %s: %llu mismatches\n", (unsigned long long)n, (unsigned long long)seed, bits, what, (unsigned long long)(bad - seen)); seen = bad; };
		std::vector<u32> hk(n), hv(n);
		u32 mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
		for (u64 i = 0; i < n; i++) { hk[i] = (u32)splitmix64(seed + i) & mask; hv[i] = (u32)i; }
		if (seed & 1) for (u64 i = 0; i < n; i++) hk[i] = (u32)((i * 3) / 2) & mask;   // sorted-ish keys (path steps by vertex)
		{	// stable sort of pairs
			Buf<u32> k(c, n ? n : 1), v(c, n ? n : 1);
			c.h2d(k.p, hk.data(), n * 4); c.h2d(v.p, hv.data(), n * 4);
			radix_sort_pairs(c, k, v, n, bits);
			std::vector<u32> gk = k.to_host(), gv = v.to_host();
			std::vector<u32> idx(n); for (u64 i = 0; i < n; i++) idx[i] = (u32)i;
			std::stable_sort(idx.begin(), idx.end(), [&](u32 a, u32 b) { return hk[a] < hk[b]; });
			for (u64 i = 0; i < n; i++) if (gk[i] != hk[idx[i]] || gv[i] != idx[i]) bad++;
			section("sort");
			// pointer flavour
			Buf<u32> k2(c, n ? n : 1), v2(c, n ? n : 1);
			c.h2d(k2.p, hk.data(), n * 4); c.h2d(v2.p, hv.data(), n * 4);
			radix_sort_pairs_inplace(c, k2.p, v2.p, n, bits);
			gk = k2.to_host(); gv = v2.to_host();
			for (u64 i = 0; i < n; i++) if (gk[i] != hk[idx[i]] || gv[i] != idx[i]) bad++;
			section("sort (pointer flavour)");
		}
		{	// scans: u32 add (exclusive, in place), u64 add, i32 inclusive, u128 xor, max-plus composition (non-commutative)
			Buf<u32> a(c, n + 1);
			std::vector<u32> h(n + 1); for (u64 i = 0; i <= n; i++) h[i] = (u32)(splitmix64(seed ^ i) & 1023u);
			c.h2d(a.p, h.data(), (n + 1) * 4);
			exclusive_scan(c, a.p, a.p, n + 1, OpAddU32(), 0u);
			std::vector<u32> g = a.to_host(); u32 run = 0;
			for (u64 i = 0; i <= n; i++) { if (g[i] != run) bad++; run += h[i]; }
			section("scan u32 exclusive in place");
			// misaligned views (scalar path)
			if (n > 8) {
				c.h2d(a.p, h.data(), (n + 1) * 4);
				Buf<u32> o(c, n + 1);
				inclusive_scan(c, a.p + 1, o.p + 3, n - 4, OpAddU32(), 0u);
				g = o.to_host(); run = 0;
				for (u64 i = 0; i + 4 < n; i++) { run += h[i + 1]; if (g[i + 3] != run) bad++; }
			}
			section("scan u32 inclusive misaligned");
			Buf<u64> b(c, n + 1), bo(c, n + 1);
			std::vector<u64> hb(n + 1); for (u64 i = 0; i <= n; i++) hb[i] = splitmix64(seed + 7 * i) >> 20;
			c.h2d(b.p, hb.data(), (n + 1) * 8);
			exclusive_scan(c, b.p, bo.p, n + 1, OpAddU64(), (u64)0);
			std::vector<u64> gb = bo.to_host(); u64 r64 = 0;
			for (u64 i = 0; i <= n; i++) { if (gb[i] != r64) bad++; r64 += hb[i]; }
			section("scan u64");
			Buf<u128> xk(c, n + 1), xo(c, n + 1);
			std::vector<u128> hx(n + 1); for (u64 i = 0; i <= n; i++) hx[i] = u128{splitmix64(seed + 2 * i), splitmix64(seed + 2 * i + 1)};
			c.h2d(xk.p, hx.data(), (n + 1) * 16);
			exclusive_scan(c, (const u128 *)xk.p, xo.p, n + 1, OpXor128(), u128{0, 0});
			std::vector<u128> gx = xo.to_host(); u128 rx{0, 0};
			for (u64 i = 0; i <= n; i++) { if (gx[i].lo != rx.lo || gx[i].hi != rx.hi) bad++; rx.lo ^= hx[i].lo; rx.hi ^= hx[i].hi; }
			section("scan u128 xor");
			Buf<MP> m(c, n ? n : 1), mo(c, n ? n : 1);
			std::vector<MP> hm(n);
			for (u64 i = 0; i < n; i++) { u64 z = splitmix64(seed + 11 * i); i32 o = (i32)(z & 1), u = (i32)((z >> 1) & 1); hm[i] = MP{(z >> 2) % 97 == 0 ? -(1 << 29) : o - u, o}; }
			c.h2d(m.p, hm.data(), n * sizeof(MP));
			inclusive_scan(c, m.p, mo.p, n, OpMP(), MP{0, -(1 << 29)});
			std::vector<MP> gm = mo.to_host(); MP rm{0, -(1 << 29)}; OpMP op;
			for (u64 i = 0; i < n; i++) { rm = op(rm, hm[i]); if ((gm[i].a > gm[i].b ? gm[i].a : gm[i].b) != (rm.a > rm.b ? rm.a : rm.b)) bad++; }   // the function applied to 0 (the clamp at -2^29 makes `a` grouping-dependent far below any b)
			section("scan max-plus");
			// segmented running maximum (an operator whose `identity` only works on the left)
			Buf<u64> sg(c, n ? n : 1), sgo(c, n ? n : 1);
			std::vector<u64> hs(n); u64 key = 0;
			for (u64 i = 0; i < n; i++) { u64 z = splitmix64(seed + 13 * i); if (z % 37 == 0) key += 1 + (z >> 8) % 3; hs[i] = (key << 32) | (u32)(z >> 40); }
			c.h2d(sg.p, hs.data(), n * 8);
			inclusive_scan(c, sg.p, sgo.p, n, SegMax(), (u64)0);
			std::vector<u64> gs = sgo.to_host(); u64 rs = 0; SegMax sop;
			for (u64 i = 0; i < n; i++) { rs = i ? sop(rs, hs[i]) : hs[i]; if (gs[i] != rs) bad++; }
			section("scan segmented max");
		}
		if (n >= 2) {	// range minimum + previous smaller
			Buf<u32> a(c, n);
			std::vector<u32> h(n); for (u64 i = 0; i < n; i++) h[i] = (u32)(splitmix64(seed * 3 + i) % 1000003u);
			c.h2d(a.p, h.data(), n * 4);
			Rmq r; rmq_build(c, r, a.p, (u32)n);
			const u32 Q = 4096;
			Buf<u32> ql(c, Q), qr(c, Q), out(c, Q), out2(c, Q);
			std::vector<u32> hl(Q), hr(Q);
			for (u32 q = 0; q < Q; q++) { u64 z = splitmix64(seed + 99 + q); u32 l = (u32)(z % n), len = (u32)((z >> 32) % ((q & 1) ? 64 : n)); u32 rr = l + len < n ? l + len : (u32)n - 1; hl[q] = l; hr[q] = rr; }
			c.h2d(ql.p, hl.data(), Q * 4); c.h2d(qr.p, hr.data(), Q * 4);
			RmqView rv = r.view(); const u32 *ql_ = ql.p, *qr_ = qr.p; u32 *o_ = out.p, *o2_ = out2.p; const u32 *a_ = a.p;
			parallel_for<K_rmq_blocks>(c, Q, PV_LAMBDA(size_t q) { o_[q] = rv.query(ql_[q], qr_[q]); o2_[q] = rv.prev_smaller(qr_[q], a_[ql_[q]]); });
			std::vector<u32> g = out.to_host(), g2 = out2.to_host();
			for (u32 q = 0; q < Q; q++) {
				u32 mn = 0xFFFFFFFFu; for (u32 i = hl[q]; i <= hr[q]; i++) mn = h[i] < mn ? h[i] : mn;
				if (g[q] != mn) bad++;
				section("rmq query");
				u32 ps = NONE; for (u32 i = hr[q] + 1; i-- > 0;) if (h[i] < h[hl[q]]) { ps = i; break; }
				if (g2[q] != ps) { bad++; if (bad - seen == 1) fprintf(stderr, "  prev_smaller(r=%u, bound=a[%u]=%u): got %u want %u\n", hr[q], hl[q], h[hl[q]], g2[q], ps); }
				section("rmq prev_smaller");
			}
		}
		c.sync();
		*mismatches = bad;
		return 0;
	} catch (std::exception &e) { x->err = e.what(); return 1; }
}

/* ---------------------------------------------------------------- fired-arc fixpoint self-test (tests/test_prims_*.py)
 * a random class sequence with crossing arcs; device fixpoint vs the reference's stack machine (flubbles.cpp:295-367) on the host */
int pvgpu_selftest_arcs(pvgpu_ctx *x, uint32_t n, uint32_t n_classes, uint64_t seed, uint64_t *mismatches, int *rounds) {
	if (!x || !mismatches || n == 0 || n_classes == 0) return 1;
	*mismatches = 0;
	try {
		PV_SET_DEVICE(x);
		using namespace pv;
		Ctx &c = x->c;
		std::vector<u32> cls(n), prv(n, NONE), last(n_classes, NONE);
		for (u32 i = 0; i < n; i++) {
			u64 z = splitmix64(seed + i);
			// mostly fresh classes with bursts of re-encounters of recent ones: nested AND crossing arcs
			cls[i] = (z % 3 == 0) ? (u32)((z >> 8) % n_classes) : (u32)(((z >> 8) % 7 + i / 5) % n_classes);
			prv[i] = last[cls[i]]; last[cls[i]] = i;
		}
		// the reference's stack machine
		std::vector<u32> want(n, 2), stk; std::vector<u8> on(n, 0);
		for (u32 i = 0; i < n; i++) {
			u32 p = prv[i];
			if (p != NONE && on[p]) { want[i] = 1; while (!stk.empty()) { u32 q = stk.back(); stk.pop_back(); on[q] = 0; if (q == p) break; } }
			stk.push_back(i); on[i] = 1;
		}
		Buf<u32> dprv(c, n), fire(c, n);
		c.h2d(dprv.p, prv.data(), (size_t)n * 4);
		int r = fired_arcs(c, dprv.p, n, fire.p, nullptr, nullptr, nullptr, nullptr);
		if (rounds) *rounds = r;
		std::vector<u32> got = fire.to_host();
		u64 bad = 0;
		if (r >= 0) for (u32 i = 0; i < n; i++) if (got[i] != want[i]) bad++;
		*mismatches = bad;
		return 0;
	} catch (std::exception &e) { x->err = e.what(); return 1; }
}

/* ---------------------------------------------------------------- text formatting self-test (tests/test_text_*.py) */
namespace pv { PV_KERNEL(X_selftest, 8); }
int pvgpu_selftest_text(pvgpu_ctx *x, uint32_t max_an, uint64_t *mismatches) {
	if (!x || !mismatches || max_an == 0) return 1;
	*mismatches = 0;
	try {
		PV_SET_DEVICE(x);
		using namespace pv;
		Ctx &c = x->c;
		// AF of every (ac, an) with 1 <= an <= max_an, 0 <= ac <= an, and plain integers, against the C library
		const size_t n = (size_t)(max_an + 1) * (max_an + 1);
		Buf<u8> out(c, n * 16);
		u8 *out_ = out.p;
		parallel_for<X_selftest>(c, n, PV_LAMBDA(size_t i) {
			u32 an = (u32)(i / (max_an + 1)), ac = (u32)(i % (max_an + 1));
			u8 *o = out_ + i * 16;
			for (int k = 0; k < 16; k++) o[k] = 0;
			if (an == 0 || ac > an) return;
			WriteSink s(o); put_af(s, ac, an); s.ch(' '); put_u32(s, ac * 7919u + an); s.finish();
		});
		std::vector<u8> h = out.to_host();
		u64 bad = 0;
		for (size_t i = 0; i < n; i++) {
			u32 an = (u32)(i / (max_an + 1)), ac = (u32)(i % (max_an + 1));
			if (an == 0 || ac > an) continue;
			char want[32]; snprintf(want, sizeof want, "%.1f %u", (double)ac / (double)an, ac * 7919u + an);
			if (strncmp(want, (const char *)&h[i * 16], 16) != 0) { if (bad < 5) fprintf(stderr, "[selftest text] %u/%u: got '%.16s' want '%s'\n", ac, an, (const char *)&h[i * 16], want); bad++; }
		}
		*mismatches = bad;
		return 0;
	} catch (std::exception &e) { x->err = e.what(); return 1; }
}

/* ---------------------------------------------------------------- sessions */
struct pvgpu_session { pvgpu_ctx *x; pve::Session s; explicit pvgpu_session(pvgpu_ctx *x_) : x(x_), s(x_->c) {} };
static char *dup_text(const std::string &t, size_t *len) {
	char *p = (char *)malloc(t.size() + 1);
	if (!p) return nullptr;
	memcpy(p, t.data(), t.size()); p[t.size()] = 0;
	if (len) *len = t.size();
	return p;
}
#define PV_TRY(x, body) try { PV_SET_DEVICE(x); body; return 0; } catch (std::exception &e) { (x)->err = e.what(); return 1; }

int pvgpu_session_open(pvgpu_ctx *x, const char *gfa, int labels, int refs, pvgpu_session **out) {
	if (!x || !gfa || !out) return 1;
	*out = nullptr;
	pvgpu_session *s = nullptr;
	try { PV_SET_DEVICE(x); s = new pvgpu_session(x); s->s.load_gfa(gfa, labels != 0, refs != 0); *out = s; return 0; }
	catch (std::exception &e) { x->err = e.what(); delete s; return 1; }
}
int pvgpu_session_open_hostparse(pvgpu_ctx *x, const char *gfa, int labels, int refs, pvgpu_session **out) {
	if (!x || !gfa || !out) return 1;
	*out = nullptr;
	pvgpu_session *s = nullptr;
	try { PV_SET_DEVICE(x); s = new pvgpu_session(x); s->s.load_gfa(gfa, labels != 0, refs != 0, true); *out = s; return 0; }
	catch (std::exception &e) { x->err = e.what(); delete s; return 1; }
}
void pvgpu_session_gfa_stats(const pvgpu_session *s, uint64_t *text_bytes, double ms[2]) {
	if (text_bytes) *text_bytes = s->s.gfa_text_bytes;
	if (ms) { ms[0] = s->s.t_gfa_read_ms; ms[1] = s->s.t_gfa_parse_ms; }
}
uint64_t pvgpu_session_label_bytes(pvgpu_session *s) {
	try { PV_SET_DEVICE(s->x); s->s.ensure_host(false, false, true); return s->s.g.lblOff.empty() ? 0 : s->s.g.lblOff.back(); } catch (std::exception &e) { s->x->err = e.what(); return 0; }
}
const char *pvgpu_session_path_name(const pvgpu_session *s, uint32_t k) { return k < s->s.g.H() ? s->s.g.paths[k].tag.c_str() : nullptr; }
int pvgpu_session_copy_graph(pvgpu_session *s, uint32_t *ids, uint32_t *seg_len, uint32_t *ep1, uint32_t *ep2, uint32_t *step_off, uint32_t *steps,
			     uint32_t *hap_len, uint64_t *label_off, char *labels) {
	PV_TRY(s->x, {
		pvh::HostGraph &g = s->s.g;
		s->s.ensure_host(ep1 || ep2, steps != nullptr, label_off || labels);
		if (ids) memcpy(ids, g.ids.data(), g.ids.size() * 4);
		if (seg_len) memcpy(seg_len, g.segLen.data(), g.segLen.size() * 4);
		if (ep1) memcpy(ep1, g.ep1.data(), g.ep1.size() * 4);
		if (ep2) memcpy(ep2, g.ep2.data(), g.ep2.size() * 4);
		if (step_off) memcpy(step_off, g.stepOff.data(), g.stepOff.size() * 4);
		if (steps) memcpy(steps, g.steps.data(), g.steps.size() * 4);
		if (hap_len) for (uint32_t k = 0; k < g.H(); k++) hap_len[k] = g.paths[k].hap_len;
		if (label_off) memcpy(label_off, g.lblOff.data(), g.lblOff.size() * 8);
		if (labels && g.has_labels) memcpy(labels, g.lbl.data(), g.lbl.size());
	})
}
int pvgpu_session_from_arrays(pvgpu_ctx *x, uint32_t V, const uint32_t *ids, const uint32_t *seg_len, const uint64_t *label_off, const char *labels,
			      uint64_t E, const uint32_t *ep1, const uint32_t *ep2, uint32_t H, const char *const *names, const uint32_t *step_off,
			      const uint32_t *steps, pvgpu_session **out) {
	if (!x || !out) return 1;
	*out = nullptr;
	pvgpu_session *s = nullptr;
	try {
		PV_SET_DEVICE(x);
		s = new pvgpu_session(x);
		pvh::HostGraph &g = s->s.g;
		g.ids.assign(ids, ids + V);
		for (uint32_t i = 1; i < V; i++) if (ids[i - 1] >= ids[i]) throw std::runtime_error("segment ids must be strictly ascending");
		g.lblOff.assign((size_t)V + 1, 0);
		if (labels && label_off) {
			g.has_labels = true; g.lblOff.assign(label_off, label_off + V + 1); g.lbl.assign(labels, labels + label_off[V]);
			g.segLen.resize(V); for (uint32_t i = 0; i < V; i++) g.segLen[i] = (uint32_t)(label_off[i + 1] - label_off[i]);
		} else if (seg_len) g.segLen.assign(seg_len, seg_len + V);
		else g.segLen.assign(V, 1);
		g.build_id_table();
		g.ep1.assign(ep1, ep1 + E); g.ep2.assign(ep2, ep2 + E);
		for (uint64_t e = 0; e < E; e++) if (ep1[e] >= 2 * V || ep2[e] >= 2 * V) throw std::runtime_error("edge endpoint out of range");
		g.stepOff.assign(1, 0);
		for (uint32_t h = 0; h < H; h++) {
			pvh::HostPath p = pvh::parse_path_name(names[h]);
			uint64_t len = 0;
			for (uint32_t k = step_off[h]; k < step_off[h + 1]; k++) { if ((steps[k] >> 1) >= V) throw std::runtime_error("path step out of range"); len += g.segLen[steps[k] >> 1]; }
			p.hap_len = (uint32_t)len;
			g.paths.push_back(std::move(p)); g.stepOff.push_back(step_off[h + 1] - step_off[0]);
		}
		if (H) g.steps.assign(steps + step_off[0], steps + step_off[H]);
		s->s.pin();
		*out = s; return 0;
	} catch (std::exception &e) { x->err = e.what(); delete s; return 1; }
}
void pvgpu_session_close(pvgpu_session *s) { delete s; }
int pvgpu_session_hot_path(pvgpu_session *s, const char *const *prefixes, size_t n, int keep, double ms[4]) {
	PV_TRY(s->x, {
		std::vector<std::string> p; for (size_t i = 0; i < n; i++) p.push_back(prefixes[i]);
		if (p != s->s.prefixes || !s->s.gm_ready) s->s.set_prefixes(p);
		s->s.hot_path(keep != 0);
		if (ms) { ms[0] = s->s.t_upload_ms; ms[1] = s->s.t_dec_ms; ms[2] = s->s.t_build_ms; ms[3] = s->s.t_call_ms; }
	})
}
void pvgpu_session_sizes(const pvgpu_session *s, uint64_t *v, uint64_t *e, uint64_t *h, uint64_t *st) {
	if (v) *v = s->s.g.V(); if (e) *e = s->s.g.E(); if (h) *h = s->s.g.H(); if (st) *st = s->s.g.stepOff.empty() ? 0 : s->s.g.stepOff.back();
}
int pvgpu_session_decompose(pvgpu_session *s) { PV_TRY(s->x, s->s.decompose()) }
uint64_t pvgpu_session_flubble_count(const pvgpu_session *s) { return s->s.forest.n(); }
uint32_t pvgpu_session_tree_count(const pvgpu_session *s) { return (uint32_t)s->s.forest.comp_id.size(); }
uint32_t pvgpu_session_tree_component(const pvgpu_session *s, uint32_t k) { return s->s.forest.comp_id[k]; }
char *pvgpu_session_pvst_text(const pvgpu_session *s, uint32_t k, size_t *len) {
	try { return dup_text(s->s.pvst_text(k), len); } catch (std::exception &e) { s->x->err = e.what(); return nullptr; }
}
int pvgpu_session_set_debug_numbering(pvgpu_session *s, int on) { s->s.debug_numbering = on != 0; return 0; }
char *pvgpu_session_hairpin_text(const pvgpu_session *s, size_t *len) {
	try { return dup_text(pve::hairpin_text(s->s), len); } catch (std::exception &e) { s->x->err = e.what(); return nullptr; }
}
char *pvgpu_session_flubble_debug_frames(const pvgpu_session *s, size_t *len) {
	try { return dup_text(pve::flubble_debug_frames(s->s), len); } catch (std::exception &e) { s->x->err = e.what(); return nullptr; }
}
int pvgpu_session_structure_export(pvgpu_session *s, const char *json_path, const char *input_name, size_t chunk, const char *frames_jsonl) {
	PV_TRY(s->x, pve::structure_export(s->s, json_path, input_name ? input_name : "", chunk, frames_jsonl ? frames_jsonl : ""))
}
int pvgpu_session_write_pvst(const pvgpu_session *s, const char *dir) { PV_TRY(s->x, s->s.write_pvsts(dir)) }
int pvgpu_session_read_pvst(pvgpu_session *s, const char *dir) { PV_TRY(s->x, s->s.read_pvsts(dir)) }
int pvgpu_session_call(pvgpu_session *s, const char *const *prefixes, size_t n) {
	PV_TRY(s->x, { std::vector<std::string> p; for (size_t i = 0; i < n; i++) p.push_back(prefixes[i]); s->s.set_prefixes(p); s->s.call(); })
}
int pvgpu_session_set_prefixes(pvgpu_session *s, const char *const *prefixes, size_t n) {
	PV_TRY(s->x, { std::vector<std::string> p; for (size_t i = 0; i < n; i++) p.push_back(prefixes[i]); s->s.set_prefixes(p); s->s.rec.reset(); s->s.rec.rec_grpOff.push_back(0); s->s.rec.rec_gtOff.push_back(0); })
}
int pvgpu_session_set_restrict(pvgpu_session *s, const char *region) { PV_TRY(s->x, s->s.set_restrict(region)) }
uint64_t pvgpu_session_record_count(const pvgpu_session *s) { return s->s.rec.n(); }
void pvgpu_session_call_stats(const pvgpu_session *s, uint64_t o[12]) {
	const pv::Records &c = s->s.rec;
	o[0] = c.n_called; o[1] = c.n_regions; o[2] = c.n_tangled; o[3] = c.n_pin_pairs; o[4] = c.n_sort_failed; o[5] = c.n_hard_errors; o[6] = c.n_chunks; o[7] = c.n_too_small; o[8] = c.max_J; o[9] = c.n_big_bfs; o[10] = c.n_subr_clamped; o[11] = c.n_slice_tree_replays;
}
// header (host strings) + the device-written record text (pv_text.cuh)
static void vcf_parts(const pve::Session &S, size_t chunk, int with_date, const char *prefix, std::string &header, pv::TextOut &body, int fd = -1, pv::u64 *streamed = nullptr) {
	auto put_header = [&]() {   // streaming: the header goes out before the record text
		if (fd < 0) return;
		const char *p = header.data(); size_t n = header.size();
		while (n) { ssize_t w = write(fd, p, n); if (w < 0) { if (errno == EINTR) continue; throw std::runtime_error(std::string("write failed: ") + strerror(errno)); } p += w; n -= (size_t)w; }
	};
	if (!prefix) { header = S.vcf_header(S.prefixes, with_date != 0); put_header(); S.vcf_records_device(chunk, nullptr, body, fd, streamed); return; }
	std::set<pv::u32> only = pvh::refs_in_sample(S.g, prefix);
	header = S.vcf_header({std::string(prefix)}, with_date != 0);
	put_header();
	S.vcf_records_device(chunk, &only, body, fd, streamed);
}
char *pvgpu_session_vcf_text(const pvgpu_session *s, size_t chunk, int with_date, const char *prefix, size_t *len) {
	try {
		PV_SET_DEVICE(s->x);
		std::string header; pv::TextOut &body = s->s.text_buf;   // page-locked, capacity reused across calls
		vcf_parts(s->s, chunk, with_date, prefix, header, body);
		char *p = (char *)malloc(header.size() + body.text.size() + 1);
		if (!p) return nullptr;
		memcpy(p, header.data(), header.size());
		if (!body.text.empty()) memcpy(p + header.size(), body.text.data(), body.text.size());   // (memcpy from a null pointer, even of 0 bytes, lets the compiler assume non-null)
		p[header.size() + body.text.size()] = 0;
		if (len) *len = header.size() + body.text.size();
		return p;
	} catch (std::exception &e) { s->x->err = e.what(); return nullptr; }
}
int pvgpu_session_vcf_write(const pvgpu_session *s, size_t chunk, int with_date, const char *prefix, int fd, uint64_t *bytes) {
	PV_TRY(s->x, {
		std::string header; pv::TextOut body; pv::u64 streamed = 0;
		vcf_parts(s->s, chunk, with_date, prefix, header, body, fd, &streamed);   // device text -> ring of page-locked blocks -> write(2)
		if (bytes) *bytes = header.size() + streamed;
	})
}
int pvgpu_session_shard(pvgpu_session *s, uint32_t rank, uint32_t world, uint32_t *n_total) {
	PV_TRY(s->x, { s->s.shard(rank, world); if (n_total) *n_total = s->s.n_comp_total; })
}
int pvgpu_session_reshard(pvgpu_session *s, uint32_t rank, uint32_t world, int from_host, uint32_t *n_total, double *ms) {
	PV_TRY(s->x, { s->s.shard(rank, world, from_host != 0); if (n_total) *n_total = s->s.n_comp_total; if (ms) *ms = s->s.t_shard_ms; })
}
uint64_t pvgpu_session_records_packed_bytes(const pvgpu_session *s) { return s->s.records_packed_bytes(); }
int pvgpu_session_records_pack(pvgpu_session *s, void *dst_device, uint64_t capacity, uint64_t *bytes) {
	PV_TRY(s->x, { pv::u64 b = 0; s->s.records_pack(dst_device, capacity, &b); if (bytes) *bytes = b; })
}
void pvgpu_session_tree_regions(const pvgpu_session *s, uint32_t *comp_id, uint64_t *n_regions) {
	std::vector<uint64_t> n = s->s.regions_per_tree();
	for (size_t t = 0; t < n.size(); t++) { if (comp_id) comp_id[t] = s->s.forest.comp_id[t]; if (n_regions) n_regions[t] = n[t]; }
}
char *pvgpu_session_record_lines(const pvgpu_session *s, size_t *len, uint64_t **line_off, uint32_t **keys, size_t *n) {
	try {
		PV_SET_DEVICE(s->x);
		pv::TextOut &body = s->s.text_buf; std::vector<uint32_t> k;
		s->s.record_lines(body, k);
		*n = body.n;
		*line_off = (uint64_t *)malloc(body.lineOff.size() * 8); memcpy(*line_off, body.lineOff.data(), body.lineOff.size() * 8);
		*keys = (uint32_t *)malloc(k.size() * 4 + 4); if (!k.empty()) memcpy(*keys, k.data(), k.size() * 4);
		char *p = (char *)malloc(body.text.size() + 1);
		if (!p) return nullptr;
		if (!body.text.empty()) memcpy(p, body.text.data(), body.text.size());
		p[body.text.size()] = 0;
		if (len) *len = body.text.size();
		return p;
	} catch (std::exception &e) { s->x->err = e.what(); return nullptr; }
}
void pvgpu_free(void *p) { free(p); }

}
