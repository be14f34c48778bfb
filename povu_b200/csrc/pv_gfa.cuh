// povu_b200 -- GFA text -> flat graph arrays, parsed on the device.
//
// The step immediately before the hot path.  The reference parses the text twice per gfa2vcf on one host thread
// (mto::from_gfa::to_bd, /root/reference/src/mto/from_gfa.cpp:28-113 pre-scan, :141-280 graph build, through the
// liteseq parser whose call-site contract is restated in SURVEY.md Appendix A); here the file is streamed through two
// page-locked staging buffers into HBM and split / validated / converted by byte-parallel kernels:
//   G_nl_count / G_nl_write   newline positions                      (16 B of text per thread, SWAR byte compare)
//   G_classify / G_split      record type per line, the pre-scan's validation of S lines, S / L / P line lists
//   G_seg (+ sort, G_lbl)     segment ids, sequence spans, ascending-id vertex order (first S line of an id wins)
//   G_link                    L lines -> endpoint ports in L-line order
//   G_phead / G_ptab / G_pcount / G_pstep    P lines: a walk is split at its commas by all threads, not one per line
//   G_haplen                  contig lengths
// Same results and the same error text as the host loader pvh::load_gfa (which stays as the checker of this file's
// tests): errors are reduced to (first offending line, code) with atomicMin, so the message names the line the
// sequential pre-scan would have stopped at.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <mutex>
#include <thread>
#include "pv_prims.cuh"

namespace pv {

#ifndef PV_EMU
PV_DEV u32 pv_popc64(u64 x) { return (u32)__popcll((unsigned long long)x); }
PV_DEV u32 pv_ctz64(u64 x) { return (u32)(__ffsll((long long)x) - 1); }
#else
inline u32 pv_popc64(u64 x) { return (u32)__builtin_popcountll(x); }
inline u32 pv_ctz64(u64 x) { return (u32)__builtin_ctzll(x); }
#endif
// 0x80 in every byte of x that equals c (exact: no false positives next to a match)
PV_HD u64 byte_eq(u64 x, u8 c) {
	const u64 lo7 = 0x7F7F7F7F7F7F7F7Full;
	u64 y = x ^ (0x0101010101010101ull * c);
	u64 t = (y & lo7) + lo7;
	return ~(t | y | lo7);
}
// 0x80 in the bytes of the 8-byte word at absolute offset `at` that lie inside [lo, hi)
PV_HD u64 byte_range(u64 at, u64 lo, u64 hi) {
	u64 m = 0x8080808080808080ull;
	if (lo > at) { u64 k = lo - at; m = k >= 8 ? 0 : m & (~0ull << (8 * k)); }
	if (hi < at + 8) { u64 k = hi > at ? hi - at : 0; m = k == 0 ? 0 : m & (~0ull >> (8 * (8 - k))); }
	return m;
}
// first position in [from, to) holding c, else `to` (8 bytes per step once aligned)
PV_DEV u64 find_byte(const u8 *T, u64 from, u64 to, u8 c) {
	u64 p = from;
	while (p < to && (p & 7u)) { if (T[p] == c) return p; p++; }
	while (p + 8 <= to) { u64 m = byte_eq(*(const u64 *)(T + p), c); if (m) return p + (pv_ctz64(m) >> 3); p += 8; }
	while (p < to) { if (T[p] == c) return p; p++; }
	return to;
}
PV_DEV bool gfa_parse_u64(const u8 *T, u64 b, u64 e, u64 &out) {   // digits only, non-empty (pvh::parse_u64)
	if (b == e) return false;
	u64 v = 0;
	for (u64 q = b; q < e; q++) { u8 ch = T[q]; if (ch < '0' || ch > '9') return false; v = v * 10 + (u64)(ch - '0'); }
	out = v; return true;
}
PV_DEV u64 gfa_strtoull(const u8 *T, u64 b, u64 e) {   // what strtoull(.., 10) makes of a non-numeric segment name (parity unpinned)
	u64 q = b;
	while (q < e && (T[q] == ' ' || (T[q] >= 9 && T[q] <= 13))) q++;
	bool neg = false;
	if (q < e && (T[q] == '+' || T[q] == '-')) { neg = T[q] == '-'; q++; }
	u64 v = 0; bool over = false;
	for (; q < e && T[q] >= '0' && T[q] <= '9'; q++) {
		u64 d = (u64)(T[q] - '0');
		if (v > (~0ull - d) / 10) over = true;
		v = v * 10 + d;
	}
	if (over) return ~0ull;
	return neg ? (u64)(0 - v) : v;
}

struct u32x4 { u32 a, b, c, d; };
struct OpAdd4 { PV_HD u32x4 operator()(u32x4 x, u32x4 y) const { return u32x4{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; } };

// line i of the text: [b, e) without the terminator (and without a trailing '\r')
struct GfaLines {
	const u8 *T; const u64 *nl; u64 n, nNl;
	PV_DEV void get(u64 i, u64 &b, u64 &e) const {
		b = i ? nl[i - 1] + 1 : 0;
		e = i < nNl ? nl[i] : n;
		if (e > b && T[e - 1] == '\r') e--;
	}
};
// segment id -> vertex index: dense table when the ids are dense enough, else binary search (HostGraph::idx_of)
struct GfaIdMap {
	const u32 *table; u64 tableN; const u32 *ids; u32 V;
	PV_DEV u32 idx_of(u64 id) const {
		if (table) return id < tableN ? table[id] : NONE;
		if (id > 0xFFFFFFFFull) return NONE;
		u32 lo = 0, hi = V;
		while (lo < hi) { u32 mid = lo + ((hi - lo) >> 1); if (ids[mid] < (u32)id) lo = mid + 1; else hi = mid; }
		return (lo < V && ids[lo] == (u32)id) ? lo : NONE;
	}
};

// chunk g of the P lines' walk fields -> (line r, absolute offset of its 16 bytes)
struct GfaPLoc {
	const u64 *cb, *pB; u32 nP;
	PV_DEV void locate(u64 g, u32 &r, u64 &at) const {
		u32 lo = 0, hi = nP;
		while (hi - lo > 1) { u32 mid = lo + ((hi - lo) >> 1); if (cb[mid] <= g) lo = mid; else hi = mid; }
		r = lo; at = (pB[4 * (size_t)lo + 2] / 16 + (g - cb[lo])) * 16;
	}
};
// one walk step: token [q, te) with te = the next comma or the end of the walk
struct GfaStepper {
	const u8 *T; GfaIdMap M; u32 *st, *fl;
	PV_DEV void at(u64 q, u64 se, u64 idx) const {
		u64 te = q;
		while (te < se && T[te] != ',') te++;
		u64 id = 0; u32 v = NONE;
		bool ok = te - q >= 2 && (T[te - 1] == '+' || T[te - 1] == '-') && gfa_parse_u64(T, q, te - 1, id);
		if (ok) { v = M.idx_of(id); ok = v != NONE; }
		if (!ok) { atomic_or(fl + 1, 1u); st[idx] = 0; return; }
		st[idx] = (v << 1) | (T[te - 1] == '-' ? 1u : 0u);
	}
};

PV_KERNEL(G_nl_count, 20);
PV_KERNEL(G_nl_write, 20);
PV_KERNEL(G_classify, 24);
PV_KERNEL(G_split, 20);
PV_KERNEL(G_seg, 48);
PV_KERNEL(G_seg_sorted, 8);
PV_KERNEL(G_seg_first, 8);
PV_KERNEL(G_seg_gather, 32);
PV_KERNEL(G_len64, 12);
PV_KERNEL(G_lbl, 32);
PV_KERNEL(G_idtab, 8);
PV_KERNEL(G_link, 40);
PV_KERNEL(G_phead, 64);
PV_KERNEL(G_ptab, 16);
PV_KERNEL(G_pse, 32);
PV_KERNEL(G_pcount, 20);
PV_KERNEL(G_pstep, 24);
PV_KERNEL(G_haplen, 8);

// everything the session needs from one GFA: large arrays stay in HBM, small tables come to the host
struct GfaDev {
	u32 V = 0, H = 0; u64 E = 0, S = 0, textBytes = 0;
	Buf<u32> ids, segLen, ep1, ep2, steps; Buf<u64> lblOff; Buf<u8> lbl;
	std::vector<u32> h_ids, h_segLen, h_stepOff;
	std::vector<u64> hapLen, nameB, nameE;   // per P line: walk length in bases, byte span of the path name in the file
	// first error of the reference's pre-scan: (line number << 16) | (code << 8) | offending byte; ~0 = none
	u64 err = ~0ull; bool lq = false;   // lq: a link / path the liteseq stage rejects
	double ms_read = 0, ms_parse = 0;
};

// file -> device memory.  A few host threads pread(2) interleaved 1 MiB blocks of the file into page-locked staging buffers
// (two per thread: the read of one block overlaps the H2D of the previous one) and enqueue the copies on the context's
// stream, so the page-cache -> staging memcpy (the slow leg, ~3 GB/s per core) runs on several cores while PCIe streams.
// The staging buffers come from the per-process pool of pv_common.cuh.  T gets n + 32 bytes, the tail zeroed, so that 16-byte chunk loads never
// run off the allocation.
inline void gfa_text_to_device(Ctx &c, int fd, u64 n, Buf<u8> &T) {
	T = Buf<u8>(c, n + 32);
	c.memset0(T.p + n, 32);
#ifndef PV_EMU
	const size_t BLK = PinnedPool::BLK;
	const u64 nBlk = (n + BLK - 1) / BLK;
	if (nBlk == 0) return;
	if (n <= (8u << 20)) {   // small file: page-locking staging blocks (~0.3 ms per MiB, cold) would cost more than the copy itself
		std::vector<char> tmp((size_t)n);
		for (u64 got = 0; got < n;) { ssize_t r = pread(fd, tmp.data() + got, (size_t)(n - got), (off_t)got); if (r <= 0) throw std::runtime_error("read error"); got += (u64)r; }
		PV_CUDA_CHECK(cudaMemcpyAsync(T.p, tmp.data(), (size_t)n, cudaMemcpyHostToDevice, c.stream));
		PV_CUDA_CHECK(cudaStreamSynchronize(c.stream));
		return;
	}
	unsigned hw = std::thread::hardware_concurrency();
	int NT = (int)(hw / 2 > 8 ? 8 : (hw / 2 < 1 ? 1 : hw / 2));
	if (const char *e = getenv("POVU_B200_READ_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) NT = v; }
	if ((u64)NT > nBlk) NT = (int)nBlk;
	std::atomic<int> fail(0);
	u8 *dst = T.p; cudaStream_t stream = c.stream; const int dev = c.device;
	auto worker = [&](int tid) {
		if (cudaSetDevice(dev) != cudaSuccess) { fail = 2; return; }
		PinnedPool &pool = PinnedPool::inst();
		char *st[2] = {pool.get(), pool.get()}; cudaEvent_t ev[2] = {nullptr, nullptr}; bool used[2] = {false, false};
		if (!st[0] || !st[1] || cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess) fail = 2;
		for (u64 b = (u64)tid, k = 0; b < nBlk && !fail; b += (u64)NT, k++) {
			int s = (int)(k & 1);
			if (used[s]) cudaEventSynchronize(ev[s]);
			const u64 off = b * BLK; const size_t want = (size_t)((n - off) < BLK ? (n - off) : BLK);
			size_t got = 0;
			while (got < want) { ssize_t r = pread(fd, st[s] + got, want - got, (off_t)(off + got)); if (r <= 0) { fail = 1; break; } got += (size_t)r; }
			if (fail) break;
			if (cudaMemcpyAsync(dst + off, st[s], want, cudaMemcpyHostToDevice, stream) != cudaSuccess || cudaEventRecord(ev[s], stream) != cudaSuccess) { fail = 2; break; }
			used[s] = true;
		}
		for (int s = 0; s < 2; s++) { if (used[s]) cudaEventSynchronize(ev[s]); if (ev[s]) cudaEventDestroy(ev[s]); pool.put(st[s]); }
	};
	if (NT == 1) worker(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < NT; t++) th.emplace_back(worker, t);
		for (auto &t : th) t.join();
	}
	cudaStreamSynchronize(c.stream);
	if (fail == 1) throw std::runtime_error("read error");
	if (fail) { cudaGetLastError(); throw std::runtime_error("povu_b200: could not stage the GFA text for the device (page-locked memory / copy failed)"); }
#else
	for (u64 got = 0; got < n;) { ssize_t r = read(fd, T.p + got, (size_t)(n - got)); if (r <= 0) throw std::runtime_error("read error"); got += (u64)r; }
#endif
}

// Parses the text T[0, n) (device memory, padded as gfa_text_to_device does).  Never throws for malformed input: the
// caller turns out.err / out.lq / the counts into the reference's messages.
inline void gfa_parse_device(Ctx &c, const u8 *T, u64 n, bool last_is_nl, bool inc_labels, bool inc_refs, GfaDev &out) {
	Trace tr(c, "gfa");
	out = GfaDev(); out.textBytes = n;
	out.h_stepOff.assign(1, 0);
	if (n == 0) return;
	// ---- A: newline positions
	const u64 nChunks = (n + 15) / 16;
	if (nChunks >= 0xFFFFFFF0ull) throw std::runtime_error("povu_b200: GFA text too large for one device parse");
	Buf<u32> cnt(c, nChunks + 1), pos(c, nChunks + 1);
	{
		u32 *cnt_ = cnt.p;
		parallel_for<G_nl_count>(c, nChunks + 1, PV_LAMBDA(size_t j) {
			if (j == nChunks) { cnt_[j] = 0; return; }
			const u64 *w = (const u64 *)(T + 16 * j);
			cnt_[j] = pv_popc64(byte_eq(w[0], '\n')) + pv_popc64(byte_eq(w[1], '\n'));   // the zero padding holds no newline
		});
	}
	exclusive_scan(c, cnt.p, pos.p, nChunks + 1, OpAddU32(), 0u);
	const u64 nNl = pos.read(nChunks);
	Buf<u64> nl(c, nNl + 1);
	{
		u64 *nl_ = nl.p; const u32 *pos_ = pos.p;
		parallel_for<G_nl_write>(c, nChunks, PV_LAMBDA(size_t j) {
			const u64 *w = (const u64 *)(T + 16 * j);
			u32 o = pos_[j];
			for (int h = 0; h < 2; h++) {
				u64 m = byte_eq(w[h], '\n');
				while (m) { nl_[o++] = 16 * (u64)j + 8 * h + (pv_ctz64(m) >> 3); m &= m - 1; }
			}
		});
	}
	cnt.release(); pos.release();
	const u64 nLines = nNl + (last_is_nl ? 0 : 1);
	if (nLines >= 0xFFFFFFF0ull) throw std::runtime_error("povu_b200: too many lines for one device parse");
	GfaLines L{T, nl.p, n, nNl};
	tr.mark("A newlines");
	// ---- B: record types, the pre-scan's S-line validation, per-type line lists
	Buf<u32x4> kind(c, nLines + 1), kpos(c, nLines + 1);
	Buf<u64> err(c, 1); err.fill_ff();
	{
		u32x4 *kind_ = kind.p; u64 *err_ = err.p;
		parallel_for<G_classify>(c, nLines + 1, PV_LAMBDA(size_t i) {
			u32x4 k{0, 0, 0, 0};
			if (i < nLines) {
				u64 b, e; L.get(i, b, e);
				if (e > b) {
					u8 t = T[b]; u32 code = 0;
					if (t == 'S') {
						u64 t1 = find_byte(T, b, e, '\t');
						if (t1 == e) code = 1;
						else {
							u64 t2 = find_byte(T, t1 + 1, e, '\t');
							if (t2 == e) code = 2;
							else if (t2 + 1 >= e || T[t2 + 1] == '\t') code = 3;
							else k.a = 1;
						}
					} else if (t == 'L') k.b = 1;
					else if (t == 'P') k.c = 1;
					else if (t != 'H' && t != 'W') code = 4;
					if (code) atomic_min(err_, ((u64)(i + 1) << 16) | ((u64)code << 8) | t);
				}
			}
			kind_[i] = k;
		});
	}
	exclusive_scan(c, kind.p, kpos.p, nLines + 1, OpAdd4(), u32x4{0, 0, 0, 0});
	u32x4 tot; c.d2h_async(&tot, kpos.p + nLines, sizeof tot);
	out.err = err.read(0);
	if (out.err != ~0ull) return;
	const u32 nS = tot.a, nP = inc_refs ? tot.c : 0; const u64 nL = tot.b;
	if (nS == 0) return;   // "liteseq returned no vertices"
	Buf<u32> sLine(c, nS), lLine(c, nL ? nL : 1), pLine(c, tot.c ? tot.c : 1);
	{
		const u32x4 *kind_ = kind.p, *kpos_ = kpos.p; u32 *s_ = sLine.p, *l_ = lLine.p, *p_ = pLine.p;
		parallel_for<G_split>(c, nLines, PV_LAMBDA(size_t i) {
			u32x4 k = kind_[i], p = kpos_[i];
			if (k.a) s_[p.a] = (u32)i;
			if (k.b) l_[p.b] = (u32)i;
			if (k.c) p_[p.c] = (u32)i;
		});
	}
	kind.release(); kpos.release();
	tr.mark("B classify");
	// ---- C: segments in ascending id order
	Buf<u32> segId(c, nS), segLen(c, nS); Buf<u64> segB(c, nS);
	Buf<u32> flags(c, 4); flags.zero();   // [0] ids not strictly ascending in file order, [1] link / path rejected
	{
		u32 *id_ = segId.p, *len_ = segLen.p; u64 *sb_ = segB.p; const u32 *s_ = sLine.p;
		parallel_for<G_seg>(c, nS, PV_LAMBDA(size_t k) {
			u64 b, e; L.get(s_[k], b, e);
			u64 t1 = find_byte(T, b, e, '\t'), t2 = find_byte(T, t1 + 1, e, '\t');
			u64 sb = t2 + 1, se = find_byte(T, sb, e, '\t');
			u64 id;
			if (!gfa_parse_u64(T, t1 + 1, t2, id)) id = gfa_strtoull(T, t1 + 1, t2);
			id_[k] = (u32)id; sb_[k] = sb; len_[k] = (u32)(se - sb);
		});
		u32 *fl = flags.p;
		parallel_for<G_seg_sorted>(c, nS, PV_LAMBDA(size_t k) { if (k && id_[k - 1] >= id_[k]) atomic_or(fl, 1u); });
	}
	u32 V = nS;
	if (flags.read(0) & 1u) {
		// stable sort by id; the first S line of an id wins (std::stable_sort + std::unique in the host loader)
		Buf<u32> key(c, nS), ord(c, nS);
		u32 *key_ = key.p, *ord_ = ord.p; const u32 *id_ = segId.p;
		parallel_for<G_seg_first>(c, nS, PV_LAMBDA(size_t k) { key_[k] = id_[k]; ord_[k] = (u32)k; });
		radix_sort_pairs(c, key, ord, nS, 32);
		Buf<u32> first(c, nS), keep;
		u32 *first_ = first.p; key_ = key.p; ord_ = ord.p;
		parallel_for<G_seg_first>(c, nS, PV_LAMBDA(size_t k) { first_[k] = (k == 0 || key_[k] != key_[k - 1]) ? 1u : 0u; });
		V = compact_flags(c, first.p, nS, keep);
		Buf<u32> id2(c, V), len2(c, V); Buf<u64> sb2(c, V);
		u32 *id2_ = id2.p, *len2_ = len2.p; u64 *sb2_ = sb2.p; const u32 *keep_ = keep.p, *len_ = segLen.p; const u64 *sb_ = segB.p;
		parallel_for<G_seg_gather>(c, V, PV_LAMBDA(size_t i) { u32 k = ord_[keep_[i]]; id2_[i] = id_[k]; len2_[i] = len_[k]; sb2_[i] = sb_[k]; });
		c.sync();
		segId = std::move(id2); segLen = std::move(len2); segB = std::move(sb2);
	}
	sLine.release();
	out.V = V;
	out.h_ids.resize(V); out.h_segLen.resize(V);
	c.d2h_async(out.h_ids.data(), segId.p, (size_t)V * 4); c.d2h_async(out.h_segLen.data(), segLen.p, (size_t)V * 4);
	{
		Buf<u64> len64(c, (size_t)V + 1);
		u64 *l64 = len64.p; const u32 *len_ = segLen.p;
		parallel_for<G_len64>(c, (size_t)V + 1, PV_LAMBDA(size_t i) { l64[i] = i < V ? (u64)len_[i] : 0; });
		out.lblOff = Buf<u64>(c, (size_t)V + 1);
		exclusive_scan(c, len64.p, out.lblOff.p, (size_t)V + 1, OpAddU64(), (u64)0);
	}
	const u64 lblBytes = out.lblOff.read(V);   // also completes the two copies above
	if (inc_labels) {
		out.lbl = Buf<u8>(c, lblBytes + 16);
		u8 *lbl_ = out.lbl.p; const u64 *off_ = out.lblOff.p, *sb_ = segB.p;
		parallel_for<G_lbl>(c, (lblBytes + 15) / 16, PV_LAMBDA(size_t j) {
			u64 q = 16 * (u64)j, qe = q + 16 < lblBytes ? q + 16 : lblBytes;
			u32 lo = 0, hi = V;   // last vertex whose sequence starts at or before q
			while (hi - lo > 1) { u32 mid = lo + ((hi - lo) >> 1); if (off_[mid] <= q) lo = mid; else hi = mid; }
			u32 v = lo;
			for (; q < qe; q++) { while (q >= off_[v + 1]) v++; lbl_[q] = T[sb_[v] + (q - off_[v])]; }
		});
	}
	const u64 maxId = out.h_ids[V - 1];
	Buf<u32> table;
	if (maxId <= 4ull * V + 1024) {
		table = Buf<u32>(c, maxId + 1); table.fill_ff();
		u32 *t_ = table.p; const u32 *id_ = segId.p;
		parallel_for<G_idtab>(c, V, PV_LAMBDA(size_t i) { t_[id_[i]] = (u32)i; });
	}
	GfaIdMap M{table.p, table.p ? maxId + 1 : 0, segId.p, V};
	tr.mark("C segments");
	// ---- D: links
	out.E = nL;
	out.ep1 = Buf<u32>(c, nL ? nL : 1); out.ep2 = Buf<u32>(c, nL ? nL : 1);
	{
		u32 *e1 = out.ep1.p, *e2 = out.ep2.p, *fl = flags.p; const u32 *l_ = lLine.p;
		parallel_for<G_link>(c, nL, PV_LAMBDA(size_t k) {
			u64 b, e; L.get(l_[k], b, e);
			u64 fb[6], fe[6]; bool ok = true;
			u64 q = b;
			for (int f = 0; f < 6; f++) {
				if (q > e) { ok = false; break; }
				u64 t = find_byte(T, q, e, '\t');
				fb[f] = q; fe[f] = t; q = t + 1;   // after the last field q = e + 1: a seventh field would be missing, six are not
			}
			u64 a = 0, z = 0; u32 ia = NONE, iz = NONE;
			if (ok) {
				u64 ol = fe[5] - fb[5];
				ok = (ol == 2 && T[fb[5]] == '0' && T[fb[5] + 1] == 'M') || (ol == 1 && T[fb[5]] == '*');
			}
			if (ok) ok = gfa_parse_u64(T, fb[1], fe[1], a) && gfa_parse_u64(T, fb[3], fe[3], z);
			if (ok) { ia = M.idx_of(a); iz = M.idx_of(z); ok = ia != NONE && iz != NONE; }
			if (!ok) { atomic_or(fl + 1, 1u); e1[k] = 0; e2[k] = 0; return; }
			bool o1 = fe[2] - fb[2] == 1 && T[fb[2]] == '+', o2 = fe[4] - fb[4] == 1 && T[fb[4]] == '+';
			e1[k] = 2 * ia + (o1 ? 1u : 0u);   // `L a + b +` joins a.right to b.left; '-' flips the side
			e2[k] = 2 * iz + (o2 ? 0u : 1u);
		});
	}
	lLine.release();
	tr.mark("D links");
	// ---- E: paths
	out.H = nP;
	if (nP) {
		Buf<u64> pB(c, (size_t)nP * 4), tab3(c, nP);   // per P line: name begin, name end (= tab 2), walk begin, line end
		tab3.fill_ff();
		{
			u64 *pB_ = pB.p; u32 *fl = flags.p; const u32 *p_ = pLine.p;
			parallel_for<G_phead>(c, nP, PV_LAMBDA(size_t r) {
				u64 b, e; L.get(p_[r], b, e);
				u64 t1 = find_byte(T, b, e, '\t');
				u64 t2 = t1 < e ? find_byte(T, t1 + 1, e, '\t') : e;
				if (t2 >= e) { atomic_or(fl + 1, 1u); pB_[4 * r] = pB_[4 * r + 1] = pB_[4 * r + 2] = pB_[4 * r + 3] = e; return; }
				pB_[4 * r] = t1 + 1; pB_[4 * r + 1] = t2; pB_[4 * r + 2] = t2 + 1; pB_[4 * r + 3] = e;
			});
		}
		std::vector<u64> hB = pB.to_host();
		out.nameB.resize(nP); out.nameE.resize(nP);
		std::vector<u64> chunkBase((size_t)nP + 1, 0);   // 16-byte chunks (absolute alignment) that cover a walk field up to the line end
		for (u32 r = 0; r < nP; r++) {
			out.nameB[r] = hB[4 * (size_t)r]; out.nameE[r] = hB[4 * (size_t)r + 1];
			u64 sb = hB[4 * (size_t)r + 2], e = hB[4 * (size_t)r + 3];
			chunkBase[r + 1] = chunkBase[r] + (e > sb ? (e - 1) / 16 - sb / 16 + 1 : 0);
		}
		const u64 nPC = chunkBase[nP];
		if (nPC >= 0xFFFFFFF0ull) throw std::runtime_error("povu_b200: path lines too large for one device parse");
		Buf<u64> cb(c, (size_t)nP + 1);
		c.h2d(cb.p, chunkBase.data(), ((size_t)nP + 1) * 8);
		const u64 *cb_ = cb.p, *pB_ = pB.p; u64 *tab3_ = tab3.p;
		GfaPLoc PL{cb_, pB_, nP};
		parallel_for<G_ptab>(c, nPC, PV_LAMBDA(size_t g) {
			u32 r; u64 at; PL.locate(g, r, at);
			const u64 sb = pB_[4 * (size_t)r + 2], e = pB_[4 * (size_t)r + 3];
			const u64 *w = (const u64 *)(T + at);
			for (int h = 0; h < 2; h++) {
				u64 m = byte_eq(w[h], '\t') & byte_range(at + 8 * h, sb, e);
				if (m) { atomic_min(tab3_ + r, at + 8 * h + (pv_ctz64(m) >> 3)); break; }
			}
		});
		{
			u32 *fl = flags.p;
			// the walk is field 2 of 4: a line without the overlaps column, or with an empty walk, is rejected
			parallel_for<G_pse>(c, nP, PV_LAMBDA(size_t r) {
				u64 sb = pB_[4 * r + 2];
				if (tab3_[r] == ~0ull || tab3_[r] == sb) { atomic_or(fl + 1, 1u); tab3_[r] = sb; }
			});
		}
		Buf<u32> cc(c, nPC + 1), cpos(c, nPC + 1);
		{
			u32 *cc_ = cc.p;
			parallel_for<G_pcount>(c, nPC + 1, PV_LAMBDA(size_t g) {
				if (g == nPC) { cc_[g] = 0; return; }
				u32 r; u64 at; PL.locate(g, r, at);
				const u64 sb = pB_[4 * (size_t)r + 2], se = tab3_[r];
				u32 k = 0;
				if (at < se) {
					const u64 *w = (const u64 *)(T + at);
					k = pv_popc64(byte_eq(w[0], ',') & byte_range(at, sb, se)) + pv_popc64(byte_eq(w[1], ',') & byte_range(at + 8, sb, se));
				}
				cc_[g] = k;
			});
		}
		exclusive_scan(c, cc.p, cpos.p, nPC + 1, OpAddU32(), 0u);
		cc.release();
		const u64 nCommas = cpos.read(nPC);
		if (flags.read(1)) { out.lq = true; return; }
		const u64 S = nCommas + nP;
		out.S = S;
		if (S >= 0xFFFFFFF0ull) return;   // the caller reports the 2^32 step limit
		out.steps = Buf<u32>(c, S ? S : 1);
		Buf<u32> stepOff(c, (size_t)nP + 1);
		{
			u32 *st_ = out.steps.p, *so_ = stepOff.p, *fl = flags.p; const u32 *cpos_ = cpos.p;
			GfaStepper SA{T, M, st_, fl};
			parallel_for<G_pstep>(c, nPC, PV_LAMBDA(size_t g) {
				u32 r; u64 at; PL.locate(g, r, at);
				const u64 sb = pB_[4 * (size_t)r + 2], se = tab3_[r];
				u64 o = cpos_[g];   // commas before this chunk, all lines
				if (g == cb_[r]) { so_[r] = (u32)(o + r); SA.at(sb, se, o + r); if (r + 1 == nP) so_[nP] = (u32)(nCommas + nP); }
				if (at >= se) return;
				const u64 *w = (const u64 *)(T + at);
				for (int h = 0; h < 2; h++) {
					u64 m = byte_eq(w[h], ',') & byte_range(at + 8 * h, sb, se);
					while (m) { u64 p = at + 8 * h + (pv_ctz64(m) >> 3); SA.at(p + 1, se, o + r + 1); o++; m &= m - 1; }
				}
			});
		}
		out.h_stepOff = stepOff.to_host();
		if (flags.read(1)) { out.lq = true; return; }
		// contig lengths: sum of the segment lengths along the walk
		Buf<u64> hl(c, nP); hl.zero();
		{
			const u32 *st_ = out.steps.p, *so_ = stepOff.p, *len_ = segLen.p; u64 *hl_ = hl.p;
			const u64 RUN = 256;
			parallel_for<G_haplen>(c, (S + RUN - 1) / RUN, PV_LAMBDA(size_t t) {
				u64 k = t * RUN, ke = k + RUN < S ? k + RUN : S;
				u32 lo = 0, hi = nP;
				while (hi - lo > 1) { u32 mid = lo + ((hi - lo) >> 1); if (so_[mid] <= k) lo = mid; else hi = mid; }
				u32 r = lo; u64 acc = 0;
				for (; k < ke; k++) {
					while (k >= so_[r + 1]) { if (acc) atomic_add(hl_ + r, acc); acc = 0; r++; }
					acc += len_[st_[k] >> 1];
				}
				if (acc) atomic_add(hl_ + r, acc);
			});
		}
		out.hapLen = hl.to_host();
	} else if (flags.read(1)) { out.lq = true; return; }
	tr.mark("E paths");
	out.ids = std::move(segId); out.segLen = std::move(segLen);
	c.sync();
}

} // namespace pv
