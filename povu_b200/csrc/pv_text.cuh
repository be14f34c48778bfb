// povu_b200 -- VCF 4.2 record text written on the device: binary records (pv_call.cuh)
// -> the exact bytes the reference prints, two kernels over the same emitter:
//   X_len    one thread per record, counting sink          -> line lengths -> u64 exclusive scan = line offsets
//   X_write  one thread per record, word-buffered sink     -> the text, each line at its offset
// Replaces the reference's
//   VcfRec text / to_vcf::write_rec   /root/reference/include/ita/genomics/vcf.hpp:295-383, src/mto/to_vcf.cpp:84-119
//   hap_slice::{as_str, as_dna_str}   /root/reference/include/ita/genomics/allele.hpp:134-209
//   pu::reverse_complement            /root/reference/src/povu/common/utils.cpp:14-45
//   VcfRec::get_af ("{:.1f}")         /root/reference/include/ita/genomics/vcf.hpp:310-332
//   gen_gt_data (GT slot order)       /root/reference/src/ita/genomics/vcf.cpp:51-84
//   record order                      /root/reference/src/ita/genomics/genomics.cpp:114-156, src/mto/to_vcf.cpp:121-139
// HBM-bound byte work: every record reads its few steps / labels / genotype codes and writes ~170 B (16 paths) to ~1 KB
// (464 paths) of text; stores leave a thread as aligned 8-byte words.
#pragma once
#include <cmath>
#include <cerrno>
#include <unistd.h>
#include "pv_call.cuh"

namespace pv {

PV_KERNEL(X_key, 16);
PV_KERNEL(X_sorted, 12);
PV_KERNEL(X_len, 96);      // record fields + genotype codes in, one length out
PV_KERNEL(X_write, 272);   // the same reads + ~174 B of text out (cfg2 average)
PV_KERNEL(X_keep, 8);

// everything the emitter reads, device pointers only
struct TextView {
	// graph + paths
	const u32 *ids, *segLen, *steps, *stepOff; const u64 *lblOff; const u8 *lbl;
	const u32 *tagOff; const u8 *tags;
	// genotype columns: per component the (column, phase, local path index) triples sorted by (column, phase)
	const u32 *ploidy; u32 nCols;
	const u32 *ceOff, *ceCol, *cePhase, *ceLocal;
	// PVST forest
	const u32 *fa, *fz, *fheight; const u8 *faf, *fzf;
	// records
	const u32 *rov_fl, *rov_comp, *rov_tangled, *cpOff;
	const u32 *rec_rov, *rec_ref, *rec_cxtL, *rec_cxtR, *rec_vt, *rec_pos, *rec_refStart, *rec_refLen, *rec_grpOff; const u64 *rec_gtOff;
	const u32 *grp_hap, *grp_start, *grp_len, *grp_count; const u16 *gt;
};

struct CountSink {
	u64 n = 0;
	PV_DEV void ch(char) { n++; }
	PV_DEV void bytes(const u8 *, u32 k) { n += k; }
	PV_DEV void finish() {}
};
// bytes are collected in a 64-bit register and leave as aligned 8-byte stores; the partial words at both ends of a
// line (shared with the neighbouring lines) are written byte by byte
struct WriteSink {
	u8 *base; u64 acc; u32 fill, lo;   // base: 8-byte aligned; acc holds bytes [lo, fill) of the current word
	PV_DEV explicit WriteSink(u8 *p) {
		uintptr_t a = (uintptr_t)p; u32 off = (u32)(a & 7u);
		base = p - off; acc = 0; fill = off; lo = off;
	}
	PV_DEV void flush() {
		if (lo == 0 && fill == 8) *(u64 *)base = acc;
		else for (u32 k = lo; k < fill; k++) base[k] = (u8)(acc >> (8 * k));
		base += 8; acc = 0; fill = 0; lo = 0;
	}
	PV_DEV void ch(char c) { acc |= (u64)(u8)c << (8 * fill); if (++fill == 8) flush(); }
	PV_DEV void bytes(const u8 *s, u32 k) { for (u32 i = 0; i < k; i++) ch((char)s[i]); }
	PV_DEV void finish() { if (fill > lo) flush(); }
};

template <class Sink> PV_DEV void put_lit(Sink &o, const char *s) { while (*s) o.ch(*s++); }
template <class Sink> PV_DEV void put_u32(Sink &o, u32 v) {
	char b[10]; int n = 0;
	do { b[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
	while (n) o.ch(b[--n]);
}
PV_DEV char text_comp_base(char c) {
	switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; default: return c; }
}
template <class Sink> PV_DEV void put_step_label(const TextView &t, Sink &o, u32 step) {   // utils.cpp:14-45 for '<'
	u32 v = step >> 1; const u8 *l = t.lbl + t.lblOff[v]; u32 n = t.segLen[v];
	if (!(step & 1u)) o.bytes(l, n);
	else for (u32 i = n; i-- > 0;) o.ch(text_comp_base((char)l[i]));
}
template <class Sink> PV_DEV void put_step_str(const TextView &t, Sink &o, u32 step) { o.ch((step & 1u) ? '<' : '>'); put_u32(o, t.ids[step >> 1]); }
// hap_slice::as_str / as_dna_str (allele.hpp:134-209); start = step index inside path `hap`
template <class Sink> PV_DEV void put_slice_str(const TextView &t, Sink &o, u32 hap, u32 start, u32 len, u32 vt) {
	u32 i = t.stepOff[hap] + start, N = i + len;
	if (vt == VT_SUB) { i++; N--; } else if (vt == VT_INS || vt == VT_DEL) N--;
	for (; i < N; i++) put_step_str(t, o, t.steps[i]);
}
template <class Sink> PV_DEV void put_slice_dna(const TextView &t, Sink &o, u32 hap, u32 start, u32 len, u32 vt) {
	u32 i = t.stepOff[hap] + start, N = i + len;
	if (vt == VT_INS || vt == VT_DEL) { u32 s = t.steps[i]; u32 v = s >> 1; u32 n = t.segLen[v]; const u8 *l = t.lbl + t.lblOff[v]; o.ch(!(s & 1u) ? (char)l[n - 1] : text_comp_base((char)l[0])); }
	if (vt != VT_SUBR) { i++; N--; }
	for (; i < N; i++) put_step_label(t, o, t.steps[i]);
}
// "{:.1f}" of double(ac) / double(an): the correctly rounded (ties to even) one-decimal rendering of the BINARY quotient q.
// The nearest tenth m/10 satisfies (2m-1)/20 <= q <= (2m+1)/20; fma(q, 20, -(2m+1)) has the sign of the exact difference
// 20q - (2m+1) and is zero only on an exact tie (q = 0.25, 0.75: the dyadic odd/20), which goes to the even digit.
template <class Sink> PV_DEV void put_af(Sink &o, u32 ac, u32 an) {
	double q = (double)ac / (double)an;
	if (!(q <= 4.0e8)) { put_lit(o, q != q ? "nan" : "inf"); return; }   // an == 0 cannot happen (a record has >= 1 allele); fmt's spelling
	u64 m = (u64)(q * 10.0);
	while (m > 0 && fma(q, 20.0, -(double)(2 * m - 1)) <= 0.0) m--;
	for (;;) {
		double d = fma(q, 20.0, -(double)(2 * m + 1));
		if (d > 0.0) { m++; continue; }
		if (d == 0.0 && (m & 1u)) m++;
		break;
	}
	put_u32(o, (u32)(m / 10)); o.ch('.'); o.ch((char)('0' + m % 10));
}
template <class Sink> PV_DEV void put_flubble_label(const TextView &t, Sink &o, u32 fl) {
	o.ch(t.faf[fl] ? '>' : '<'); put_u32(o, t.ids[t.fa[fl]]); o.ch(t.fzf[fl] ? '>' : '<'); put_u32(o, t.ids[t.fz[fl]]);
}

// one VCF line (vcf.hpp:295-383 + to_vcf.cpp:84-119)
template <class Sink> PV_DEV void emit_record(const TextView &t, size_t i, Sink &o) {
	const u32 vt = t.rec_vt[i], ref = t.rec_ref[i], rov = t.rec_rov[i];
	const u32 g0 = t.rec_grpOff[i], g1 = t.rec_grpOff[i + 1];
	const u32 comp = t.rov_comp[rov];
	const u32 nh = t.cpOff[comp + 1] - t.cpOff[comp];
	const u16 *gt = t.gt + t.rec_gtOff[i];
	u32 nref = 0;
	for (u32 k = 0; k < nh; k++) if (gt[k] == 0) nref++;
	u32 an = nref;
	for (u32 k = g0; k < g1; k++) an += t.grp_count[k];
	o.bytes(t.tags + t.tagOff[ref], t.tagOff[ref + 1] - t.tagOff[ref]); o.ch('\t');
	put_u32(o, t.rec_pos[i]); o.ch('\t');
	put_step_str(t, o, t.rec_cxtL[i]); put_step_str(t, o, t.rec_cxtR[i]); o.ch('\t');
	put_slice_dna(t, o, ref, t.rec_refStart[i], t.rec_refLen[i], vt); o.ch('\t');
	for (u32 k = g0; k < g1; k++) { if (k > g0) o.ch(','); put_slice_dna(t, o, t.grp_hap[k], t.grp_start[k], t.grp_len[k], vt); }
	put_lit(o, "\t60\tPASS\tAC=");
	for (u32 k = g0; k < g1; k++) { if (k > g0) o.ch(','); put_u32(o, t.grp_count[k]); }
	put_lit(o, ";AF=");
	for (u32 k = g0; k < g1; k++) { if (k > g0) o.ch(','); put_af(o, t.grp_count[k], an); }
	// NS = genotype columns that received any call (vcf.cpp:51-84: a path with data marks its column even when its
	// phase lies outside the column's ploidy)
	const u32 e0 = t.ceOff[comp], e1 = t.ceOff[comp + 1];
	u32 ns = 0;
	for (u32 e = e0; e < e1;) {
		u32 c = t.ceCol[e]; bool has = false;
		for (; e < e1 && t.ceCol[e] == c; e++) if (gt[t.ceLocal[e]] != 0xFFFFu) has = true;
		ns += has ? 1u : 0u;
	}
	put_lit(o, ";AN="); put_u32(o, an);
	put_lit(o, ";NS="); put_u32(o, ns);
	put_lit(o, ";AT="); put_slice_str(t, o, ref, t.rec_refStart[i], t.rec_refLen[i], vt);
	o.ch(',');
	for (u32 k = g0; k < g1; k++) { if (k > g0) o.ch(','); put_slice_str(t, o, t.grp_hap[k], t.grp_start[k], t.grp_len[k], vt); }
	put_lit(o, ";VARTYPE="); put_lit(o, vt == VT_DEL ? "DEL" : (vt == VT_INS ? "INS" : (vt == VT_SUB ? "SUB" : "SUBR")));
	put_lit(o, ";TANGLED="); o.ch(t.rov_tangled[rov] ? 'T' : 'F');
	if (vt != VT_SUBR) {
		const u32 fl = t.rov_fl[rov];
		put_lit(o, ";ES="); put_flubble_label(t, o, fl);
		put_lit(o, ";LV="); put_u32(o, t.fheight[fl] - 1);
	}
	put_lit(o, "\tGT\t");
	// genotype columns.  The reference fills the slots ref matches first, then alt group 1, 2, ... (each in path order),
	// later writes replacing earlier ones: a slot shared by several paths of the component ends up with the largest code.
	u32 e = e0;
	for (u32 c = 0; c < t.nCols; c++) {
		if (c) o.ch('\t');
		u32 f = e;
		while (f < e1 && t.ceCol[f] == c) f++;
		const u32 pl = t.ploidy[c];
		bool any = false;
		for (u32 x = e; x < f; x++) if (t.cePhase[x] < pl && gt[t.ceLocal[x]] != 0xFFFFu) { any = true; break; }
		if (!any) o.ch('.');
		else {
			u32 x = e;
			for (u32 p = 0; p < pl; p++) {
				if (p) o.ch('|');
				u32 code = NONE;
				for (; x < f && t.cePhase[x] == p; x++) { u32 v = gt[t.ceLocal[x]]; if (v != 0xFFFFu && (code == NONE || v > code)) code = v; }
				if (code == NONE) o.ch('.'); else put_u32(o, code);
			}
		}
		e = f;
	}
	o.ch('\n');
	o.finish();
}

// device-resident tables that depend only on the loaded graph (built once per session)
struct TextTables {
	Buf<u32> ids, tagOff, ploidy, ceOff, ceCol, cePhase, ceLocal; Buf<u64> lblOff; Buf<u8> lbl, tags;
	u32 nCols = 0; bool ready = false;
};

struct TextOut {
	hvec<char> text;            // page-locked: the D2H runs at PCIe speed
	hvec<u64> lineOff;          // n + 1 offsets into text (kept lines only have non-zero length)
	size_t n = 0;
};

// order of the reference (genomics.cpp:114-156, to_vcf.cpp:121-139): chunks of `chunk_size` regions; inside a chunk by
// ref path id, then region, then the device's emission order (context / INS, SUB, DEL).  SUBR records hang off
// pseudo regions after the real ones and are emitted with the final chunk.  perm[k] = record index of output line k.
inline void record_order_device(Ctx &c, const u32 *rec_rov, const u32 *rec_ref, size_t n, u64 n_regions, size_t chunk_size, u32 H, Buf<u32> &perm) {
	perm = Buf<u32>(c, n ? n : 1);
	if (!n) return;
	if (chunk_size == 0) chunk_size = 1;
	const u32 last = n_regions ? (u32)(n_regions - 1) : 0u;
	const u64 cs = chunk_size;
	Buf<u32> flag(c, 4);
	flag.zero();
	u32 *perm_ = perm.p, *flag_ = flag.p;
	parallel_for<X_sorted>(c, n, PV_LAMBDA(size_t i) {
		perm_[i] = (u32)i;
		if (i == 0) return;
		u32 ra = rec_rov[i - 1], rb = rec_rov[i];
		u64 ca = (u64)(ra < last ? ra : last) / cs, cb = (u64)(rb < last ? rb : last) / cs;
		bool ok = ca != cb ? ca < cb : (rec_ref[i - 1] != rec_ref[i] ? rec_ref[i - 1] < rec_ref[i] : ra <= rb);
		if (!ok) flag_[0] = 1;
	});
	if (!flag.read(0)) return;   // a single ref path without inversions: already in order
	Buf<u32> key(c, n);
	u32 *key_ = key.p;
	u32 maxRov = (u32)std::min<u64>(0xFFFFFFFFull, n_regions + H + 1);
	for (int pass = 0; pass < 3; pass++) {   // LSD: region, ref path, chunk
		const u32 *p = perm.p;
		parallel_for<X_key>(c, n, PV_LAMBDA(size_t i) {
			u32 r = p[i]; u32 rv = rec_rov[r];
			key_[i] = pass == 0 ? rv : (pass == 1 ? rec_ref[r] : (u32)((u64)(rv < last ? rv : last) / cs));
		});
		radix_sort_pairs(c, key, perm, n, pass == 0 ? bits_for(maxRov) : (pass == 1 ? bits_for(H) : bits_for((u64)last / cs + 1)));
		key_ = key.p;
	}
}

// device text -> file descriptor through a small ring of page-locked blocks: the D2H of block k+1 runs on the copy stream
// while block k is written with write(2); no text-sized page-locked buffer is ever allocated
inline void stream_text_to_fd(Ctx &c, const u8 *d_text, u64 total, int fd) {
	auto put = [&](const char *p, size_t n) { while (n) { ssize_t w = write(fd, p, n); if (w < 0) { if (errno == EINTR) continue; throw std::runtime_error(std::string("write failed: ") + strerror(errno)); } p += w; n -= (size_t)w; } };
#ifndef PV_EMU
	const size_t BLK = PinnedPool::BLK; const int NB = 4;
	PinnedPool &pool = PinnedPool::inst();
	char *blk[NB]; cudaEvent_t ev[NB]; size_t len[NB];
	for (int k = 0; k < NB; k++) { blk[k] = pool.get(); if (!blk[k]) throw std::runtime_error("povu_b200: no page-locked memory for the text stream"); PV_CUDA_CHECK(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming)); }
	c.bg_sync(); c.bg_begin();
	const u64 nBlk = (total + BLK - 1) / BLK;
	try {
		for (u64 b = 0; b < nBlk + NB - 1; b++) {
			if (b < nBlk) {   // issue block b
				int k = (int)(b % NB); len[k] = (size_t)((total - b * BLK) < BLK ? (total - b * BLK) : BLK);
				PV_CUDA_CHECK(cudaMemcpyAsync(blk[k], d_text + b * BLK, len[k], cudaMemcpyDeviceToHost, c.copy_stream));
				PV_CUDA_CHECK(cudaEventRecord(ev[k], c.copy_stream));
			}
			if (b >= (u64)NB - 1) {   // drain block b - (NB - 1)
				int k = (int)((b - (NB - 1)) % NB);
				PV_CUDA_CHECK(cudaEventSynchronize(ev[k]));
				put(blk[k], len[k]);
			}
		}
	} catch (...) { cudaStreamSynchronize(c.copy_stream); for (int k = 0; k < NB; k++) { pool.put(blk[k]); cudaEventDestroy(ev[k]); } c.bg_pending = false; throw; }
	for (int k = 0; k < NB; k++) { pool.put(blk[k]); cudaEventDestroy(ev[k]); }
	c.bg_sync();
#else
	put((const char *)d_text, (size_t)total);
#endif
}

// text of the records selected by keepRef (per ref path id: 1 = print; nullptr = all), in the reference's order.
// fd < 0: the text and its line offsets land in `out`; fd >= 0: the text is streamed to that descriptor instead (out.text stays
// empty, *streamed = bytes written)
inline void records_text_device(Ctx &c, const TextView &t, size_t n, u64 n_regions, size_t chunk_size, u32 H, const u8 *keepRef /*host, H entries*/, TextOut &out,
				int fd = -1, u64 *streamed = nullptr) {
	out.n = n; out.text.clear(); out.lineOff.clear();
	if (streamed) *streamed = 0;
	if (fd < 0) out.lineOff.resize(n + 1);
	if (!n) { if (fd < 0) out.lineOff[0] = 0; return; }
	Buf<u32> perm;
	record_order_device(c, t.rec_rov, t.rec_ref, n, n_regions, chunk_size, H, perm);
	Buf<u8> keep(c, H ? H : 1);
	if (keepRef) c.h2d(keep.p, keepRef, H);
	const u8 *keep_ = keepRef ? keep.p : nullptr; const u32 *perm_ = perm.p;
	Buf<u64> len(c, n + 1), off(c, n + 1);
	u64 *len_ = len.p;
	parallel_for<X_len>(c, n + 1, PV_LAMBDA(size_t k) {
		if (k >= n) { len_[k] = 0; return; }
		u32 i = perm_[k];
		if (keep_ && !keep_[t.rec_ref[i]]) { len_[k] = 0; return; }
		CountSink s; emit_record(t, i, s); len_[k] = s.n;
	});
	exclusive_scan(c, len.p, off.p, n + 1, OpAddU64(), (u64)0);
	u64 total; c.d2h(&total, off.p + n, 8);
	Buf<u8> text(c, total + 8);
	u8 *text_ = text.p; const u64 *off_ = off.p;
	parallel_for<X_write>(c, n, PV_LAMBDA(size_t k) {
		if (off_[k + 1] == off_[k]) return;
		WriteSink s(text_ + off_[k]); emit_record(t, perm_[k], s);
	});
	if (fd >= 0) {
		stream_text_to_fd(c, text.p, total, fd);
		if (streamed) *streamed = total;
		return;
	}
	out.text.resize(total);
	c.d2h_async(out.text.data(), text.p, total);
	c.d2h_async(out.lineOff.data(), off.p, (n + 1) * 8);
	c.sync();
}

// ---------------------------------------------------------------------------------------------
// .pvst v0.0.3 text written on the device (second half of SURVEY.md 8 f2; replaces mto::to_pvst::write_pvst,
// /root/reference/src/mto/to_pvst.cpp:23-109): per component file
//     H\t0.0.3\t.\t.\t.\n
//     D\t0\t.\t<children>\t.\n
//     <family>\t<idx>\t<label>\t<children>\tL\n        children = "." or "c1, c2, ..." in creation order
// A top-level chain gives the dummy root a million children, so the unit of work is not a line but an ITEM: per PVST
// vertex a prefix item (everything up to the children column) and a suffix item, plus one item per child entry.  Items
// are laid out in file order (vertex v owns items 2v + childOff[v] .. 2v + childOff[v+1] + 1), every item computes its
// length, one u64 scan places them, every item writes its few bytes.
// ---------------------------------------------------------------------------------------------
PV_KERNEL(P_node, 12);
PV_KERNEL(P_len, 40);
PV_KERNEL(P_write, 64);
struct PvstView {
	const u32 *ids, *a_v, *z_v, *parent, *nodeTree, *treeNode0; const u8 *a_fwd, *z_fwd, *family;   // flubble arrays by global flubble index
	const u32 *childOff, *childList;   // per node (global): its child flubbles (global flubble indices), ascending
	u32 nNodes, nTrees;
	// node x of tree t: x == treeNode0[t] is the dummy root, otherwise flubble x - t - 1 (global), local PVST index x - treeNode0[t]
	template <class Sink> PV_DEV void prefix(u32 x, Sink &o) const {
		u32 t = nodeTree[x], loc = x - treeNode0[t];
		if (loc == 0) { put_lit(o, "H\t0.0.3\t.\t.\t.\nD\t0\t.\t"); }
		else {
			u32 f = x - t - 1;
			o.ch((char)family[f]); o.ch('\t'); put_u32(o, loc); o.ch('\t');
			o.ch(a_fwd[f] ? '>' : '<'); put_u32(o, ids[a_v[f]]); o.ch(z_fwd[f] ? '>' : '<'); put_u32(o, ids[z_v[f]]); o.ch('\t');
		}
		if (childOff[x] == childOff[x + 1]) o.ch('.');
	}
	template <class Sink> PV_DEV void suffix(u32 x, Sink &o) const { put_lit(o, x == treeNode0[nodeTree[x]] ? "\t.\n" : "\tL\n"); }
	template <class Sink> PV_DEV void child(u32 x, u32 k, Sink &o) const {   // k-th entry of the global child list, a child of node x
		if (k > childOff[x]) { o.ch(','); o.ch(' '); }
		u32 f = childList[k]; u32 t = nodeTree[x];
		put_u32(o, f + t + 1 - treeNode0[t]);
	}
};
struct PvstParentNode {   // key of flubble f for the children grouping: the global node index of its parent
	const u32 *parent, *flTree, *treeNode0;
	PV_DEV u32 operator()(size_t f) const { return treeNode0[flTree[f]] + parent[f]; }
};
struct PvstTextOut { hvec<char> text; std::vector<u64> treeOff; };   // treeOff[t] .. treeOff[t+1]: the file of tree t

// forest arrays are HOST arrays (flubbles of all trees concatenated; h_treeOff[T+1] = first flubble of every tree; parent = local
// PVST index, 0 = root); ids is the device segment-id table
inline void pvst_text_device(Ctx &c, const u32 *d_ids, u32 nFl, u32 T, const u32 *h_treeOff, const u32 *h_a_v, const u8 *h_a_fwd, const u32 *h_z_v,
			     const u8 *h_z_fwd, const u32 *h_parent, const char *h_family, PvstTextOut &out) {
	out.text.clear(); out.treeOff.assign((size_t)T + 1, 0);
	if (T == 0) return;
	const u32 N = nFl + T;
	std::vector<u32> node0((size_t)T + 1);
	for (u32 t = 0; t <= T; t++) node0[t] = h_treeOff[t] + t;
	Buf<u32> a_v(c, nFl ? nFl : 1), z_v(c, nFl ? nFl : 1), parent(c, nFl ? nFl : 1), treeOff(c, (size_t)T + 1), treeNode0(c, (size_t)T + 1), nodeTree(c, N), flTree(c, nFl ? nFl : 1);
	Buf<u8> a_fwd(c, nFl ? nFl : 1), z_fwd(c, nFl ? nFl : 1), family(c, nFl ? nFl : 1);
	c.h2d(a_v.p, h_a_v, (size_t)nFl * 4); c.h2d(z_v.p, h_z_v, (size_t)nFl * 4); c.h2d(parent.p, h_parent, (size_t)nFl * 4);
	c.h2d(a_fwd.p, h_a_fwd, nFl); c.h2d(z_fwd.p, h_z_fwd, nFl); c.h2d(family.p, h_family, nFl);
	c.h2d(treeOff.p, h_treeOff, ((size_t)T + 1) * 4); c.h2d(treeNode0.p, node0.data(), ((size_t)T + 1) * 4);
	{
		u32 *nodeTree_ = nodeTree.p, *flTree_ = flTree.p; const u32 *n0 = treeNode0.p;
		parallel_for<P_node>(c, N, PV_LAMBDA(size_t x) {
			u32 t = ub_u32(n0, 0, T + 1, (u32)x) - 1;
			nodeTree_[x] = t;
			if (x != n0[t]) flTree_[x - t - 1] = t;
		});
	}
	Buf<u32> childOff(c, (size_t)N + 1), childList(c, nFl ? nFl : 1);
	group_indices(c, PvstParentNode{parent.p, flTree.p, treeNode0.p}, nFl, N, childOff.p, childList.p);
	PvstView pw{d_ids, a_v.p, z_v.p, parent.p, nodeTree.p, treeNode0.p, a_fwd.p, z_fwd.p, (const u8 *)family.p, childOff.p, childList.p, N, T};
	const size_t nItems = 2 * (size_t)N + nFl;
	Buf<u64> len(c, nItems + 1), off(c, nItems + 1);
	u64 *len_ = len.p; const u32 *childOff_ = childOff.p, *childList_ = childList.p, *parent_ = parent.p, *flTree_ = flTree.p, *n0 = treeNode0.p;
	parallel_for<P_len>(c, (size_t)N + 1, PV_LAMBDA(size_t x) {
		if (x == N) { len_[nItems] = 0; return; }
		CountSink a; pw.prefix((u32)x, a); len_[2 * x + childOff_[x]] = a.n;
		CountSink b; pw.suffix((u32)x, b); len_[2 * x + childOff_[x + 1] + 1] = b.n;
	});
	parallel_for<P_len>(c, nFl, PV_LAMBDA(size_t k) {
		u32 f = childList_[k]; u32 x = n0[flTree_[f]] + parent_[f];
		CountSink s; pw.child(x, (u32)k, s); len_[2 * (size_t)x + 1 + k] = s.n;
	});
	exclusive_scan(c, len.p, off.p, nItems + 1, OpAddU64(), (u64)0);
	u64 total; c.d2h(&total, off.p + nItems, 8);
	Buf<u8> text(c, total + 8);
	u8 *text_ = text.p; const u64 *off_ = off.p;
	parallel_for<P_write>(c, N, PV_LAMBDA(size_t x) {
		WriteSink a(text_ + off_[2 * x + childOff_[x]]); pw.prefix((u32)x, a); a.finish();
		WriteSink b(text_ + off_[2 * x + childOff_[x + 1] + 1]); pw.suffix((u32)x, b); b.finish();
	});
	parallel_for<P_write>(c, nFl, PV_LAMBDA(size_t k) {
		u32 f = childList_[k]; u32 x = n0[flTree_[f]] + parent_[f];
		WriteSink s(text_ + off_[2 * (size_t)x + 1 + k]); pw.child(x, (u32)k, s); s.finish();
	});
	// file boundaries: the prefix item of every tree's root
	Buf<u64> tOff(c, (size_t)T + 1);
	{ u64 *tOff_ = tOff.p; parallel_for<P_node>(c, (size_t)T + 1, PV_LAMBDA(size_t t) { tOff_[t] = t < T ? off_[2 * (size_t)n0[t] + childOff_[n0[t]]] : total; }); }
	out.text.resize(total);
	c.d2h_async(out.text.data(), text.p, total);
	c.d2h_async(out.treeOff.data(), tOff.p, ((size_t)T + 1) * 8);
	c.sync();
}

} // namespace pv
