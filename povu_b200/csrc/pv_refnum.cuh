// povu_b200 -- the reference's INTERNAL NUMBERING of the cycle-equivalence pass, derived in parallel from the arrays decompose
// already holds (exact DFS tree in preorder, back edges, simplifying edges).  Two users: (1) the debug surfaces of the reference,
// `--hairpins` (stderr "Boundary: b1 b2", /root/reference/src/povu/algorithms/flubbles.cpp:503-719) and the per-component
// "povu.flubble-debug.frame.v1" frames of `--structure-export` (flubbles.cpp:108-213); (2) decompose itself, on the rare trees
// where the reference's class names are not cycle equivalence (refnum_quirk_flag below): the classes then come from here.
// On every named benchmark configuration the hot path never runs this.
//
// What the reference computes sequentially, vertex by vertex in descending DFS index, and what stands in for it here:
//   * bracket list of v (std::list, children spliced on top in ascending index order, own pushes on top of those)
//       = the brackets whose source lies in subtree(v) and whose target is a proper ancestor of v, in ONE global order: the
//         POSTORDER of their source (own pushes in push order: real back edges by adjacency slot, then the capping edge, then the
//         simplifying edge).  subtree(v) is a contiguous range of that order, so
//           top(v)  = the LAST bracket of the range whose target index is < v   (a "previous smaller value" query),
//           size(v) = a difference of prefix sums over (+1 at the source, -1 at the target).
//   * capping back edges (flubbles.cpp:556-573,611-617): v -> hi_2 when hi_2 < hi_0, with the reference's own (quirky) choice of
//     hi_2 = hi of the first child that is not hi_child and reaches above v; hi(x) = range minimum of the targets in subtree(x).
//   * new_class() numbering (flubbles.cpp:591-603,667-684): a tree edge opens a class iff (top, size) differs from the previous
//     vertex at which the same bracket was on top -- one stable sort of the vertices by their top bracket puts those neighbours
//     next to each other; a back edge gets its own class at its target unless it was on top of a list of size one before; the ids
//     are then a suffix sum of the per-vertex counts (the reference walks the vertices downwards).
//   * edge ids (spanning_tree.cpp:784-805): tree-edge index + number of back edges created before the vertex pair was discovered;
//     a back edge seen at adjacency slot s of port p is created right before p's next child behind s is discovered (or, without
//     one, before the first vertex after subtree(p)).
//   * hairpin boundaries: a run opens at a vertex whose list came up empty (= gets the simplifying edge), runs down the indices
//     to the next leaf / the root, and its second end is the last vertex of the run whose top bracket is a simplifying edge.
#pragma once
#include "pv_prims.cuh"

namespace pv {

PV_KERNEL(R_count, 48);
PV_KERNEL(R_cap, 40);
PV_KERNEL(R_fill, 48);
PV_KERNEL(R_top, 32);
PV_KERNEL(R_class, 24);
PV_KERNEL(R_out, 32);

struct OpMaxU32 { PV_HD u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };

// real back edges that leave port p, in the order the reference creates them while scanning p's side (ascending adjacency slot;
// a side without links: the one back edge to the dummy root): fn(target preorder, label id, slot or NONE).  Shared by the label
// pass of decompose (K_backedge) and the numbering pass, so both see the same edge set.
template <class Fn>
PV_DEV void for_each_backedge(u32 p, const u32 *adjOff_, const u32 *adjNbr_, const u32 *pre_, const u32 *parent_, const u32 *entrySide_, u32 rootPre, u64 M, Fn fn) {
	const u32 v = p >> 1, myPre = pre_[p];
	const u32 s0 = adjOff_[p], s1 = adjOff_[p + 1];
	const bool isExit = (p & 1u) != entrySide_[v];
	bool loopDone = false;
	if (s0 == s1) {   // tip side: back edge to the dummy root unless the tree already links them (spanning_tree.cpp:431-438)
		if (parent_[p] != NONE) fn(rootPre, (u64)M + p, NONE);
		return;
	}
	for (u32 s = s0; s < s1; s++) {
		u32 q = adjNbr_[s];
		if (q == (p ^ 1u)) {   // self-loop: one extra back edge per segment, from the child side (spanning_tree.cpp:387-395)
			if (isExit && !loopDone) { loopDone = true; fn(pre_[q], (u64)s, s); }
			continue;
		}
		u32 qPre = pre_[q];
		if (qPre > myPre) continue;            // recorded from the deeper endpoint
		if (parent_[p] == q) continue;          // the tree edge itself, or parallel to it
		bool dup = false;
		for (u32 t = s0; t < s; t++) if (adjNbr_[t] == q) { dup = true; break; }   // parallel edges collapse (spanning_tree.cpp:379-386)
		if (!dup) fn(qPre, (u64)s, s);
	}
}

// what decompose_device lends to the numbering pass (device pointers; g = global preorder index, ports p = 2 * vertex + side)
struct RefNumIn {
	u32 V, P, N, nComp, nS; u64 M;
	const u32 *adjOff, *adjNbr, *treeSlot, *parent, *entrySide, *pre, *sz, *depth, *pre2node;
	const u32 *compIdxOfRoot, *gcomp, *cNodeBase, *cNv, *cHasTips, *cSegBase;
	const u32 *bridge, *bsum;      // bridge flags + their exclusive prefix sums (the simplifying edge sits where exactly one lies in the subtree)
	const i32 *PC;                 // prefix sums of (+1 source, -1 target) over real + simplifying edges
	RmqView hiAll;                 // range minimum over the per-vertex lowest target (real + simplifying edges)
	const u32 *minT;               // that array
	const u32 *stV, *segComp;      // class stack: vertex per position, component per position
};
struct RefNumOut {
	std::vector<u32> edge_idx, edge_id, class_id;      // per class-stack position
	std::vector<u32> hp_comp, hp_b1, hp_b2;            // hairpin boundaries in the reference's print order (vertex indices; NONE = unset)
	u32 errors = 0;                                    // vertices without a top bracket (never on a consistent tree)
	bool valid = false;
	void reset() { edge_idx.clear(); edge_id.clear(); class_id.clear(); hp_comp.clear(); hp_b1.clear(); hp_b2.clear(); errors = 0; valid = false; }
};

// Can the reference's class naming differ from true cycle equivalence on this tree?  Its capping back edge goes to hi_2 = hi of
// the FIRST child (ascending index) that is not hi_child and reaches above v (flubbles.cpp:564-573), where cycle equivalence
// needs the LOWEST hi among those children.  Where the two differ the bracket ends below brackets of a sibling it should cover,
// and two tree edges with different bracket sets can share (top bracket, list size): the reference gives them one class and
// reports a flubble that cycle equivalence does not have.  flag[0] += 1 for every vertex where the two choices differ (three or
// more children must reach above it for that; the usual multi-allelic bubble has them all reach the same vertex and is not counted).
PV_KERNEL(R_quirk, 24);
inline void refnum_quirk_flag(Ctx &c, const RefNumIn &t, u32 *flag) {
	const u32 P = t.P, N = t.N;
	const u32 *adjOff_ = t.adjOff, *treeSlot_ = t.treeSlot, *entrySide_ = t.entrySide, *pre_ = t.pre, *sz_ = t.sz, *pre2node_ = t.pre2node;
	const RmqView hv = t.hiAll;
	parallel_for<R_quirk>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg, node = pre2node_[g];
		if (node >= P || sz_[node] < 6) return;      // three children need three vertex pairs below
		const bool isEntry = (node & 1u) == entrySide_[node >> 1];
		u32 nGray = 0;
		for (u32 s = adjOff_[node], e = adjOff_[node + 1]; s < e; s++) if (treeSlot_[s] != NONE) nGray++;
		if (nGray + (isEntry ? 1u : 0u) < 3) return;
		auto hiOf = [&](u32 ch) -> u32 { u32 gc = pre_[ch]; return hv.query(gc, gc + sz_[ch] - 1); };
		// hi_child = the first child with the lowest hi (flubbles.cpp:546-562)
		u32 hi1 = NONE, hiChild = NONE;
		if (isEntry) { hi1 = hiOf(node ^ 1u); hiChild = node ^ 1u; }
		for (u32 s = adjOff_[node], e = adjOff_[node + 1]; s < e; s++) { u32 ch = treeSlot_[s]; if (ch == NONE) continue; u32 h = hiOf(ch); if (hiChild == NONE || h < hi1) { hi1 = h; hiChild = ch; } }
		// among the other children that reach above g: the first one's hi (the reference's hi_2) against the lowest
		u32 first = NONE, lowest = NONE;
		auto other = [&](u32 ch) { if (ch == hiChild) return; u32 h = hiOf(ch); if (h >= g) return; if (first == NONE) first = h; if (h < lowest) lowest = h; };
		if (isEntry) other(node ^ 1u);
		for (u32 s = adjOff_[node], e = adjOff_[node + 1]; s < e; s++) { u32 ch = treeSlot_[s]; if (ch != NONE) other(ch); }
		if (first != lowest) atomic_add(flag, 1u);
	});
}

// out (optional): the host-side numbering + hairpin boundaries of the two debug surfaces.  stackClass (optional): receives, on the
// device, the reference's class of every class-stack position as a label that is unique per class inside the whole forest (what
// decompose groups by when the quirk above can bite).
inline void refnum_device(Ctx &c, const RefNumIn &t, RefNumOut *outp, Buf<u32> *stackClass = nullptr) {
	RefNumOut scratch; RefNumOut &out = outp ? *outp : scratch;
	out.reset();
	const u32 P = t.P, N = t.N, nS = t.nS; const u64 M = t.M;
	const u32 *adjOff_ = t.adjOff, *adjNbr_ = t.adjNbr, *treeSlot_ = t.treeSlot, *parent_ = t.parent, *entrySide_ = t.entrySide, *pre_ = t.pre, *sz_ = t.sz, *depth_ = t.depth, *pre2node_ = t.pre2node;
	const u32 *compIdxOfRoot_ = t.compIdxOfRoot, *gcomp_ = t.gcomp, *cNodeBase_ = t.cNodeBase, *cNv_ = t.cNv, *cHasTips_ = t.cHasTips;
	const u32 *bridge_ = t.bridge, *bsum_ = t.bsum, *minT_ = t.minT; const i32 *PC_ = t.PC;
	const RmqView hv = t.hiAll;
	enum : u32 { KIND_REAL = 0, KIND_CAP = 1, KIND_SIMP = 2 };

	// ---- per vertex: component, subtree size, depth, "is a vertex of the reference's tree", simplifying flag
	Buf<u32> gK(c, N), gSz(c, N), gDepth(c, N), gFlags(c, N);   // flags: 1 = reference vertex, 2 = reference root, 4 = simplifying edge starts here
	u32 *gK_ = gK.p, *gSz_ = gSz.p, *gDepth_ = gDepth.p, *gFlags_ = gFlags.p;
	parallel_for<R_count>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg, node = pre2node_[g];
		u32 k, sz, d, fl = 0;
		if (node >= P) { k = node - P; sz = 2 * cNv_[k] + 1; d = 0; if (cHasTips_[k]) fl = 1 | 2; }      // the dummy root exists only when the component has tips
		else {
			k = compIdxOfRoot_[gcomp_[node >> 1]]; sz = sz_[node]; d = depth_[node]; fl = 1;
			if (!cHasTips_[k] && g == cNodeBase_[k] + 1) fl |= 2;
			if (bridge_[g] && bsum_[g + sz] - bsum_[g] == 1) fl |= 4;
		}
		gK_[g] = k; gSz_[g] = sz; gDepth_[g] = d; gFlags_[g] = fl;
	});

	// ---- own real back edges per vertex, creation-time histogram (ND = the vertex discovered next after the edge was created)
	Buf<u32> nReal(c, N), ndCnt(c, (size_t)N + 1), ndEx(c, (size_t)N + 1);
	nReal.zero(); ndCnt.zero();
	u32 *nReal_ = nReal.p, *ndCnt_ = ndCnt.p, *ndEx_ = ndEx.p;
	parallel_for<R_count>(c, P, PV_LAMBDA(size_t pp) {
		const u32 p = (u32)pp, g = pre_[p], k = gK_[g];
		const u32 compEnd = cNodeBase_[k] + 2 * cNv_[k] + 1;
		u32 n = 0;
		for_each_backedge(p, adjOff_, adjNbr_, pre_, parent_, entrySide_, cNodeBase_[k], M, [&](u32, u64, u32 s) {
			n++;
			u32 nd = g + sz_[p];
			if (s != NONE) for (u32 x = s + 1, e = adjOff_[p + 1]; x < e; x++) { u32 ch = treeSlot_[x]; if (ch != NONE) { nd = pre_[ch]; break; } }
			if (nd < compEnd) atomic_add(ndCnt_ + nd, 1u);
		});
		nReal_[g] = n;
	});
	exclusive_scan(c, ndCnt.p, ndEx.p, (size_t)N + 1, OpAddU32(), 0u);

	// ---- capping back edges (flubbles.cpp:517-573,611-617)
	Buf<u32> capTgt(c, N);
	Buf<i32> capD(c, (size_t)N + 1), PCap(c, (size_t)N + 1);
	capD.zero();
	u32 *capTgt_ = capTgt.p; i32 *capD_ = capD.p, *PCap_ = PCap.p;
	parallel_for<R_cap>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg, node = pre2node_[g];
		capTgt_[g] = NONE;
		if (node >= P || gSz_[g] < 4) return;      // two children need at least two vertex pairs below
		const bool isEntry = (node & 1u) == entrySide_[node >> 1];
		// children in ascending index order: the black child first (entry side), then the gray children by adjacency slot
		auto hiOf = [&](u32 ch) -> u32 { u32 gc = pre_[ch]; return hv.query(gc, gc + sz_[ch] - 1); };
		u32 hi1 = NONE, hiChild = NONE, nCh = 0;
		if (isEntry) { u32 h = hiOf(node ^ 1u); hi1 = h; hiChild = node ^ 1u; nCh++; }
		for (u32 s = adjOff_[node], e = adjOff_[node + 1]; s < e; s++) {
			u32 ch = treeSlot_[s];
			if (ch == NONE) continue;
			u32 h = hiOf(ch); nCh++;
			if (hiChild == NONE || h < hi1) { hi1 = h; hiChild = ch; }
		}
		if (nCh < 2) return;
		u32 hi2 = NONE;
		if (isEntry && (node ^ 1u) != hiChild) { u32 h = hiOf(node ^ 1u); if (h < g) hi2 = h; }
		if (hi2 == NONE) for (u32 s = adjOff_[node], e = adjOff_[node + 1]; s < e; s++) {
			u32 ch = treeSlot_[s];
			if (ch == NONE || ch == hiChild) continue;
			u32 h = hiOf(ch);
			if (h < g) { hi2 = h; break; }
		}
		const u32 hi0 = (gFlags_[g] & 4u) ? NONE : minT_[g];      // own real targets only (a vertex with the simplifying edge has none)
		if (hi2 != NONE && hi2 < hi0) { capTgt_[g] = hi2; atomic_add((u32 *)capD_ + g, 1u); atomic_add((u32 *)capD_ + hi2, 0xFFFFFFFFu); }
	});
	exclusive_scan(c, capD.p, PCap.p, (size_t)N + 1, OpAddI32(), 0);

	// ---- all brackets in list order (postorder of the source; own pushes in push order)
	Buf<u32> cntPost(c, (size_t)N + 1), bbase(c, (size_t)N + 1);
	cntPost.zero();
	u32 *cntPost_ = cntPost.p, *bbase_ = bbase.p;
	parallel_for<R_fill>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg; const u32 fl = gFlags_[g];
		if (!(fl & 1u)) return;
		cntPost_[g - gDepth_[g] + gSz_[g] - 1] = nReal_[g] + (capTgt_[g] != NONE ? 1u : 0u) + ((fl & 4u) ? 1u : 0u);
	});
	exclusive_scan(c, cntPost.p, bbase.p, (size_t)N + 1, OpAddU32(), 0u);
	const u32 NB = bbase.read(N);
	if (NB == 0) {
		out.valid = true; out.edge_idx.assign(nS, 0); out.edge_id.assign(nS, 0); out.class_id.assign(nS, 0);
		if (stackClass) { *stackClass = Buf<u32>(c, nS ? nS : 1); u32 *sc = stackClass->p; parallel_for<R_out>(c, nS, PV_LAMBDA(size_t i) { sc[i] = (u32)i; }); }
		return;
	}
	Buf<u32> bT(c, NB), bKind(c, NB);
	u32 *bT_ = bT.p, *bKind_ = bKind.p;
	parallel_for<R_fill>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg; const u32 fl = gFlags_[g];
		if (!(fl & 1u)) return;
		u32 w = bbase_[g - gDepth_[g] + gSz_[g] - 1];
		const u32 node = pre2node_[g], k = gK_[g];
		if (node < P) for_each_backedge(node, adjOff_, adjNbr_, pre_, parent_, entrySide_, cNodeBase_[k], M, [&](u32 tgt, u64, u32) { bT_[w] = tgt; bKind_[w] = KIND_REAL; w++; });
		if (capTgt_[g] != NONE) { bT_[w] = capTgt_[g]; bKind_[w] = KIND_CAP; w++; }
		if (fl & 4u) { bT_[w] = cHasTips_[k] ? cNodeBase_[k] : cNodeBase_[k] + 1; bKind_[w] = KIND_SIMP; w++; }
	});

	// ---- top bracket and list size of every vertex below the root
	Rmq rmqB;
	rmq_build(c, rmqB, bT.p, NB);
	Buf<u32> topOf(c, N), lsize(c, N), err(c, 1);
	err.zero();
	u32 *topOf_ = topOf.p, *lsize_ = lsize.p, *err_ = err.p;
	{
		RmqView rb = rmqB.view();
		parallel_for<R_top>(c, N, PV_LAMBDA(size_t gg) {
			const u32 g = (u32)gg; const u32 fl = gFlags_[g];
			topOf_[g] = NB; lsize_[g] = 0;
			if (!(fl & 1u) || (fl & 2u)) return;
			const u32 sz = gSz_[g], gpost = g - gDepth_[g] + sz - 1;
			const u32 lo = bbase_[gpost + 1 - sz], hi = bbase_[gpost + 1];
			u32 j = hi > lo ? rb.prev_smaller(hi - 1, g) : NONE;
			if (j == NONE || j < lo) { atomic_add(err_, 1u); return; }
			topOf_[g] = j;
			lsize_[g] = (u32)((PC_[g + sz] - PC_[g]) + (PCap_[g + sz] - PCap_[g]));
		});
	}

	// ---- class numbering: vertices grouped by top bracket, inside a group in processing (descending index) order
	Buf<u32> skey(c, N), sval(c, N);
	{
		u32 *k_ = skey.p, *v_ = sval.p;
		parallel_for<R_class>(c, N, PV_LAMBDA(size_t i) { u32 g = N - 1 - (u32)i; k_[i] = topOf_[g]; v_[i] = g; });
	}
	radix_sort_pairs(c, skey, sval, N, bits_for(NB));
	const u32 *skey_ = skey.p, *sval_ = sval.p;
	Buf<u32> lead(c, N), leadS(c, N), newTree(c, N), defined(c, NB), cntBE(c, N);
	newTree.zero(); defined.zero(); cntBE.zero();
	u32 *lead_ = lead.p, *leadS_ = leadS.p, *newTree_ = newTree.p, *defined_ = defined.p, *cntBE_ = cntBE.p;
	parallel_for<R_class>(c, N, PV_LAMBDA(size_t i) {
		const u32 k = skey_[i];
		lead_[i] = 0;
		if (k >= NB) return;
		const u32 g = sval_[i];
		const bool nw = i == 0 || skey_[i - 1] != k || lsize_[sval_[i - 1]] != lsize_[g];
		if (nw) { lead_[i] = (u32)i + 1; newTree_[g] = 1; }
		if (lsize_[g] == 1) defined_[k] = 1;      // the bracket shares the class of this tree edge (flubbles.cpp:678-683)
	});
	inclusive_scan(c, lead.p, leadS.p, N, OpMaxU32(), 0u);
	parallel_for<R_class>(c, NB, PV_LAMBDA(size_t j) { if (bKind_[j] != KIND_CAP && !defined_[j]) atomic_add(cntBE_ + bT_[j], 1u); });
	Buf<u32> tot(c, N), totEx(c, N), classNew(c, N), classOf(c, N);
	u32 *tot_ = tot.p, *totEx_ = totEx.p, *classNew_ = classNew.p, *classOf_ = classOf.p;
	parallel_for<R_class>(c, N, PV_LAMBDA(size_t i) { u32 g = N - 1 - (u32)i; tot_[i] = cntBE_[g] + newTree_[g]; });
	exclusive_scan(c, tot.p, totEx.p, N, OpAddU32(), 0u);
	parallel_for<R_class>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg, k = gK_[g];
		const u32 last = cNodeBase_[k] + 2 * cNv_[k];      // highest index of the component
		classNew_[g] = totEx_[N - 1 - g] - totEx_[N - 1 - last] + cntBE_[g];
	});
	parallel_for<R_class>(c, N, PV_LAMBDA(size_t i) { if (skey_[i] < NB) classOf_[sval_[i]] = classNew_[sval_[leadS_[i] - 1]]; });

	const u32 *stV_ = t.stV;
	if (stackClass) {      // class label = position of the class's opening vertex in the sorted order (unique across components)
		*stackClass = Buf<u32>(c, nS ? nS : 1);
		Buf<u32> labOf(c, N);
		u32 *labOf_ = labOf.p, *sc = stackClass->p;
		parallel_for<R_class>(c, N, PV_LAMBDA(size_t i) { if (skey_[i] < NB) labOf_[sval_[i]] = leadS_[i]; });
		parallel_for<R_out>(c, nS, PV_LAMBDA(size_t i) { const u32 v = stV_[i], cx = (2 * v + entrySide_[v]) ^ 1u; sc[i] = labOf_[pre_[cx]]; });
	}
	if (!outp) { if (err.read(0)) throw std::runtime_error("decompose: inconsistent bracket lists (internal error)"); return; }
	// ---- per class-stack position: tree edge index, edge id, class id
	Buf<u32> oIdx(c, nS ? nS : 1), oId(c, nS ? nS : 1), oCls(c, nS ? nS : 1);
	u32 *oIdx_ = oIdx.p, *oId_ = oId.p, *oCls_ = oCls.p;
	parallel_for<R_out>(c, nS, PV_LAMBDA(size_t i) {
		const u32 v = stV_[i], cx = (2 * v + entrySide_[v]) ^ 1u, gc = pre_[cx], k = gK_[gc];
		const u32 base = cNodeBase_[k];
		const u32 child = gc - base - (cHasTips_[k] ? 0u : 1u);
		oIdx_[i] = child - 1;
		oId_[i] = child - 1 + (ndEx_[gc] - ndEx_[base]);      // back edges created before the pair (gc - 1, gc) was discovered: ND <= gc - 1
		oCls_[i] = classOf_[gc];
	});

	// ---- hairpin boundaries
	Buf<u32> closeP(c, N), closeS(c, N), simpTop(c, N), startF(c, N);
	u32 *closeP_ = closeP.p, *closeS_ = closeS.p, *simpTop_ = simpTop.p, *startF_ = startF.p;
	parallel_for<R_out>(c, N, PV_LAMBDA(size_t gg) {
		const u32 g = (u32)gg, fl = gFlags_[g];
		const bool cl = (fl & 1u) && ((fl & 2u) || gSz_[g] == 1);
		closeP_[g] = cl ? g + 1 : 0u;
		const u32 tp = topOf_[g];
		simpTop_[g] = (tp < NB && bKind_[tp] == KIND_SIMP) ? g : NONE;
		startF_[g] = (fl & 4u) ? 1u : 0u;
	});
	inclusive_scan(c, closeP.p, closeS.p, N, OpMaxU32(), 0u);
	Rmq rmqS;
	rmq_build(c, rmqS, simpTop.p, N);
	Buf<u32> starts;
	const u32 nH = compact_flags(c, startF.p, N, starts);
	Buf<u32> hK(c, nH ? nH : 1), hB1(c, nH ? nH : 1), hB2(c, nH ? nH : 1);
	{
		u32 *hK_ = hK.p, *hB1_ = hB1.p, *hB2_ = hB2.p; const u32 *starts_ = starts.p;
		RmqView rs = rmqS.view();
		parallel_for<R_out>(c, nH, PV_LAMBDA(size_t i) {
			const u32 g0 = starts_[i];
			hK_[i] = gK_[g0]; hB1_[i] = pre2node_[g0] >> 1;
			const u32 cpos = closeS_[g0 - 1];      // (the root closes, and lies below every start)
			u32 b2 = NONE;
			if (cpos != 0 && cpos <= g0 - 1) { u32 m = rs.query(cpos, g0 - 1); if (m != NONE) b2 = pre2node_[m] >> 1; }      // vertices (close, g0): cpos = close + 1
			hB2_[i] = b2;
		});
	}

	out.edge_idx = oIdx.to_host(nS); out.edge_id = oId.to_host(nS); out.class_id = oCls.to_host(nS);
	out.errors = err.read(0);
	if (nH) {
		std::vector<u32> k = hK.to_host(nH), b1 = hB1.to_host(nH), b2 = hB2.to_host(nH);
		// the reference pushes a boundary when its run closes, walking the indices downwards: per component, descending start
		size_t i = 0;
		while (i < nH) {
			size_t j = i;
			while (j < nH && k[j] == k[i]) j++;
			for (size_t x = j; x-- > i;) { out.hp_comp.push_back(k[x]); out.hp_b1.push_back(b1[x]); out.hp_b2.push_back(b2[x]); }
			i = j;
		}
	}
	out.valid = true;
}

} // namespace pv
