// povu_b200 -- call: (graph + haplotype paths + PVST flubbles) -> binary VCF records, every step a
// data-parallel kernel over flat HBM arrays.  Replaces, with identical results, the reference's
//   ic::color_pvst              /root/reference/src/ita/variation/color.cpp:15-30,104-163
//   ir::gen_rov / find_pvst_rovs /root/reference/src/ita/variation/rov.cpp:185-268
//   bfs_sort / comp_bfs_tree    /root/reference/src/ita/graph/graph.cpp:75-163,557-577
//   BfsTree::sort (Kahn)        /root/reference/include/ita/graph/bfs_tree.hpp:257-328
//   depth_matrix::fill          /root/reference/include/ita/genomics/allele.hpp:804-862
//   gen_race (lap 0)            /root/reference/src/ita/genomics/untangle.cpp:29-103
//   comp_exps + helpers         /root/reference/src/ita/variation/overlay.cpp:84-588
//   minimal_rov grouping        /root/reference/include/ita/genomics/allele.hpp:302-475
//   gen_exp_vcf_recs / gen_gt_data  /root/reference/src/ita/genomics/vcf.cpp:20-121,241-273
//
// Device data layout (DESIGN.md "call: HBM layout"):
//   steps[S]      packed path steps, paths concatenated in P-line order: (vertex idx << 1) | strand
//   stepOff[H+1]  first step of every path;  loci[S] 1-based start coordinate of every step
//   vsOff[V+1], vsStep[S], vsHap[S]   per-vertex list of the (global) steps that visit it, ascending; vsHap = (path << 1) | strand --
//                 since paths are concatenated in hap order this is also (hap, local step) order, i.e. the
//                 reference's vertex_to_step_matrix_ flattened (bidirected.cpp:347-355)
//   adjOff[2V+1], adjNbr[..]          per-port neighbour ports in ascending L-line order (bidirected.hpp:68-69)
//   per region r: verts[rovOff[r] .. +J)  sorted vertices;  depth matrix cells u16 [nh_r x J_r] at matBase[r]
//                 (rows = the paths of the region's component, ascending path id)
#pragma once
#include "pv_decompose.cuh"

namespace pv {

typedef uint16_t u16;

#ifndef PV_EMU
PV_DEV u64 atomic_cas64(u64 *p, u64 cmp, u64 v) { return (u64)atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)v); }
#else
inline u64 atomic_cas64(u64 *p, u64 cmp, u64 v) { u64 o = *p; if (o == cmp) *p = v; return o; }
#endif

PV_KERNEL(C_adj_fill, 16);
PV_KERNEL(C_loci_len, 12);
PV_KERNEL(C_loci, 20);
PV_KERNEL(C_vs_key, 12);
PV_KERNEL(C_vs_hap, 12);
PV_KERNEL(C_uf, 16);
PV_KERNEL(C_color, 32);
PV_KERNEL(C_child, 12);
PV_KERNEL(C_restrict, 24);
PV_KERNEL(C_bfs, 128);
PV_KERNEL(C_gather, 16);
PV_KERNEL(C_rovmeta, 24);
PV_KERNEL(C_hash, 16);
PV_KERNEL(C_region, 72);        // per region lane: its share of the step lists in, cells + laps + context counts out (C3 + C4 + C5a fused)
PV_KERNEL(C_matfill, 24);       // per (region, column): the vertex's step list in, one cell per visiting path out
PV_KERNEL(C_lap, 16);
PV_KERNEL(C_pair_count, 16);    // per (region, path): two matrix rows in, a count out
PV_KERNEL(C_pair_fill, 80);    // per candidate slot: two matrix rows + the steps of both slices in, 9 words out (the row search and the step-list searches are not compulsory traffic)
PV_KERNEL(C_group, 64);
PV_KERNEL(C_emit, 64);
PV_KERNEL(C_gt_ref, 16);
PV_KERNEL(C_gt_alt, 16);
PV_KERNEL(C_fill16, 2);
PV_KERNEL(C_pack, 24);
PV_KERNEL(C_keep, 12);
PV_KERNEL(C_sizes, 24);

PV_DEV u32 lb_u32(const u32 *a, u32 lo, u32 hi, u32 key) {   // first index in [lo, hi) with a[idx] >= key
	while (lo < hi) { u32 m = lo + ((hi - lo) >> 1); if (a[m] < key) lo = m + 1; else hi = m; }
	return lo;
}
PV_DEV u32 ub_u32(const u32 *a, u32 lo, u32 hi, u32 key) {   // first index in [lo, hi) with a[idx] > key
	while (lo < hi) { u32 m = lo + ((hi - lo) >> 1); if (a[m] <= key) lo = m + 1; else hi = m; }
	return lo;
}
PV_DEV u64 ub_u64(const u64 *a, u64 lo, u64 hi, u64 key) {
	while (lo < hi) { u64 m = lo + ((hi - lo) >> 1); if (a[m] <= key) lo = m + 1; else hi = m; }
	return lo;
}

// ---------------------------------------------------------------------------------------------
// device-resident graph + paths (built once, used by every call)
// ---------------------------------------------------------------------------------------------
struct AdjEntryPort {   // port of adjacency entry j (2e = first end of link e, 2e+1 = second end; P = dropped duplicate)
	const u32 *ep1, *ep2; u32 P;
	PV_DEV u32 operator()(size_t j) const { u32 a = ep1[j >> 1], b = ep2[j >> 1]; return (j & 1u) ? (a == b ? P : b) : a; }
};
struct StepVertex { const u32 *steps; PV_DEV u32 operator()(size_t s) const { return steps[s] >> 1; } };

struct DevGraph {
	u32 V = 0, H = 0, S = 0; u64 E = 0;
	Buf<u32> adjOff, adjNbr;
	Buf<u32> steps, loci, stepOff, segLen;
	Buf<u32> vsOff, vsStep, vsHap;
	Buf<u32> vcomp;
	// host copies of the small per-path tables
	std::vector<u32> h_stepOff;
	std::vector<u32> pathComp;        // dense component index per path (NONE for empty paths)
	std::vector<u32> pathLabel;       // component label (lowest vertex index of the component) per path (NONE for empty paths)
	std::vector<u32> cpOff, cpList;   // paths of every component, ascending
	std::vector<u32> hapLocal;        // index of the path inside its component's list
	Buf<u32> d_pathComp, d_cpOff, d_cpList, d_hapLocal;
};

struct GraphView {
	u32 V, H, S;
	const u32 *adjOff, *adjNbr, *steps, *loci, *stepOff, *vsOff, *vsStep, *vsHap, *pathComp, *cpOff, *cpList, *hapLocal;
	// first / last step in [lo, hi) that visits vertex v (NONE if none)
	PV_DEV u32 occ_first(u32 v, u32 lo, u32 hi) const {
		u32 a = vsOff[v], b = vsOff[v + 1];
		u32 k = lb_u32(vsStep, a, b, lo);
		return (k < b && vsStep[k] < hi) ? vsStep[k] : NONE;
	}
	PV_DEV u32 occ_last(u32 v, u32 lo, u32 hi) const {
		u32 a = vsOff[v], b = vsOff[v + 1];
		u32 k = lb_u32(vsStep, a, b, hi);
		return (k > a && vsStep[k - 1] >= lo) ? vsStep[k - 1] : NONE;
	}
	PV_DEV bool has_hap(u32 v, u32 h) const {   // any step of path h on vertex v
		return occ_first(v, stepOff[h], stepOff[h + 1]) != NONE;
	}
};

inline GraphView view(const DevGraph &g) {
	return GraphView{g.V, g.H, g.S, g.adjOff.p, g.adjNbr.p, g.steps.p, g.loci.p, g.stepOff.p, g.vsOff.p, g.vsStep.p, g.vsHap.p,
			 g.d_pathComp.p, g.d_cpOff.p, g.d_cpList.p, g.d_hapLocal.p};
}

// weakly connected components: label[v] = lowest vertex index of v's component (VG::componetize discovers
// components in exactly that order, bidirected.cpp:585-596)
inline void component_labels(Ctx &c, u32 V, u64 E, const u32 *ep1, const u32 *ep2, u32 *label) {
	parallel_for<C_uf>(c, V, PV_LAMBDA(size_t v) { label[v] = (u32)v; });
	parallel_for<C_uf>(c, E, PV_LAMBDA(size_t e) { uf_union(label, ep1[e] >> 1, ep2[e] >> 1); });
	// Flattening in place must not compress paths: a halving store of another thread (a stale ancestor) could land after
	// the owner's final store and leave a non-root label behind.  Read-only walks see either the old parent or the root.
	parallel_for<C_uf>(c, V, PV_LAMBDA(size_t v) {
		u32 x = (u32)v, p = ld_volatile(label + x);
		while (p != x) { x = p; p = ld_volatile(label + x); }
		label[v] = x;
	});
}

// d_ep1/d_ep2: E edge endpoint ports (device).  d_steps: S packed steps (device).  d_segLen: V (device).
// labels (optional): component labels already computed on this graph (decompose hands them over on the hot path).
inline void graph_build(Ctx &c, u32 V, u64 E, const u32 *ep1, const u32 *ep2, u32 H, const u32 *h_stepOff, const u32 *d_steps,
			const u32 *d_segLen, DevGraph &g, Buf<u32> *labels = nullptr) {
	Trace tr(c, "index");
	g.V = V; g.E = E; g.H = H; g.S = h_stepOff ? h_stepOff[H] : 0;
	const u32 P = 2 * V; const u64 M = 2 * E; const u32 S = g.S;
	if (M >= 0xFFFFFFF0ull) throw std::runtime_error("graph too large for 32-bit adjacency slots");
	// Two groupings (adjacency by port, steps by vertex) and the component label of every path are all queued before the ONE
	// host wait of the index build: their "has large groups" flags and the path labels arrive together.
	// ---- adjacency in whole-graph edge order (what `call` iterates: bidirected.hpp:68-69)
	g.adjOff = Buf<u32>(c, (size_t)P + 2); g.adjNbr = Buf<u32>(c, M ? M : 1);
	// entries 2e / 2e+1 = the two ends of link e, grouped by port in link order (group_indices: ports have a handful of
	// entries; hubs take the restricted stable sort).  Bucket P collects the second end of a v+ -> v+ style loop, which
	// is one std::set entry in the reference.
	Buf<u32> alist(c, M ? M : 1);
	const AdjEntryPort AEP{ep1, ep2, P};
	GroupState gsAdj, gsVs;
	MultiRead mr(c);
	group_begin(c, AEP, NoOrd(), M, P + 1, g.adjOff.p, alist.p, 48u, gsAdj, &mr);
	// ---- paths
	g.h_stepOff.assign(h_stepOff, h_stepOff + H + 1);
	g.stepOff = Buf<u32>(c, (size_t)H + 1); c.h2d(g.stepOff.p, h_stepOff, ((size_t)H + 1) * 4);
	g.steps = Buf<u32>(c, S ? S : 1); c.d2d(g.steps.p, d_steps, (size_t)S * 4);
	g.segLen = Buf<u32>(c, V ? V : 1); c.d2d(g.segLen.p, d_segLen, (size_t)V * 4);
	g.loci = Buf<u32>(c, S ? S : 1);
	g.vsOff = Buf<u32>(c, (size_t)V + 1); g.vsHap = Buf<u32>(c, S ? S : 1);
	const u32 *steps_ = g.steps.p, *stepOff_ = g.stepOff.p, *segLen_ = g.segLen.p;
	{
		// loci[k] = 1-based start coordinate of step k (liteseq ref_walk::loci).  One u32 scan over all
		// walks: the coordinate is a difference of two prefixes, so 32-bit wrap-around cancels (loci itself is 32 bits wide).
		Buf<u32> len(c, (size_t)S + 1), pre(c, (size_t)S + 1);
		u32 *len_ = len.p, *pre_ = pre.p; u32 *loci_ = g.loci.p;
		parallel_for<C_loci_len>(c, (size_t)S + 1, PV_LAMBDA(size_t s) { len_[s] = s < S ? segLen_[steps_[s] >> 1] : 0u; });
		exclusive_scan(c, len.p, pre.p, (size_t)S + 1, OpAddU32(), 0u);
		parallel_for<C_loci>(c, S, PV_LAMBDA(size_t s) {
			u32 h = ub_u32(stepOff_, 0, H + 1, (u32)s) - 1;
			loci_[s] = 1u + pre_[s] - pre_[stepOff_[h]];
		});
	}
	// vertex -> steps index = the step indices grouped by vertex, ascending inside a vertex (group_indices, pv_prims.cuh).
	// DENSE walks (dozens to hundreds of haplotypes: every vertex owns a long list) skip the small-group machinery altogether: one stable
	// radix sort of (packed step, step index) on the VERTEX bits of the key -- the strand bit rides along below them, so the
	// (path, strand) word of every list entry is later formed from the sorted key instead of a random read of steps[].
	u32 vs_small = 48;
	if (const char *e = getenv("POVU_B200_VS_SMALL")) vs_small = (u32)strtoul(e, nullptr, 10);   // tests: force the hub path
	u64 dense_min = 20;      // measured: at 26 steps per vertex (cfg4, 90 paths) the sort path builds the index in 5.6 ms against 9.1 ms; cfg2 (12) and cfg3 (10) keep the small-group path
	if (const char *e = getenv("POVU_B200_DENSE_INDEX")) dense_min = strtoull(e, nullptr, 10);       // tests: 0 forces the dense path
	const bool dense = S > 0 && (u64)S >= dense_min * (u64)V;
	const StepVertex SV{steps_};
	Buf<u32> dKey;      // dense: the sorted packed steps
	if (!dense) { g.vsStep = Buf<u32>(c, S ? S : 1); group_begin(c, SV, NoOrd(), S, V, g.vsOff.p, g.vsStep.p, vs_small, gsVs, &mr); }
	else {
		Buf<u32> cnt(c, (size_t)V + 1), vv(c, S);
		cnt.zero();
		dKey = Buf<u32>(c, S);
		u32 *cnt_ = cnt.p, *kk_ = dKey.p, *vv_ = vv.p;
		parallel_for<K_grp_count>(c, S, PV_LAMBDA(size_t s) { const u32 x = steps_[s]; atomic_add(cnt_ + (x >> 1), 1u); kk_[s] = x; vv_[s] = (u32)s; });
		exclusive_scan(c, cnt.p, g.vsOff.p, (size_t)V + 1, OpAddU32(), 0u);
		radix_sort_pairs(c, dKey, vv, S, bits_for(V), 1);
		g.vsStep = std::move(vv);
	}
	// ---- weakly connected components -> the component label of every path
	if (labels && labels->p && labels->n >= V && V) g.vcomp = std::move(*labels);
	else { g.vcomp = Buf<u32>(c, V ? V : 1); component_labels(c, V, E, ep1, ep2, g.vcomp.p); }
	Buf<u32> pr(c, H ? H : 1);
	std::vector<u32> roots(H ? H : 1, NONE);
	{
		u32 *uf_ = g.vcomp.p; u32 *pr_ = pr.p;
		parallel_for<C_uf>(c, H, PV_LAMBDA(size_t h) { pr_[h] = stepOff_[h + 1] > stepOff_[h] ? uf_[steps_[stepOff_[h]] >> 1] : NONE; });
		if ((size_t)H * 4 + 64 < ((size_t)1 << 20)) mr.add(roots.data(), pr.p, (size_t)H * 4);
	}
	mr.wait();
	if ((size_t)H * 4 + 64 >= ((size_t)1 << 20)) roots = pr.to_host();
	roots.resize(H);
	tr.mark("groupings queued");
	group_finish(c, AEP, NoOrd(), gsAdj);
	{
		u32 *adjNbr_ = g.adjNbr.p; const u32 *list_ = alist.p;
		parallel_for<C_adj_fill>(c, M, PV_LAMBDA(size_t s) {
			u32 j = list_[s]; u32 a = ep1[j >> 1], b = ep2[j >> 1];
			bool self = (a >> 1) == (b >> 1);
			// Edge::get_other_vtx(v, side) answers the complement side for any self-loop (bidirected.cpp:66-77)
			adjNbr_[s] = (j & 1u) ? (self ? (b ^ 1u) : a) : (self ? (a ^ 1u) : b);
		});
	}
	tr.mark("adjacency");
	if (!dense) group_finish(c, SV, NoOrd(), gsVs);
	{
		// (path << 1) | strand next to every entry: the depth matrix needs the strand per visit, and reading it here (one pass, the
		// step index at hand) is cheaper than a random access per matrix cell later
		u32 *vsStep_ = g.vsStep.p, *vsHap_ = g.vsHap.p; const u32 *dk = dKey.p;
		if (dense) parallel_for<C_vs_hap>(c, S, PV_LAMBDA(size_t i) { vsHap_[i] = ((ub_u32(stepOff_, 0, H + 1, vsStep_[i]) - 1) << 1) | (dk[i] & 1u); });
		else parallel_for<C_vs_hap>(c, S, PV_LAMBDA(size_t i) { u32 st = vsStep_[i]; vsHap_[i] = ((ub_u32(stepOff_, 0, H + 1, st) - 1) << 1) | (steps_[st] & 1u); });
	}
	tr.mark("loci + vertex->steps");
	{
		g.pathLabel = roots;
		std::vector<u32> uniq;
		for (u32 r : roots) if (r != NONE) uniq.push_back(r);
		std::sort(uniq.begin(), uniq.end()); uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
		g.pathComp.assign(H, NONE); g.hapLocal.assign(H, 0);
		g.cpOff.assign(uniq.size() + 1, 0);
		for (u32 h = 0; h < H; h++) if (roots[h] != NONE) { u32 k = (u32)(std::lower_bound(uniq.begin(), uniq.end(), roots[h]) - uniq.begin()); g.pathComp[h] = k; g.cpOff[k + 1]++; }
		for (size_t k = 0; k < uniq.size(); k++) g.cpOff[k + 1] += g.cpOff[k];
		g.cpList.assign(g.cpOff.back(), 0);
		std::vector<u32> fill(g.cpOff.begin(), g.cpOff.end() - 1);
		for (u32 h = 0; h < H; h++) if (g.pathComp[h] != NONE) { u32 k = g.pathComp[h]; g.hapLocal[h] = fill[k] - g.cpOff[k]; g.cpList[fill[k]++] = h; }
		auto up = [&](Buf<u32> &b, const std::vector<u32> &h) { b = Buf<u32>(c, h.size() ? h.size() : 1); c.h2d(b.p, h.data(), h.size() * 4); };
		up(g.d_pathComp, g.pathComp); up(g.d_cpOff, g.cpOff); up(g.d_cpList, g.cpList); up(g.d_hapLocal, g.hapLocal);   // (sources are members of g: they outlive the copies)
	}
	tr.mark("components");
}

// ---------------------------------------------------------------------------------------------
// region vertex order: BFS tree + Kahn-style sort (graph.cpp:75-163, bfs_tree.hpp:257-328), one thread
// per region over caller-provided scratch.  Returns the number of sorted vertices, 0 if the sort fails
// (cycle -> the reference falls back to gen_sort), or NONE on scratch overflow (retry with more).
// ---------------------------------------------------------------------------------------------
struct BfsScratch {
	u32 capV, capE;
	u32 *tv;      // [capV]   tree vertex -> (vertex idx << 1) | reverse
	u32 *estart;  // [capV+1] out-edge range of a tree vertex
	u32 *eto;     // [capE]   (to << 1) | tree-edge flag
	u32 *deg;     // [capV]
	u32 *order;   // [capV]   Kahn queue == output order
	u32 *hkey;    // [hcap]   open-addressing vertex -> tree vertex map (hcap = 0: linear scan of tv)
	u32 *hval;
	u32 hcap;
};
static const u32 MAX_UNBLOCK_CTR = 10000;   // graph.hpp:22-23

PV_DEV u32 bfs_find(const BfsScratch &s, u32 n, u32 v) {
	if (s.hcap == 0) { for (u32 t = 0; t < n; t++) if ((s.tv[t] >> 1) == v) return t; return NONE; }
	u32 m = s.hcap - 1, i = (u32)(splitmix64(v) & m);
	while (true) { u32 k = s.hkey[i]; if (k == v) return s.hval[i]; if (k == NONE) return NONE; i = (i + 1) & m; }
}
PV_DEV void bfs_put(const BfsScratch &s, u32 v, u32 t) {
	if (s.hcap == 0) return;
	u32 m = s.hcap - 1, i = (u32)(splitmix64(v) & m);
	while (s.hkey[i] != NONE) i = (i + 1) & m;
	s.hkey[i] = v; s.hval[i] = t;
}

PV_DEV u32 bfs_kahn_sort(const GraphView &g, u32 a_v, u32 a_fwd, u32 z_v, u32 z_fwd, const BfsScratch &s, u32 *out /*[capV]*/) {
	for (u32 i = 0; i < s.hcap; i++) s.hkey[i] = NONE;
	u32 n = 0, ne = 0, qh = 0;
	s.tv[n] = (a_v << 1) | (a_fwd ? 0u : 1u); bfs_put(s, a_v, 0); n = 1;
	const u32 endKey = (z_v << 1) | (z_fwd ? 0u : 1u);
	while (qh < n) {
		u32 t = qh++; u32 cur = s.tv[t];
		s.estart[t] = ne;
		if (n - qh > MAX_UNBLOCK_CTR) continue;
		u32 v = cur >> 1; bool fwd = !(cur & 1u);
		u32 port = 2 * v + (fwd ? 1u : 0u);       // OUT end: right side when forward
		for (u32 k = g.adjOff[port], ke = g.adjOff[port + 1]; k < ke; k++) {
			u32 q = g.adjNbr[k]; u32 w = q >> 1;
			u32 nb = (w << 1) | ((q & 1u) ? 1u : 0u);   // entering by the left side = forward
			u32 tw = bfs_find(s, n, w);
			if (tw != NONE) {
				if (ne >= s.capE) return NONE;
				u32 to = s.tv[tw] == nb ? tw : 0u;     // parent_map[] default-inserts tree vertex 0 (graph.cpp:130-136)
				s.eto[ne++] = to << 1;
				continue;
			}
			if (cur == endKey) continue;
			if (n >= s.capV || ne >= s.capE) return NONE;
			s.tv[n] = nb; bfs_put(s, w, n);
			s.eto[ne++] = (n << 1) | 1u;
			n++;
		}
	}
	for (u32 t = qh; t <= n; t++) s.estart[t] = ne;   // (qh == n here) closes the last range
	// Kahn
	for (u32 t = 0; t < n; t++) s.deg[t] = 0;
	for (u32 t = 0; t < n; t++)
		for (u32 e = s.estart[t]; e < s.estart[t + 1]; e++) {
			u32 x = s.eto[e]; u32 to = x >> 1;
			if ((x & 1u) || to != t) s.deg[to]++;
		}
	// done flags live in the top bit of deg
	const u32 DONE = 0x80000000u;
	u32 head = 0, tail = 0;
	s.order[tail++] = 0; s.deg[0] |= DONE;
	while (head < tail) {
		u32 cur = s.order[head++];
		for (u32 e = s.estart[cur]; e < s.estart[cur + 1]; e++) {
			u32 x = s.eto[e]; u32 u = x >> 1;
			if (!(x & 1u) && u == cur) continue;
			if (s.deg[u] & DONE) continue;
			s.deg[u]--;
			if (s.deg[u] > 0) continue;
			s.order[tail++] = u; s.deg[u] |= DONE;
		}
	}
	if (tail != n) return 0;
	for (u32 k = 0; k < n; k++) out[k] = s.tv[s.order[k]] >> 1;
	return n;
}

// ---------------------------------------------------------------------------------------------
// fallback vertex order when the BFS/Kahn sort meets a cycle: the reference's haplotype-lap order
// (find_laps + gen_sort, graph.cpp:165-395).  Inherently sequential (every lap is inserted into the list the
// earlier laps built), so one thread replays it per failed region; such regions are rare (a repeat unit of two
// or more vertices inside a bubble).  Scratch: a doubly linked list in arrays + an open-addressing table
// vertex -> {list node, left / right context of the current lap}.
// ---------------------------------------------------------------------------------------------
PV_KERNEL(G_count, 64);
PV_KERNEL(G_sort, 256);
struct GsScratch {
	u32 cap, hcap;
	u32 *val, *nxt, *prv;                         // [cap]  list nodes
	u32 *hkey, *hnode, *hlE, *hlV, *hrE, *hrV;   // [hcap] vertex -> node, lap-stamped contexts
};
static const u32 GS_CRASH = 0xFFFFFFFEu;   // the reference dereferences an empty optional / map::at throws / reads v_ids[-1] here

// i-th lap of path h between the i-th visits of u and v: [start, start + len) in global step indices
struct GsLaps {
	u32 loU, loV, n;
	PV_DEV GsLaps(const GraphView &g, u32 u, u32 v, u32 h) {
		u32 s0 = g.stepOff[h], s1 = g.stepOff[h + 1];
		u32 ua = g.vsOff[u], ub = g.vsOff[u + 1], va = g.vsOff[v], vb = g.vsOff[v + 1];
		loU = lb_u32(g.vsStep, ua, ub, s0); u32 hiU = lb_u32(g.vsStep, ua, ub, s1);
		loV = lb_u32(g.vsStep, va, vb, s0); u32 hiV = lb_u32(g.vsStep, va, vb, s1);
		u32 nu = hiU - loU, nv = hiV - loV;
		n = nu < nv ? nu : nv;
	}
	PV_DEV void lap(const GraphView &g, u32 i, u32 &start, u32 &len) const {
		u32 s = g.vsStep[loU + i], e = g.vsStep[loV + i];
		if (s > e) { u32 t = s; s = e; e = t; }
		start = s; len = e - s + 1;
	}
};
PV_DEV u32 gs_slot(const GsScratch &s, u32 v) {
	u32 m = s.hcap - 1, i = (u32)(splitmix64(v) & m);
	while (true) {
		u32 k = s.hkey[i];
		if (k == v) return i;
		if (k == NONE) { s.hkey[i] = v; s.hnode[i] = NONE; s.hlE[i] = 0; s.hrE[i] = 0; return i; }
		i = (i + 1) & m;
	}
}
// returns the list length (the order in out[]), 0 for an empty list, GS_CRASH where the reference has undefined behaviour
PV_DEV u32 gen_sort_laps(const GraphView &g, u32 a_v, u32 z_v, const GsScratch &s, u32 *out) {
	for (u32 i = 0; i < s.hcap; i++) s.hkey[i] = NONE;
	u32 n = 0, head = NONE, tail = NONE, epoch = 0;
	for (u32 h = 0; h < g.H; h++) {
		GsLaps L(g, a_v, z_v, h);
		for (u32 li = 0; li < L.n; li++) {
			u32 start, len; L.lap(g, li, start, len);
			if (n == 0) {   // first lap: the list is the lap's vertices in first-visit order (graph.cpp:282-296)
				for (u32 j = start; j < start + len; j++) {
					u32 sl = gs_slot(s, g.steps[j] >> 1);
					if (s.hnode[sl] != NONE) continue;
					s.val[n] = g.steps[j] >> 1; s.nxt[n] = NONE; s.prv[n] = tail;
					if (tail != NONE) s.nxt[tail] = n; else head = n;
					tail = n; s.hnode[sl] = n; n++;
				}
				continue;
			}
			if (start == g.stepOff[h]) return GS_CRASH;   // `for (j = start + len - 1; j >= start; j--)` with u32 j and start == 0
			epoch++;
			u32 x = NONE;
			for (u32 j = start; j < start + len; j++) {   // left contexts (graph.cpp:246-262)
				u32 sl = gs_slot(s, g.steps[j] >> 1);
				if (s.hlE[sl] == epoch) continue;
				if (s.hnode[sl] != NONE) { s.hlE[sl] = epoch; s.hlV[sl] = s.hnode[sl]; }
				else if (j > start) {
					u32 ps = gs_slot(s, g.steps[j - 1] >> 1);
					if (s.hnode[ps] != NONE) x = s.hnode[ps];
					if (x == NONE) return GS_CRASH;
					s.hlE[sl] = epoch; s.hlV[sl] = x;
				}
			}
			u32 y = NONE;
			for (u32 j = start + len - 1;; j--) {   // right contexts (graph.cpp:264-280)
				u32 sl = gs_slot(s, g.steps[j] >> 1);
				if (s.hrE[sl] != epoch) {
					if (s.hnode[sl] != NONE) { s.hrE[sl] = epoch; s.hrV[sl] = s.hnode[sl]; }
					else if (j + 1 < start + len) {
						u32 ns = gs_slot(s, g.steps[j + 1] >> 1);
						if (s.hnode[ns] != NONE) y = s.hnode[ns];
						if (y == NONE) return GS_CRASH;
						s.hrE[sl] = epoch; s.hrV[sl] = y;
					}
				}
				if (j == start) break;
			}
			for (u32 j = start; j < start + len; j++) {   // insert the new vertices (graph.cpp:336-372)
				u32 sl = gs_slot(s, g.steps[j] >> 1);
				if (s.hnode[sl] != NONE) continue;
				if (s.hlE[sl] != epoch || s.hrE[sl] != epoch) return GS_CRASH;   // map::at on a vertex without a context
				u32 l = s.hlV[sl], r = s.hrV[sl];
				u32 ip = l;
				if (l != r) { u32 q = head; while (q != l && q != r) q = s.nxt[q]; if (q == l) ip = r; }   // distance(l) < distance(r) ? r : l
				if (n >= s.cap) return GS_CRASH;   // cannot happen: cap counts every lap step
				u32 p = s.prv[ip];
				s.val[n] = g.steps[j] >> 1; s.nxt[n] = ip; s.prv[n] = p; s.prv[ip] = n;
				if (p != NONE) s.nxt[p] = n; else head = n;
				s.hnode[sl] = n; n++;
			}
		}
	}
	u32 k = 0;
	for (u32 q = head; q != NONE; q = s.nxt[q]) out[k++] = s.val[q];
	return n;
}

// ---------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------
struct CallIn {           // PVST forest, flubbles of all components concatenated (host arrays)
	u32 nFl = 0;
	const u32 *a_v = nullptr, *z_v = nullptr; const u8 *a_fwd = nullptr, *z_fwd = nullptr;
	const u32 *parent = nullptr;   // PVST idx inside the component (0 = dummy root; flubble k of the component is k+1)
	const u32 *flBase = nullptr;   // per flubble: index of the first flubble of its component
	u32 nRefs = 0; const u32 *refs = nullptr;   // ascending path ids to call against
	const u8 *excluded = nullptr;  // optional: 1 = never selected (sub-flubble clan or below one, color.cpp:117-125)
	// The arrays may be a SLICE of the forest: the flubbles of the one component that holds every selected reference path
	// (no flubble of another component can be selected: both boundaries must lie on every reference path).  fl_offset =
	// global index of the slice's first flubble; flBase values stay global, region -> flubble indices come out global.
	u32 fl_offset = 0;
	bool on_device = false;        // a_v .. flBase point into device memory (decompose's table handed over in HBM)
	// -g/--restrict ref:start-end (rov.cpp:137-183): keep a selected flubble only if a step of the named path on one of its
	// two boundary vertices starts inside [start, end).  restrict_ref == NONE: the path does not exist -> no regions.
	bool restrict_on = false; u32 restrict_ref = NONE; u64 restrict_start = 0, restrict_end = 0;
};

// binary records, written by the device into page-locked host vectors whose capacity survives across calls
// the records of one (chunk, ref) emission as they were produced in HBM; offsets (grpOff, gtOff) are already global
struct RecPiece {
	Buf<u32> rov, ref, cxtL, cxtR, vt, pos, refStart, refLen, grpOff /* nRec + 1 */, gHap, gStart, gLen, gCnt; Buf<u64> gtOff /* nRec + 1 */; Buf<u16> gt;
	u32 nRec = 0, nGrp = 0; u64 nGt = 0;
};
struct Records {
	// Records are born in HBM (`dev`) and stay there for the device text writer; the host vectors below are filled on
	// demand by to_host() (povu_ffi / record_lines / the bench's e2e leg).
	std::vector<RecPiece> dev; size_t copied = 0;       // pieces [0, copied) are already in the host vectors
	size_t nRecTot = 0, nGrpTot = 0; u64 nGtTot = 0;
	// regions (RoVs) in enumeration order: ascending (component file, PVST vertex idx), >= 3 sorted vertices
	hvec<u32> rov_fl, rov_comp, rov_tangled;
	// records: grouped per (region, ref) in context order, INS / SUB / DEL
	hvec<u32> rec_rov, rec_ref, rec_cxtL, rec_cxtR, rec_vt, rec_pos, rec_refStart, rec_refLen, rec_grpOff;
	hvec<u64> rec_gtOff;
	hvec<u32> grp_hap, grp_start, grp_len, grp_count;   // start = step index inside the path
	hvec<u16> gt;           // per record: one code per path of the region's component (0 ref, i alt group, 0xFFFF missing)
	std::vector<u32> cpOff, cpList;   // paths of every component = the columns of `gt`
	u64 n_subr_clamped = 0;   // SUBR records whose merged length ran past the end of a walk (the reference reads out of bounds there)
	u64 n_regions = 0;   // real regions; rov_* entries beyond them are per-ref pseudo regions that carry the SUBR records
	u64 n_called = 0, n_too_small = 0, n_sort_failed = 0, n_tangled = 0, n_pin_pairs = 0, n_hard_errors = 0, n_chunks = 0, n_big_bfs = 0, max_J = 0;
	u64 n_slice_tree_replays = 0;   // ref paths whose slice tree was replayed sequentially (slices of one alt path overlap); 0 = every tree took the sorted fast path
	size_t n() const { return nRecTot; }
	void add_piece(RecPiece &&p) { nRecTot += p.nRec; nGrpTot += p.nGrp; nGtTot += p.nGt; dev.push_back(std::move(p)); }
	// append the device pieces that have not been copied yet to the page-locked host vectors
	void to_host(Ctx &c) {
		if (copied == dev.size()) return;
		rec_rov.reserve(nRecTot); rec_ref.reserve(nRecTot); rec_cxtL.reserve(nRecTot); rec_cxtR.reserve(nRecTot); rec_vt.reserve(nRecTot); rec_pos.reserve(nRecTot);
		rec_refStart.reserve(nRecTot); rec_refLen.reserve(nRecTot); rec_grpOff.reserve(nRecTot + 1); rec_gtOff.reserve(nRecTot + 1);
		grp_hap.reserve(nGrpTot); grp_start.reserve(nGrpTot); grp_len.reserve(nGrpTot); grp_count.reserve(nGrpTot); gt.reserve(nGtTot);   // no reallocation under the copies in flight
		for (; copied < dev.size(); copied++) {
			const RecPiece &p = dev[copied];
			const size_t recBase = rec_rov.size(); const u64 gtBase = gt.size();
			auto app = [&](hvec<u32> &dst, const Buf<u32> &src, u32 m) { size_t o = dst.size(); dst.resize(o + m); c.d2h_async(dst.data() + o, src.p, (size_t)m * 4); };
			app(rec_rov, p.rov, p.nRec); app(rec_ref, p.ref, p.nRec); app(rec_cxtL, p.cxtL, p.nRec); app(rec_cxtR, p.cxtR, p.nRec); app(rec_vt, p.vt, p.nRec); app(rec_pos, p.pos, p.nRec);
			app(rec_refStart, p.refStart, p.nRec); app(rec_refLen, p.refLen, p.nRec);
			app(grp_hap, p.gHap, p.nGrp); app(grp_start, p.gStart, p.nGrp); app(grp_len, p.gLen, p.nGrp); app(grp_count, p.gCnt, p.nGrp);
			rec_grpOff.resize(recBase + p.nRec + 1); c.d2h_async(rec_grpOff.data() + recBase, p.grpOff.p, ((size_t)p.nRec + 1) * 4);
			rec_gtOff.resize(recBase + p.nRec + 1); c.d2h_async(rec_gtOff.data() + recBase, p.gtOff.p, ((size_t)p.nRec + 1) * 8);
			gt.resize(gtBase + p.nGt); c.d2h_async(gt.data() + gtBase, p.gt.p, p.nGt * 2);
		}
		c.sync();
	}
	// all records as ONE device piece (the text writer reads flat arrays): several pieces are concatenated once
	const RecPiece &flat(Ctx &c) {
		if (dev.size() == 1) return dev[0];
		RecPiece f; f.nRec = (u32)nRecTot; f.nGrp = (u32)nGrpTot; f.nGt = nGtTot;
		const size_t R1 = nRecTot ? nRecTot : 1, G1 = nGrpTot ? nGrpTot : 1;
		f.rov = Buf<u32>(c, R1); f.ref = Buf<u32>(c, R1); f.cxtL = Buf<u32>(c, R1); f.cxtR = Buf<u32>(c, R1); f.vt = Buf<u32>(c, R1); f.pos = Buf<u32>(c, R1);
		f.refStart = Buf<u32>(c, R1); f.refLen = Buf<u32>(c, R1); f.grpOff = Buf<u32>(c, R1 + 1); f.gtOff = Buf<u64>(c, R1 + 1);
		f.gHap = Buf<u32>(c, G1); f.gStart = Buf<u32>(c, G1); f.gLen = Buf<u32>(c, G1); f.gCnt = Buf<u32>(c, G1); f.gt = Buf<u16>(c, nGtTot ? nGtTot : 1);
		size_t ro = 0, go = 0; u64 to = 0;
		if (dev.empty()) { c.memset0(f.grpOff.p, 4); c.memset0(f.gtOff.p, 8); }
		for (const RecPiece &p : dev) {
			c.d2d(f.rov.p + ro, p.rov.p, (size_t)p.nRec * 4); c.d2d(f.ref.p + ro, p.ref.p, (size_t)p.nRec * 4); c.d2d(f.cxtL.p + ro, p.cxtL.p, (size_t)p.nRec * 4);
			c.d2d(f.cxtR.p + ro, p.cxtR.p, (size_t)p.nRec * 4); c.d2d(f.vt.p + ro, p.vt.p, (size_t)p.nRec * 4); c.d2d(f.pos.p + ro, p.pos.p, (size_t)p.nRec * 4);
			c.d2d(f.refStart.p + ro, p.refStart.p, (size_t)p.nRec * 4); c.d2d(f.refLen.p + ro, p.refLen.p, (size_t)p.nRec * 4);
			c.d2d(f.grpOff.p + ro, p.grpOff.p, ((size_t)p.nRec + 1) * 4); c.d2d(f.gtOff.p + ro, p.gtOff.p, ((size_t)p.nRec + 1) * 8);
			c.d2d(f.gHap.p + go, p.gHap.p, (size_t)p.nGrp * 4); c.d2d(f.gStart.p + go, p.gStart.p, (size_t)p.nGrp * 4); c.d2d(f.gLen.p + go, p.gLen.p, (size_t)p.nGrp * 4);
			c.d2d(f.gCnt.p + go, p.gCnt.p, (size_t)p.nGrp * 4); c.d2d(f.gt.p + to, p.gt.p, p.nGt * 2);
			ro += p.nRec; go += p.nGrp; to += p.nGt;
		}
		if (copied != 0 && copied != dev.size()) to_host(c);   // a host copy that was begun from the pieces is finished from them
		const bool on_host = copied != 0 || dev.empty();
		c.sync();
		dev.clear(); dev.push_back(std::move(f)); copied = on_host ? 1 : 0;
		return dev[0];
	}
	void reset() {
		dev.clear(); copied = 0; nRecTot = nGrpTot = 0; nGtTot = 0;
		rov_fl.clear(); rov_comp.clear(); rov_tangled.clear();
		rec_rov.clear(); rec_ref.clear(); rec_cxtL.clear(); rec_cxtR.clear(); rec_vt.clear(); rec_pos.clear(); rec_refStart.clear(); rec_refLen.clear();
		rec_grpOff.clear(); rec_gtOff.clear(); grp_hap.clear(); grp_start.clear(); grp_len.clear(); grp_count.clear(); gt.clear();
		n_regions = 0; n_subr_clamped = 0; n_slice_tree_replays = 0; n_called = n_too_small = n_sort_failed = n_tangled = n_pin_pairs = n_hard_errors = n_chunks = n_big_bfs = max_J = 0;
	}
};
typedef Records CallOut;
enum { VT_DEL = 0, VT_INS = 1, VT_SUB = 2, VT_SUBR = 3 };

struct CallOpts { u64 cell_budget = 1ull << 30; };

struct RovView {
	const u32 *rovOff, *rovVerts, *rovComp;   // per region (global index): sorted vertices, component
	const u64 *rowPre, *cellPre;             // prefix sums of matrix rows / cells over the view's items
	const u64 *hkey; const u32 *hval; u64 hmask;
	u32 R; u32 uniformNh; u32 rShift;
	const u32 *instJ; // ref lap / alignment column of an unrolled matrix (nullptr for plain regions)
	const u32 *map;   // nullptr: item r of the view is region rShift + r (a chunk of regions);
	                  // else item r is an unrolled matrix of region map[r] (tangled regions) and rShift = 0
	PV_DEV u32 region(u32 r) const { return map ? map[r] : rShift + r; }
	PV_DEV u32 J(u32 r) const { u32 q = region(r); return rovOff[q + 1] - rovOff[q]; }
	PV_DEV u32 vert(u32 r, u32 j) const { return rovVerts[rovOff[region(r)] + j]; }
	PV_DEV u32 comp(u32 r) const { return rovComp[region(r)]; }
	PV_DEV u32 rowBase(u32 r) const { return (u32)(rowPre[rShift + r] - rowPre[rShift]); }   // first matrix row of item r (view-relative)
	PV_DEV u64 matBase(u32 r) const { return cellPre[rShift + r] - cellPre[rShift]; }
	PV_DEV u32 find(u32 r, u32 v) const {   // column of vertex v in item r, or NONE
		u32 q = region(r);
		u32 a = rovOff[q], n = rovOff[q + 1] - a;
		if (n <= 16) { for (u32 j = 0; j < n; j++) if (rovVerts[a + j] == v) return j; return NONE; }
		u64 key = ((u64)q << 32) | v; u64 i = splitmix64(key) & hmask;
		while (true) { u64 k = hkey[i]; if (k == key) return hval[i]; if (k == ~0ull) return NONE; i = (i + 1) & hmask; }
	}
	PV_DEV u32 row2rov(u32 row) const {
		if (uniformNh) return row / uniformNh;
		return (u32)(ub_u64(rowPre + rShift, 0, (u64)R + 1, rowPre[rShift] + row) - 1);
	}
};

struct Lap { u32 lo, hi; };   // global step range [lo, hi)

// hap_sl_from_lap (overlay.cpp:189-246): [first start .. last end] of the context vertices inside a lap
PV_DEV bool slice_bounds(const GraphView &g, u32 lv, u32 rv, Lap lap, u32 &start, u32 &len) {
	u32 lf = g.occ_first(lv, lap.lo, lap.hi), rf = g.occ_first(rv, lap.lo, lap.hi);
	if (lf == NONE || rf == NONE) return false;
	u32 end;
	if (lf > rf) { start = rf; end = g.occ_last(lv, lap.lo, lap.hi); }
	else { start = lf; end = g.occ_last(rv, lap.lo, lap.hi); }
	len = end - start + 1;
	return true;
}
PV_DEV bool walk_equal(const u32 *steps, u32 a, u32 b, u32 len) {
	for (u32 i = 0; i < len; i++) if (steps[a + i] != steps[b + i]) return false;
	return true;
}
PV_DEV bool walk_revcomp(const u32 *steps, u32 a, u32 b, u32 len) {   // is_inv_local (overlay.cpp:167-187)
	for (u32 i = 0; i < len; i++) if (steps[a + i] != (steps[b + len - 1 - i] ^ 1u)) return false;
	return true;
}
// std::vector<id_or_t> operator< (types.cpp:112-120): packed steps compare like (v_id, forward < reverse)
PV_DEV int walk_cmp(const u32 *steps, u32 a, u32 la, u32 b, u32 lb) {
	u32 n = la < lb ? la : lb;
	for (u32 i = 0; i < n; i++) { u32 x = steps[a + i], y = steps[b + i]; if (x != y) return x < y ? -1 : 1; }
	return la == lb ? 0 : (la < lb ? -1 : 1);
}

// per-chunk working set shared by the overlay kernels
struct ChunkView {
	GraphView g; RovView rv;
	u16 *mat; u32 *lapStart, *lapLen, *tangled, *hasData, *errs;

	PV_DEV Lap lap(u32 row) const { u32 s = lapStart[row]; return Lap{s, s + lapLen[row]}; }
};
struct CandView { u32 *cL, *cR, *cStart, *cLen, *cRS, *cRL, *cHap, *cCls, *cHead, *gCount, *hsort, *pinCount;
		  u32 *evald; };   // evald[region]: a context of this ref was looked at (what makes trek::has_data() true, see overlay_pass)

// comp_bounds + is_inverted over one (region, ref path, other path) pair (overlay.cpp:84-149); with
// fill = true also gen_cxt, the two hap slices and the classification of every context
// (overlay.cpp:307-327,189-246,462-530).  Returns the number of contexts.
PV_DEV u32 scan_pair_ri(const ChunkView &cv, u32 refLocal, u32 k, u32 r, u32 i, u32 row, bool fill, const CandView &cd, u32 off);
PV_DEV u32 scan_pair(const ChunkView &cv, u32 refLocal, u32 k, u32 row, bool fill, const CandView &cd, u32 off) {
	u32 r = cv.rv.row2rov(row);
	return scan_pair_ri(cv, refLocal, k, r, row - cv.rv.rowBase(r), row, fill, cd, off);
}
// the same with the region item r and the row's index i inside it already known (row = rowBase(r) + i)
PV_DEV u32 scan_pair_ri(const ChunkView &cv, u32 refLocal, u32 k, u32 r, u32 i, u32 row, bool fill, const CandView &cd, u32 off) {
	const RovView &rv = cv.rv; const GraphView &g = cv.g;
	if (i == refLocal || ld_volatile(cv.tangled + r)) return 0;      // (volatile: in the fused region kernel another lane's atomic set it moments ago)
	if (k > 0 && !cv.hasData[r]) return 0;     // comp_exps returned after the first ref (overlay.cpp:531-532)
	u32 J = rv.J(r);
	const u16 *rr = cv.mat + rv.matBase(r) + (u64)refLocal * J, *ar = cv.mat + rv.matBase(r) + (u64)i * J;
	bool empty = true, inv = true; u32 cnt = 0; u32 prev = NONE; bool diff = false;
	for (u32 j = 0; j < J; j++) {
		u32 x = rr[j], y = ar[j];
		if (y) empty = false;
		bool ci = (x == 1 && y == 2) || (x == 2 && y == 1);
		if (!ci && !(x == 0 && y == 0)) inv = false;
		bool ok = x > 0 && (x == y || ci);
		if (ok) {
			if (prev != NONE && diff) {
				if (fill) {
					cd.evald[r] = 1;      // tk.get_min_rov(ref) default-inserts the ref's map for EVERY context, before the pin / equal-slice tests (overlay.cpp:478-479)
					u32 u = rv.vert(r, prev), v = rv.vert(r, j);
					Lap rl = cv.lap(rv.rowBase(r) + refLocal), al = cv.lap(row);
					u32 fu = g.occ_first(u, rl.lo, rl.hi), lv = g.occ_last(v, rl.lo, rl.hi);
					u32 o = off + cnt; cd.cCls[o] = 3;
					u32 rs = 0, rn = 0, as = 0, an = 0;
					if (fu == NONE || lv == NONE || !slice_bounds(g, u, v, rl, rs, rn) || !slice_bounds(g, u, v, al, as, an)) atomic_add(cv.errs + 0, 1u);   // the reference exits / crashes here
					else if (rn == an && walk_revcomp(g.steps, rs, as, rn)) { cd.cCls[o] = 4; cd.cRS[o] = rs; cd.cStart[o] = as; cd.cHap[o] = i; atomic_add(cd.pinCount, 1u); }   // inversion pin pair for SNE (overlay.cpp:487-496)
					else if (rn == an && walk_equal(g.steps, rs, as, rn)) { }
					else {
						cd.cL[o] = g.steps[fu]; cd.cR[o] = g.steps[lv]; cd.cStart[o] = as; cd.cLen[o] = an; cd.cRS[o] = rs; cd.cRL[o] = rn; cd.cHap[o] = i;
						cd.cCls[o] = an == rn ? 1u : (an > rn ? 0u : 2u);   // 0 INS, 1 SUB, 2 DEL (emission order, vcf.cpp:253-266)
					}
				}
				cnt++;
			}
			prev = j; diff = false;
		} else if (x != y) diff = true;
	}
	if (empty) return 0;
	if (inv) {   // every shared column is visited in opposite orientation: one pin pair (ref lap front, alt lap back), overlay.cpp:455-460
		if (fill) {
			Lap rl = cv.lap(rv.rowBase(r) + refLocal), al = cv.lap(row);
			cd.cCls[off] = 3;
			if (rl.lo == NONE || al.lo == NONE || rl.hi == rl.lo || al.hi == al.lo) atomic_add(cv.errs + 0, 1u);
			else { cd.cCls[off] = 4; cd.cRS[off] = rl.lo; cd.cStart[off] = al.hi - 1; cd.cHap[off] = i; atomic_add(cd.pinCount, 1u); }
		}
		return 1;
	}
	return cnt;
}

// order of the reference's nested std::maps: context (l, r), then INS < SUB < DEL, then the walk
PV_DEV bool head_less(const u32 *steps, const CandView &cd, u32 a, u32 b) {
	if (cd.cL[a] != cd.cL[b]) return cd.cL[a] < cd.cL[b];
	if (cd.cR[a] != cd.cR[b]) return cd.cR[a] < cd.cR[b];
	if (cd.cCls[a] != cd.cCls[b]) return cd.cCls[a] < cd.cCls[b];
	return walk_cmp(steps, cd.cStart[a], cd.cLen[a], cd.cStart[b], cd.cLen[b]) < 0;
}
PV_DEV bool same_record(const CandView &cd, u32 a, u32 b) { return cd.cL[a] == cd.cL[b] && cd.cR[a] == cd.cR[b] && cd.cCls[a] == cd.cCls[b]; }

PV_DEV void sift_down(const u32 *steps, const CandView &cd, u32 *a, u32 root, u32 end) {
	while (2 * root + 1 < end) {
		u32 ch = 2 * root + 1;
		if (ch + 1 < end && head_less(steps, cd, a[ch], a[ch + 1])) ch++;
		if (!head_less(steps, cd, a[root], a[ch])) return;
		u32 t = a[root]; a[root] = a[ch]; a[ch] = t; root = ch;
	}
}

// inversion pin pairs collected while overlaying (ise::pin_cushion, /root/reference/include/ita/variation/sne.hpp:112-170);
// kept on the device in pieces, one per overlay pass, with the key that restores the reference's insertion order
struct PinStore {
	struct Piece { Buf<u32> keyR, keyJ, keyK, keySlot, refStep, altStep; u32 n; };
	std::vector<Piece> pieces; u64 total = 0;
};
PV_KERNEL(S_pinflag, 8);
PV_KERNEL(S_pinpack, 40);

// comp_exps over every matrix of a view (overlay.cpp:348-588) for all ref paths, then records + genotypes
// (vcf.cpp:20-121,241-273); results are appended to `out`.
// tangled_seen (optional): receives errs[3] (tangled regions counted so far) with the first host wait of the pass
inline void overlay_pass(Ctx &c, const DevGraph &G, const GraphView &g, const CallIn &in, const RovView &rv, const ChunkView &cv, u32 Rc, u32 nRows,
			 Records &out, Trace &tr, PinStore &pins, u32 *tangled_seen = nullptr, Buf<u32> *pairCnt0 = nullptr) {
	const u32 nRefs = in.nRefs; const u32 uniformNh = rv.uniformNh;
	for (u32 k = 0; k < nRefs; k++) {
		const u32 refH = in.refs[k];
		const u32 refLocal = G.hapLocal[refH];
		const u32 refStep0 = G.h_stepOff[refH];
		// ---- C5a: contexts per (region, path) pair
		Buf<u32> pairCnt, candOff(c, (size_t)nRows + 1);
		if (k == 0 && pairCnt0) pairCnt = std::move(*pairCnt0);      // counted by the fused region kernel
		else {
			pairCnt = Buf<u32>(c, (size_t)nRows + 1);
			u32 *pairCnt_ = pairCnt.p;
			if ((u64)nRows <= 32ull * Rc) {
				// few rows per region (a handful of samples): one thread per REGION walks its rows -- the region's geometry (row base,
				// matrix base, width) is looked up once instead of once per row, and its rows are one contiguous block of the matrix
				parallel_for<C_pair_count>(c, (size_t)Rc + 1, PV_LAMBDA(size_t rr_) {
					if (rr_ >= Rc) { pairCnt_[nRows] = 0; return; }
					const u32 r = (u32)rr_, rb = rv.rowBase(r), nh = rv.rowBase(r + 1) - rb;
					for (u32 i = 0; i < nh; i++) pairCnt_[rb + i] = scan_pair_ri(cv, refLocal, k, r, i, rb + i, false, CandView{}, 0);
				});
			} else
			parallel_for<C_pair_count>(c, (size_t)nRows + 1, PV_LAMBDA(size_t row) {
				pairCnt_[row] = row < nRows ? scan_pair(cv, refLocal, k, (u32)row, false, CandView{}, 0) : 0u;
			});
		}
		const u32 *candOff_ = candOff.p;
		exclusive_scan(c, pairCnt.p, candOff.p, (size_t)nRows + 1, OpAddU32(), 0u);
		u32 nCand = 0;
		{ MultiRead mr(c); mr.add(&nCand, candOff.p + nRows, 4); if (tangled_seen && k == 0) mr.add(tangled_seen, cv.errs + 3, 4); mr.wait(); }
		const size_t nC1 = nCand ? nCand : 1;
		Buf<u32> cL(c, nC1), cR(c, nC1), cStart(c, nC1), cLen(c, nC1), cRS(c, nC1), cRL(c, nC1), cHap(c, nC1), cCls(c, nC1), cHead(c, nC1), gCount(c, nC1), hsort(c, nC1);
		Buf<u32> pinCount(c, 1); pinCount.zero();
		// has_data() of the reference's trek is "some ref's map exists", and a ref's map comes into being with the first context that is
		// looked at for it -- also when that context ends as an inversion pin pair or an equal slice and leaves no record.  Regions
		// whose first ref only produced such contexts therefore DO go on to the next ref (a whole-row inversion does not count).
		Buf<u32> evald(c, (size_t)Rc + 1); evald.zero();
		const CandView cd{cL.p, cR.p, cStart.p, cLen.p, cRS.p, cRL.p, cHap.p, cCls.p, cHead.p, gCount.p, hsort.p, pinCount.p, evald.p};
		// ---- C5b: slices per context, classification
		// one thread per candidate slot; the thread of a row's FIRST slot fills the whole row.  Rows without a context
		// (most paths carry the ref allele) cost nothing and the working lanes of a warp are consecutive busy rows.
		if (nCand) parallel_for<C_pair_fill>(c, nCand, PV_LAMBDA(size_t x) {
			u32 row = ub_u32(candOff_, 0, nRows + 1, (u32)x) - 1;
			if (candOff_[row] != (u32)x) return;
			scan_pair(cv, refLocal, k, row, true, cd, (u32)x);
		});
		tr.mark("C5 pairs");
		// ---- C6a: group alt slices by (context, class, walk) per region, order groups like the reference's std::maps
		Buf<u32> recCnt(c, (size_t)Rc + 1), grpCnt(c, (size_t)Rc + 1), recOff(c, (size_t)Rc + 1), grpOff(c, (size_t)Rc + 1);
		u32 *recCnt_ = recCnt.p, *grpCnt_ = grpCnt.p;
		bool group_by_warp = false;
#ifndef PV_EMU
		// hundreds of paths: a region has ~100+ candidates that one thread would compare one by one (1.7 GB/s of algorithmic
		// traffic at 464 paths).  One warp per region instead: repeatedly take the first ungrouped candidate as the next head
		// (lanes stride over the candidates, warp minimum), let every lane compare its candidates with it, count the members
		// with a warp sum.  Heads come out in order of first occurrence and every candidate joins the first head it matches --
		// exactly the sequential grouping; the few heads are then ordered by lane 0.
		u64 warp_min = 8;      // candidates per region from which a warp beats a thread (the thread version ran at 17 GB/s with ~20 candidates per region)
		if (const char *e_ = getenv("POVU_B200_GROUP_WARP_MIN")) warp_min = strtoull(e_, nullptr, 10);
		group_by_warp = (u64)nCand >= warp_min * Rc || getenv("POVU_B200_GROUP_WARP") != nullptr;
		if (group_by_warp && !getenv("POVU_B200_GROUP_THREAD")) {
			u32 *cHead_ = cd.cHead;
			parallel_for<C_fill16>(c, nC1, PV_LAMBDA(size_t x) { cHead_[x] = NONE; });
			parallel_for<C_group>(c, ((size_t)Rc + 1) * 32, PV_LAMBDA(size_t t) {
				const u32 r = (u32)(t >> 5), lane = (u32)(t & 31);
				if (r >= Rc) { if (lane == 0) { recCnt_[r] = 0; grpCnt_[r] = 0; } return; }
				const u32 b = candOff_[rv.rowBase(r)], e = candOff_[rv.rowBase(r + 1)];
				u32 *a = cd.hsort + b; u32 nh = 0;
				u32 kL = 0, kR = 0, kC = 0, kN = 0, kS = 0, kH = NONE; bool inRegs = false;      // head number `lane` (register path, first 32 heads)
				if (e - b <= 128u) {
					inRegs = true;
					// up to four candidates per lane live in registers for the whole grouping: every field is read ONCE (ncu at 90
					// paths: 2.4 GB read for 0.6 GB of candidates, one sweep of the arrays per head), the head's fields come from its
					// owner lane by shuffle, membership is a ballot.  Candidate x = b + 32 k + lane: ascending k, then lane, is
					// ascending x, so "first ungrouped candidate" and "joins the first head it matches" are the sequential rules.
					u32 fL[4], fR[4], fC[4], fN[4], fS[4]; bool un[4];
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const u32 x = b + 32u * k + lane; const bool v = x < e;
						fC[k] = v ? cd.cCls[x] : 3u; un[k] = v && fC[k] < 3u;
						fL[k] = un[k] ? cd.cL[x] : 0u; fR[k] = un[k] ? cd.cR[x] : 0u; fN[k] = un[k] ? cd.cLen[x] : 0u; fS[k] = un[k] ? cd.cStart[x] : 0u;
					}
					while (true) {
						int hk = -1; u32 hm = 0;
#pragma unroll
						for (int k = 0; k < 4; k++) { const u32 m = __ballot_sync(0xFFFFFFFFu, un[k]); if (hk < 0 && m) { hk = k; hm = m; } }
						if (hk < 0) break;
						const int hl = __ffs((int)hm) - 1;
						u32 sL = 0, sR = 0, sC = 0, sN = 0, sS = 0;
#pragma unroll
						for (int k = 0; k < 4; k++) if (k == hk) { sL = fL[k]; sR = fR[k]; sC = fC[k]; sN = fN[k]; sS = fS[k]; }
						const u32 hL = __shfl_sync(0xFFFFFFFFu, sL, hl), hR = __shfl_sync(0xFFFFFFFFu, sR, hl), hC = __shfl_sync(0xFFFFFFFFu, sC, hl),
							  hLen = __shfl_sync(0xFFFFFFFFu, sN, hl), hSt = __shfl_sync(0xFFFFFFFFu, sS, hl);
						const u32 h = b + 32u * (u32)hk + (u32)hl;
						if (lane == nh) { kL = hL; kR = hR; kC = hC; kN = hLen; kS = hSt; kH = h; }
						u32 mine = 0;
#pragma unroll
						for (int k = 0; k < 4; k++) {
							if (!un[k]) continue;
							const u32 x = b + 32u * k + lane;
							if (x == h || (fL[k] == hL && fR[k] == hR && fC[k] == hC && fN[k] == hLen && walk_equal(g.steps, hSt, fS[k], hLen))) { cd.cHead[x] = h; un[k] = false; mine++; }
						}
#pragma unroll
						for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
						if (lane == 0) { a[nh] = h; cd.gCount[h] = mine; }
						nh++;
					}
				} else {
				u32 next = b + lane;          // this lane's first candidate that may still be ungrouped
				while (true) {
					while (next < e && (cd.cCls[next] >= 3 || cd.cHead[next] != NONE)) next += 32;
					u32 h = next < e ? next : NONE;
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) { u32 o = __shfl_xor_sync(0xFFFFFFFFu, h, d); h = o < h ? o : h; }
					if (h == NONE) break;
					const u32 hL = cd.cL[h], hR = cd.cR[h], hC = cd.cCls[h], hLen = cd.cLen[h], hSt = cd.cStart[h];
					u32 mine = 0;
					for (u32 x = next; x < e; x += 32) {
						if (x == h) { cd.cHead[x] = h; mine++; continue; }
						if (cd.cCls[x] >= 3 || cd.cHead[x] != NONE) continue;
						if (cd.cL[x] == hL && cd.cR[x] == hR && cd.cCls[x] == hC && cd.cLen[x] == hLen && walk_equal(g.steps, hSt, cd.cStart[x], hLen)) { cd.cHead[x] = h; mine++; }
					}
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
					if (lane == 0) { a[nh] = h; cd.gCount[h] = mine; }
					nh++;
				}
				}
				__syncwarp();      // (lane 0 orders the heads from the arrays: the members' stores above must have landed)
				if (inRegs && nh <= 32u) {
					// the heads' order (context, class, walk: the reference's nested std::maps) by RANKING instead of a heap sort on one
					// lane: lane j holds head j, meets every other head by shuffle and counts the smaller ones; distinct heads never
					// compare equal, so the ranks are a permutation.  A head opens a record iff no smaller head shares its context + class.
					u32 rank = 0; bool prevSame = false;
					for (u32 j = 0; j < nh; j++) {
						const u32 oL = __shfl_sync(0xFFFFFFFFu, kL, j), oR = __shfl_sync(0xFFFFFFFFu, kR, j), oC = __shfl_sync(0xFFFFFFFFu, kC, j),
							  oN = __shfl_sync(0xFFFFFFFFu, kN, j), oS = __shfl_sync(0xFFFFFFFFu, kS, j);
						if (lane >= nh || j == lane) continue;
						const bool same = oL == kL && oR == kR && oC == kC;
						const bool less = oL != kL ? oL < kL : (oR != kR ? oR < kR : (oC != kC ? oC < kC : walk_cmp(g.steps, oS, oN, kS, kN) < 0));
						if (less) { rank++; if (same) prevSame = true; }
					}
					if (lane < nh) a[rank] = kH;
					const u32 nrec = (u32)__popc(__ballot_sync(0xFFFFFFFFu, lane < nh && !prevSame));
					if (lane == 0) {
						recCnt_[r] = nrec; grpCnt_[r] = nh;
						if (k == 0) cv.hasData[r] = cd.evald[r];
						else if (!cd.evald[r] && cv.hasData[r]) atomic_add(cv.errs + 0, 1u);   // std::map::at throws in the reference
					}
					return;
				}
				if (lane != 0) return;
				for (u32 st = nh / 2; st-- > 0;) sift_down(g.steps, cd, a, st, nh);
				for (u32 end = nh; end-- > 1;) { u32 t2 = a[0]; a[0] = a[end]; a[end] = t2; sift_down(g.steps, cd, a, 0, end); }
				u32 nrec = 0;
				for (u32 q = 0; q < nh; q++) if (q == 0 || !same_record(cd, a[q - 1], a[q])) nrec++;
				recCnt_[r] = nrec; grpCnt_[r] = nh;
				if (k == 0) cv.hasData[r] = cd.evald[r];
				else if (!cd.evald[r] && cv.hasData[r]) atomic_add(cv.errs + 0, 1u);   // std::map::at throws in the reference
			});
		} else group_by_warp = false;
#endif
		if (!group_by_warp)
		parallel_for<C_group>(c, (size_t)Rc + 1, PV_LAMBDA(size_t rr_) {
			if (rr_ >= Rc) { recCnt_[rr_] = 0; grpCnt_[rr_] = 0; return; }
			u32 r = (u32)rr_;
			u32 b = candOff_[rv.rowBase(r)], e = candOff_[rv.rowBase(r + 1)];
			u32 *a = cd.hsort + b; u32 nh = 0;
			for (u32 x = b; x < e; x++) {
				if (cd.cCls[x] >= 3) continue;
				u32 found = NONE;
				for (u32 t = 0; t < nh; t++) {
					u32 y = a[t];
					if (same_record(cd, x, y) && cd.cLen[y] == cd.cLen[x] && walk_equal(g.steps, cd.cStart[y], cd.cStart[x], cd.cLen[x])) { found = y; break; }
				}
				if (found == NONE) { a[nh++] = x; cd.cHead[x] = x; cd.gCount[x] = 1; }
				else { cd.cHead[x] = found; cd.gCount[found]++; }
			}
			for (u32 st = nh / 2; st-- > 0;) sift_down(g.steps, cd, a, st, nh);
			for (u32 end = nh; end-- > 1;) { u32 t = a[0]; a[0] = a[end]; a[end] = t; sift_down(g.steps, cd, a, 0, end); }
			u32 nrec = 0;
			for (u32 t = 0; t < nh; t++) if (t == 0 || !same_record(cd, a[t - 1], a[t])) nrec++;
			recCnt_[r] = nrec; grpCnt_[r] = nh;
			if (k == 0) cv.hasData[r] = cd.evald[r];
			else if (!cd.evald[r] && cv.hasData[r]) atomic_add(cv.errs + 0, 1u);   // std::map::at throws in the reference
		});
		exclusive_scan(c, recCnt.p, recOff.p, (size_t)Rc + 1, OpAddU32(), 0u);
		exclusive_scan(c, grpCnt.p, grpOff.p, (size_t)Rc + 1, OpAddU32(), 0u);
		u32 nRec, nGrp, nPin;
		{ MultiRead mr(c); mr.add(&nRec, recOff.p + Rc, 4); mr.add(&nGrp, grpOff.p + Rc, 4); mr.add(&nPin, pinCount.p, 4); mr.wait(); }
		if (nPin) {   // pin pairs sit in candidate slots of class 4 (counted by C_pair_fill): pack them, in slot order
			Buf<u32> pf(c, nCand), pidx;
			u32 *pf_ = pf.p;
			parallel_for<S_pinflag>(c, nCand, PV_LAMBDA(size_t x) { pf_[x] = cd.cCls[x] == 4 ? 1u : 0u; });
			const u32 np = compact_flags(c, pf.p, nCand, pidx);
			if (np) {
				pins.pieces.emplace_back();
				PinStore::Piece &pc = pins.pieces.back();
				pc.n = np; pc.keyR = Buf<u32>(c, np); pc.keyJ = Buf<u32>(c, np); pc.keyK = Buf<u32>(c, np); pc.keySlot = Buf<u32>(c, np); pc.refStep = Buf<u32>(c, np); pc.altStep = Buf<u32>(c, np);
				u32 *kR = pc.keyR.p, *kJ = pc.keyJ.p, *kK = pc.keyK.p, *kS = pc.keySlot.p, *rS = pc.refStep.p, *aS = pc.altStep.p; const u32 *pidx_ = pidx.p;
				parallel_for<S_pinpack>(c, np, PV_LAMBDA(size_t q) {
					u32 x = pidx_[q];
					u32 row = ub_u32(candOff_, 0, nRows + 1, x) - 1; u32 r = rv.row2rov(row);
					kR[q] = rv.region(r); kJ[q] = rv.instJ ? rv.instJ[r] : 0u; kK[q] = k; kS[q] = x;
					rS[q] = cd.cRS[x]; aS[q] = cd.cStart[x];
				});
				pins.total += np;
			}
		}
		if (nRec == 0) continue;
		// ---- C6b: records + groups (offsets already global: the host only appends)
		const u32 grpBase = (u32)out.nGrpTot; const u64 gtBase = out.nGtTot;
		Buf<u32> oLoc(c, nRec); u32 *oLoc_ = oLoc.p;
		Buf<u32> oRov(c, nRec), oRef(c, nRec), oL(c, nRec), oR(c, nRec), oVt(c, nRec), oPos(c, nRec), oRS(c, nRec), oRL(c, nRec), oGrp(c, (size_t)nRec + 1);
		Buf<u32> gHap(c, nGrp), gStart(c, nGrp), gLen(c, nGrp), gCnt(c, nGrp), headRec(c, nCand), headOrd(c, nCand);
		Buf<u64> gtOff(c, (size_t)nRec + 1), recNh(c, (size_t)nRec + 1);
		u32 *oRov_ = oRov.p, *oRef_ = oRef.p, *oL_ = oL.p, *oR_ = oR.p, *oVt_ = oVt.p, *oPos_ = oPos.p, *oRS_ = oRS.p, *oRL_ = oRL.p, *oGrp_ = oGrp.p;
		u32 *gHap_ = gHap.p, *gStart_ = gStart.p, *gLen_ = gLen.p, *gCnt_ = gCnt.p, *headRec_ = headRec.p, *headOrd_ = headOrd.p;
		u64 *gtOff_ = gtOff.p, *recNh_ = recNh.p; const u32 *recOff_ = recOff.p, *grpOff_ = grpOff.p;
		parallel_for<C_emit>(c, Rc, PV_LAMBDA(size_t rr_) {
			u32 r = (u32)rr_;
			u32 nh = grpOff_[r + 1] - grpOff_[r];
			if (!nh) return;
			const u32 *a = cd.hsort + candOff_[rv.rowBase(r)];
			u32 rec = recOff_[r], gi = grpOff_[r]; u32 ord = 0;
			u32 nhaps = rv.rowBase(r + 1) - rv.rowBase(r);
			for (u32 t = 0; t < nh; t++) {
				u32 y = a[t];
				if (t == 0 || !same_record(cd, a[t - 1], y)) {
					if (t) rec++;
					ord = 0;
					u32 cls = cd.cCls[y];
					oRov_[rec] = rv.region(r); oLoc_[rec] = r; oRef_[rec] = refH; oL_[rec] = cd.cL[y]; oR_[rec] = cd.cR[y];
					oVt_[rec] = cls == 0 ? (u32)VT_INS : (cls == 1 ? (u32)VT_SUB : (u32)VT_DEL);
					oRS_[rec] = cd.cRS[y]; oRL_[rec] = cd.cRL[y];
					u32 l = g.loci[cd.cRS[y] + 1];                       // comp_pos (allele.hpp:161-179)
					oPos_[rec] = cls == 1 ? l : l - 1;
					oGrp_[rec] = grpBase + gi; recNh_[rec] = nhaps;
				}
				ord++;
				headRec_[y] = rec; headOrd_[y] = ord;
				u32 hh = g.cpList[g.cpOff[rv.comp(r)] + cd.cHap[y]];
				gHap_[gi] = hh; gStart_[gi] = cd.cStart[y] - g.stepOff[hh]; gLen_[gi] = cd.cLen[y]; gCnt_[gi] = cd.gCount[y];
				gi++;
			}
		});
		c.memset0(recNh.p + nRec, 8);
		exclusive_scan(c, recNh.p, gtOff.p, (size_t)nRec + 1, OpAddU64(), (u64)0);
		u64 nGt;
		if (uniformNh) nGt = (u64)nRec * uniformNh; else c.d2h(&nGt, gtOff.p + nRec, 8);   // every matrix has the same number of rows: no need to ask the device
		Buf<u16> gt(c, nGt ? nGt : 1);
		c.memset_ff(gt.p, nGt * 2);
		u16 *gt_ = gt.p;
		tr.mark("C6 group+emit");
		// ---- C7: genotypes.  Paths equal to the ref over the context columns with an identical slice -> 0 (overlay.cpp:534-580)
		// (one thread per RECORD walking its rows was measured slower at 16 paths: 256 vs 180 us -- too few threads to hide the loads)
		parallel_for<C_gt_ref>(c, nGt, PV_LAMBDA(size_t x) {
			u32 rec = uniformNh ? (u32)(x / uniformNh) : (u32)(ub_u64(gtOff_, 0, (u64)nRec + 1, (u64)x) - 1);
			u32 i = (u32)(x - gtOff_[rec]);
			u32 r = oLoc_[rec]; u32 J = rv.J(r);
			u32 lv = oL_[rec] >> 1, rvx = oR_[rec] >> 1;
			u32 ju = rv.find(r, lv), jv = rv.find(r, rvx);
			if (ju > jv) { u32 t = ju; ju = jv; jv = t; }
			const u16 *rr = cv.mat + rv.matBase(r) + (u64)refLocal * J, *ar = cv.mat + rv.matBase(r) + (u64)i * J;
			for (u32 j = ju; j <= jv; j++) if (rr[j] != ar[j]) return;
			u32 row = rv.rowBase(r) + i;
			if (cv.lapStart[row] == NONE) return;
			u32 as, an;
			if (!slice_bounds(g, lv, rvx, cv.lap(row), as, an)) return;
			if (an != oRL_[rec] || !walk_equal(g.steps, oRS_[rec], as, an)) return;
			gt_[x] = 0;
		});
		parallel_for<C_gt_alt>(c, nCand, PV_LAMBDA(size_t x) {
			if (cd.cCls[x] >= 3) return;
			u32 hd = cd.cHead[x]; u32 rec = headRec_[hd];
			gt_[gtOff_[rec] + cd.cHap[x]] = (u16)headOrd_[hd];
		});
		// ref-slice starts relative to the ref path, global gt offsets, the group sentinel
		parallel_for<C_keep>(c, (size_t)nRec + 1, PV_LAMBDA(size_t i) { if (i < nRec) oRS_[i] -= refStep0; else oGrp_[i] = grpBase + nGrp; gtOff_[i] += gtBase; });
		tr.mark("C7 genotypes");
		// ---- the records of this (chunk, ref) stay in HBM (Records::dev); Records::to_host copies them out on demand
		RecPiece pc;
		pc.rov = std::move(oRov); pc.ref = std::move(oRef); pc.cxtL = std::move(oL); pc.cxtR = std::move(oR); pc.vt = std::move(oVt); pc.pos = std::move(oPos);
		pc.refStart = std::move(oRS); pc.refLen = std::move(oRL); pc.grpOff = std::move(oGrp); pc.gHap = std::move(gHap); pc.gStart = std::move(gStart); pc.gLen = std::move(gLen);
		pc.gCnt = std::move(gCnt); pc.gtOff = std::move(gtOff); pc.gt = std::move(gt); pc.nRec = nRec; pc.nGrp = nGrp; pc.nGt = nGt;
		out.add_piece(std::move(pc));
		tr.mark("records d2h");
	}
}

PV_KERNEL(S_word, 8);
PV_KERNEL(S_item, 32);
PV_KERNEL(S_extend, 64);
PV_KERNEL(S_tree, 64);
PV_KERNEL(S_gt, 8);

// state of one slice tree (ist::slice_tree, /root/reference/include/ita/graph/slice_tree.hpp:106-848) in flat arenas
struct SliceTree {
	// vertices, keyed by ref start
	u32 *vStart, *vLeft, *vRight, *vAltHead, *vAltTail, *vKeyHead, *vLenHead; u32 nV;
	u32 *hKey, *hVal; u32 hMask;                 // ref start -> vertex
	u32 *aHap, *aStart, *aLen, *aNext, *aLive; u32 nA;   // alts of a vertex, in insertion order
	u32 *kHap, *kNext; u32 nK;                   // alt-hap keys present in a vertex's `alts` map (they survive removals)
	u32 *lLen, *lHap, *lNext, *lLive; u32 nL;    // (len, hap) members of a vertex's `len_haps`
	u32 root;
	PV_DEV u32 lookup(u32 start) const { u32 i = (u32)splitmix64(start) & hMask; while (true) { u32 k = hKey[i]; if (k == start) return hVal[i]; if (k == NONE) return NONE; i = (i + 1) & hMask; } }
	PV_DEV u32 new_vertex(u32 start) {
		u32 v = nV++; vStart[v] = start; vLeft[v] = NONE; vRight[v] = NONE; vAltHead[v] = NONE; vAltTail[v] = NONE; vKeyHead[v] = NONE; vLenHead[v] = NONE;
		u32 i = (u32)splitmix64(start) & hMask; while (hKey[i] != NONE) i = (i + 1) & hMask; hKey[i] = start; hVal[i] = v;
		return v;
	}
	PV_DEV bool has_hap(u32 v, u32 h) const { for (u32 k = vKeyHead[v]; k != NONE; k = kNext[k]) if (kHap[k] == h) return true; return false; }
	PV_DEV void add_key(u32 v, u32 h) { u32 k = nK++; kHap[k] = h; kNext[k] = vKeyHead[v]; vKeyHead[v] = k; }
	PV_DEV void push_alt(u32 v, u32 h, u32 st, u32 len) {
		u32 a = nA++; aHap[a] = h; aStart[a] = st; aLen[a] = len; aNext[a] = NONE; aLive[a] = 1;
		if (vAltTail[v] == NONE) vAltHead[v] = a; else aNext[vAltTail[v]] = a;
		vAltTail[v] = a;
	}
	PV_DEV void len_insert(u32 v, u32 len, u32 h) {
		for (u32 l = vLenHead[v]; l != NONE; l = lNext[l]) if (lLive[l] && lLen[l] == len && lHap[l] == h) return;
		u32 l = nL++; lLen[l] = len; lHap[l] = h; lLive[l] = 1; lNext[l] = vLenHead[v]; vLenHead[v] = l;
	}
	PV_DEV void len_erase(u32 v, u32 len, u32 h) { for (u32 l = vLenHead[v]; l != NONE; l = lNext[l]) if (lLive[l] && lLen[l] == len && lHap[l] == h) lLive[l] = 0; }
	PV_DEV u32 nth_alt(u32 v, u32 h, u32 idx) const { u32 n = 0; for (u32 a = vAltHead[v]; a != NONE; a = aNext[a]) if (aLive[a] && aHap[a] == h) { if (n == idx) return a; n++; } return NONE; }
	// vertex::insert_alt (slice_tree.hpp:347-356): std::map::insert does nothing when the hap key already exists
	PV_DEV void insert_alt(u32 v, u32 h, u32 st, u32 len) { if (!has_hap(v, h)) { add_key(v, h); push_alt(v, h, st, len); } len_insert(v, len, h); }
	PV_DEV void append_alt(u32 v, u32 h, u32 st, u32 len) { if (!has_hap(v, h)) add_key(v, h); push_alt(v, h, st, len); len_insert(v, len, h); }
	PV_DEV void remove_alt(u32 v, u32 h, u32 idx) { u32 a = nth_alt(v, h, idx); aLive[a] = 0; len_erase(v, aLen[a], h); }
	// vertex::check_update_type (slice_tree.hpp:284-345): returns comp_type, alt index in *ai
	PV_DEV u32 check(u32 v, u32 refStart, u32 h, u32 st, u32 len, u32 *ai) const {
		enum { NO_OVERLAP = 0, EXISTS = 1, REPLACE_ALT = 2, MERGE_EXTEND = 3, MERGE_REPLACE = 4, INSERT_ALT = 5, CONTAINED = 6, EXTEND_ALT = 7 };
		bool same = vStart[v] == refStart;
		if (same && !has_hap(v, h)) return INSERT_ALT;
		u32 n = 0;
		for (u32 a = vAltHead[v]; a != NONE; a = aNext[a]) {
			if (!aLive[a] || aHap[a] != h) continue;
			u32 as = aStart[a], al = aLen[a]; u32 ct = NONE;
			if (as == st && al == len) ct = EXISTS;
			else if (as < st && as + al >= st + len) ct = CONTAINED;
			else if (st < as && st + len >= as + al) ct = REPLACE_ALT;
			else if (as < st && as + al >= st && st + len >= as + al) ct = MERGE_EXTEND;
			else if (st < as && st + len >= as && as + al >= st + len) ct = MERGE_REPLACE;
			if (ct != NONE) { *ai = n; return ct; }
			n++;
		}
		return same ? EXTEND_ALT : NO_OVERLAP;
	}
	// slice_tree::add_vertex (slice_tree.hpp:830-846).  Returns false where the reference exits.
	PV_DEV bool add(u32 refStart, u32 h, u32 st, u32 len) {
		enum { NO_OVERLAP = 0, EXISTS = 1, REPLACE_ALT = 2, MERGE_EXTEND = 3, MERGE_REPLACE = 4, INSERT_ALT = 5, CONTAINED = 6, EXTEND_ALT = 7 };
		while (true) {
			if (nV == 0) { u32 v = new_vertex(refStart); root = v; insert_alt(v, h, st, len); return true; }
			u32 ct, ai = NONE, target;
			u32 ex = lookup(refStart);
			if (ex != NONE) {   // handle_existing_start
				target = ex; ct = check(ex, refStart, h, st, len, &ai);
				if (ct == NO_OVERLAP) return false;
			} else {            // find_insert_point_traverse
				u32 cur = root, ctr = 0; ct = NONE; target = NONE;
				while (true) {
					if (ctr++ >= nV) return false;
					if (has_hap(cur, h)) {
						u32 t = check(cur, refStart, h, st, len, &ai);
						if (t == EXTEND_ALT) return false;
						if (t != NO_OVERLAP) { ct = t; target = cur; break; }
					}
					// right() returns left_ in the reference (slice_tree.hpp:171-174): both directions test and follow the LEFT link
					if (vLeft[cur] == NONE) {
						u32 v = new_vertex(refStart);
						if (refStart < vStart[cur]) vLeft[cur] = v; else vRight[cur] = v;
						insert_alt(v, h, st, len);
						return true;
					}
					cur = vLeft[cur];
				}
			}
			switch (ct) {
			case EXISTS: case CONTAINED: return true;
			case INSERT_ALT: insert_alt(target, h, st, len); return true;
			case EXTEND_ALT: append_alt(target, h, st, len); return true;
			case REPLACE_ALT: remove_alt(target, h, ai); insert_alt(target, h, st, len); return true;
			case MERGE_EXTEND: {
				u32 a = nth_alt(target, h, ai); u32 oldLen = aLen[a]; u32 newLen = (st + len) - aStart[a];
				aStart[a] = st; aLen[a] = newLen; len_erase(target, oldLen, h); len_insert(target, newLen, h);
				return true;
			}
			case MERGE_REPLACE: {
				u32 a = nth_alt(target, h, ai); u32 as = aStart[a], al = aLen[a];
				remove_alt(target, h, ai);
				len = (as + al) - st;      // then add_vertex(ref_h_start, alt_h_idx, alt_h_start, new_len) again
				continue;
			}
			default: return false;
			}
		}
	}
};

// segmented running maximum over (segment key << 32 | value) words whose keys are sorted: a new key restarts the maximum
struct SegMax { PV_HD u64 operator()(u64 a, u64 b) const { return (a >> 32) == (b >> 32) ? (a > b ? a : b) : b; } };

PV_DEV u32 half_diff_ceil(u32 a, u32 b) { u32 d = (a > b ? a : b) - (a < b ? a : b); return (d + 1) / 2; }

// Seed-and-extend of the inversion pin pairs into maximal reverse-complement matches, the per-ref slice tree and the
// SUBR records (ise::sne /root/reference/src/ita/variation/sne.cpp:32-454, gen_inv_recs src/ita/genomics/vcf.cpp:123-239).
// Chains, limits and extensions are data-parallel; the slice tree is insertion-order dependent (quirk Q20) and is
// replayed by one thread per ref path.
inline void sne_pass(Ctx &c, const DevGraph &G, const GraphView &g, const CallIn &in, PinStore &pins, u32 R, u32 *errs, Records &out, Trace &tr) {
	const u32 nP = (u32)pins.total, nRefs = in.nRefs;
	Buf<u32> kR(c, nP), kJ(c, nP), kK(c, nP), kS(c, nP), rS(c, nP), aS(c, nP);
	{
		size_t o = 0;
		for (auto &pc : pins.pieces) {
			c.d2d(kR.p + o, pc.keyR.p, (size_t)pc.n * 4); c.d2d(kJ.p + o, pc.keyJ.p, (size_t)pc.n * 4); c.d2d(kK.p + o, pc.keyK.p, (size_t)pc.n * 4);
			c.d2d(kS.p + o, pc.keySlot.p, (size_t)pc.n * 4); c.d2d(rS.p + o, pc.refStep.p, (size_t)pc.n * 4); c.d2d(aS.p + o, pc.altStep.p, (size_t)pc.n * 4);
			o += pc.n;
		}
	}
	// the reference's insertion order: region, unrolled matrix, ref, (path, context) = slot
	Buf<u32> perm(c, nP), wk(c, nP);
	u32 *perm_ = perm.p, *wk_ = wk.p;
	parallel_for<S_word>(c, nP, PV_LAMBDA(size_t i) { perm_[i] = (u32)i; });
	const u32 *words[4] = {kS.p, kK.p, kJ.p, kR.p};
	for (int w = 0; w < 4; w++) {
		const u32 *src = words[w];
		parallel_for<S_word>(c, nP, PV_LAMBDA(size_t i) { wk_[i] = src[perm_[i]]; });
		radix_sort_pairs_inplace(c, wk.p, perm.p, nP, 32);
	}
	// work items: pin pair o seen from ref ordinal kk (a pair between two ref paths belongs to both trees, sne.hpp:128-147)
	const u32 *refs_ = nullptr;
	Buf<u32> refsD(c, nRefs); c.h2d(refsD.p, in.refs, (size_t)nRefs * 4); refs_ = refsD.p;
	const size_t nF = (size_t)nP * nRefs;
	Buf<u32> fl(c, nF), idx;
	u32 *fl_ = fl.p; const u32 *rS_ = rS.p, *aS_ = aS.p;
	parallel_for<S_item>(c, nF, PV_LAMBDA(size_t x) {
		u32 kk = (u32)(x / nP), o = (u32)(x % nP); u32 p = perm_[o];
		u32 hr = ub_u32(g.stepOff, 0, g.H + 1, rS_[p]) - 1, ha = ub_u32(g.stepOff, 0, g.H + 1, aS_[p]) - 1;
		fl_[x] = (hr == refs_[kk] || ha == refs_[kk]) ? 1u : 0u;
	});
	const u32 nItems = compact_flags(c, fl.p, nF, idx);
	if (!nItems) return;
	Buf<u32> iK(c, nItems), iAlt(c, nItems), iRef(c, nItems), iAltIdx(c, nItems), iperm(c, nItems), iw(c, nItems);
	u32 *iK_ = iK.p, *iAlt_ = iAlt.p, *iRef_ = iRef.p, *iAltIdx_ = iAltIdx.p, *iperm_ = iperm.p, *iw_ = iw.p; const u32 *idx_ = idx.p;
	parallel_for<S_item>(c, nItems, PV_LAMBDA(size_t t) {
		u32 x = idx_[t]; u32 kk = x / nP, o = x % nP; u32 p = perm_[o];
		u32 hr = ub_u32(g.stepOff, 0, g.H + 1, rS_[p]) - 1, ha = ub_u32(g.stepOff, 0, g.H + 1, aS_[p]) - 1;
		u32 ri = rS_[p] - g.stepOff[hr], ai = aS_[p] - g.stepOff[ha];
		iK_[t] = kk; iperm_[t] = (u32)t;
		if (hr == refs_[kk]) { iAlt_[t] = ha; iRef_[t] = ri; iAltIdx_[t] = ai; }
		else { iAlt_[t] = hr; iRef_[t] = ai; iAltIdx_[t] = ri; }
	});
	// order (ref ordinal, alt path, insertion order): stable sorts by alt path, then by ref ordinal
	parallel_for<S_word>(c, nItems, PV_LAMBDA(size_t t) { iw_[t] = iAlt_[iperm_[t]]; });
	radix_sort_pairs_inplace(c, iw.p, iperm.p, nItems, bits_for(g.H));
	parallel_for<S_word>(c, nItems, PV_LAMBDA(size_t t) { iw_[t] = iK_[iperm_[t]]; });
	radix_sort_pairs_inplace(c, iw.p, iperm.p, nItems, bits_for(nRefs));
	// chain limits (sne.cpp:92-122) + extension of every link (sne.cpp:124-213)
	Buf<u32> eOk(c, nItems), eRefStart(c, nItems), eAltStart(c, nItems), eLen(c, nItems), eK(c, nItems), eAlt(c, nItems);
	u32 *eOk_ = eOk.p, *eRefStart_ = eRefStart.p, *eAltStart_ = eAltStart.p, *eLen_ = eLen.p, *eK_ = eK.p, *eAlt_ = eAlt.p;
	parallel_for<S_extend>(c, nItems, PV_LAMBDA(size_t t) {
		u32 me = iperm_[t]; u32 kk = iK_[me], alt = iAlt_[me];
		bool hasPrev = t > 0 && iK_[iperm_[t - 1]] == kk && iAlt_[iperm_[t - 1]] == alt;
		bool hasNext = t + 1 < nItems && iK_[iperm_[t + 1]] == kk && iAlt_[iperm_[t + 1]] == alt;
		u32 cur = iRef_[me];
		u32 limL = NONE, limR = NONE;
		if (!hasPrev && hasNext) limR = half_diff_ceil(cur, iRef_[iperm_[t + 1]]);
		else if (!hasNext && hasPrev) limL = half_diff_ceil(cur, iRef_[iperm_[t - 1]]);
		else if (hasPrev && hasNext) { limL = half_diff_ceil(cur, iRef_[iperm_[t - 1]]); limR = half_diff_ceil(cur, iRef_[iperm_[t + 1]]); }
		const u32 refH = refs_[kk];
		const u32 rb = g.stepOff[refH], rn = g.stepOff[refH + 1] - rb, ab = g.stepOff[alt], an = g.stepOff[alt + 1] - ab;
		const u32 f0 = cur, r0 = iAltIdx_[me];
		// do_match (sne.cpp:32-57); indices past the end of a walk are undefined behaviour in the reference -> mismatch
		auto match = [&](u32 i, u32 j) -> int {
			if (i == NONE || j == NONE) return -1;
			if (i >= rn || j >= an) return 1;
			u32 x = g.steps[rb + i], y = g.steps[ab + j];
			return ((x >> 1) == (y >> 1) && (x & 1u) != (y & 1u)) ? 0 : 1;
		};
		u32 S_i = limL == NONE ? f0 : f0 - limL - 1;
		u32 S_j = limR == NONE ? r0 : r0 - limL - 1;      // (sic) tests the right limit, subtracts the left one (sne.cpp:170-172)
		u32 E_j = limR == NONE ? NONE : r0 + limR + 1;
		u32 E_i = limR == NONE ? NONE : f0 + limR + 1;
		u32 i = f0, j = r0, da = 0, db = 0;
		for (; i < E_i || j > S_j; ++i, j--) { if (match(i, j) != 0) break; da++; }
		i = f0; j = r0;
		for (; i > S_i || j < E_j; --i, ++j) { if (match(i, j) != 0) break; db++; }
		u32 x = db == 0 ? f0 : (f0 - db) + 1;
		u32 y = db == 0 ? r0 : r0 + (db - 1);
		u32 z = da + db;
		if (db > 0 && da > 0) z--;
		eK_[t] = kk; eAlt_[t] = alt;
		if (da == 0 && db == 0) { eOk_[t] = 0; return; }
		eOk_[t] = 1; eRefStart_[t] = x; eLen_[t] = z; eAltStart_[t] = y - z + 1;
	});
	std::vector<u32> hK = eK.to_host();
	tr.mark("S chains+extend");
	// one slice tree per ref path, items in (alt path, insertion) order
	u32 t0 = 0;
	for (u32 kk = 0; kk < nRefs; kk++) {
		u32 t1 = t0; while (t1 < nItems && hK[t1] == kk) t1++;
		const u32 n = t1 - t0;
		if (!n) continue;
		Buf<u32> oRec(c, (size_t)n * 12), oSl(c, (size_t)n * 2), oHap(c, (size_t)n * 2), oCnt(c, 4);
		u32 *oRec_ = oRec.p, *oSl_ = oSl.p, *oHap_ = oHap.p, *oCnt_ = oCnt.p;
		const u32 refH = in.refs[kk];
		// ---- conflict-free fast path.  Two slices interact in the slice tree only when they belong to the same alt path and
		// their ALT intervals overlap, touch or coincide (vertex::check_update_type, slice_tree.hpp:284-345); whether such a pair
		// meets at all depends on the tree's shape (quirk Q20: the traversal only follows left links), i.e. on the whole
		// insertion history.  Without such a pair every insertion is INSERT_ALT / EXTEND_ALT / EXISTS or a new vertex, the tree's
		// content is just "slices grouped by ref start" and the records follow from sorts: vertices by ref start, lengths
		// ascending, paths ascending, slices of a path in insertion order (vcf.cpp:147-215).  Any suspicious pair -> the
		// sequential replay below, unchanged.
		bool fast = false;
		{
			Buf<u32> okF(c, n), okIdx;
			u32 *okF_ = okF.p;
			parallel_for<S_word>(c, n, PV_LAMBDA(size_t i) { okF_[i] = eOk_[t0 + i] ? 1u : 0u; });
			const u32 m = compact_flags(c, okF.p, n, okIdx);
			if (m == 0) { t0 = t1; continue; }
			Buf<u32> perm2(c, m), wk2(c, m), dup(c, m), cflag(c, 1);
			Buf<u64> sv(c, m), sm(c, m);
			u32 *perm2_ = perm2.p, *wk2_ = wk2.p, *dup_ = dup.p, *cflag_ = cflag.p; u64 *sv_ = sv.p, *sm_ = sm.p; const u32 *okIdx_ = okIdx.p;
			cflag.zero();
			parallel_for<S_word>(c, m, PV_LAMBDA(size_t i) { perm2_[i] = t0 + okIdx_[i]; });
			const u32 *keys4[4] = {eRefStart_, eLen_, eAltStart_, eAlt_};   // LSD: sorted by (alt path, alt start, len, ref start), stable in insertion order
			for (int w = 0; w < 4; w++) {
				const u32 *src = keys4[w];
				parallel_for<S_word>(c, m, PV_LAMBDA(size_t i) { wk2_[i] = src[perm2_[i]]; });
				radix_sort_pairs_inplace(c, wk2.p, perm2.p, m, w == 3 ? bits_for(g.H) : 32);
			}
			parallel_for<S_item>(c, m, PV_LAMBDA(size_t i) {
				u32 t = perm2_[i]; bool d = false;
				if (i > 0) { u32 p = perm2_[i - 1]; d = eAlt_[p] == eAlt_[t] && eAltStart_[p] == eAltStart_[t] && eLen_[p] == eLen_[t] && eRefStart_[p] == eRefStart_[t]; }
				dup_[i] = d ? 1u : 0u;       // the same slice at the same ref start again: EXISTS, dropped (the first one stays)
				sv_[i] = ((u64)eAlt_[t] << 32) | (d ? 0ull : (u64)(eAltStart_[t] + eLen_[t]));
			});
			inclusive_scan(c, sv.p, sm.p, m, SegMax(), (u64)0);
			parallel_for<S_item>(c, m, PV_LAMBDA(size_t i) {
				if (i == 0 || dup_[i]) return;
				u32 t = perm2_[i]; u64 pm = sm_[i - 1];
				if ((u32)(pm >> 32) == eAlt_[t] && (u32)pm >= eAltStart_[t]) cflag_[0] = 1;   // overlaps / touches / coincides with an earlier slice of the path
			});
			if (cflag.read(0) == 0) {
				fast = true;
				// kept slices in insertion order, then sorted by (ref start, len, path) -- stable, so a path's slices keep their order
				Buf<u32> keepI(c, n), kIdx;
				u32 *keepI_ = keepI.p;
				keepI.zero();
				parallel_for<S_word>(c, m, PV_LAMBDA(size_t i) { if (!dup_[i]) keepI_[perm2_[i] - t0] = 1; });
				const u32 kN = compact_flags(c, keepI.p, n, kIdx);
				Buf<u32> perm3(c, kN), wk3(c, kN), nr(c, (size_t)kN + 1), nh(c, (size_t)kN + 1), nrS(c, (size_t)kN + 1), nhS(c, (size_t)kN + 1);
				u32 *perm3_ = perm3.p, *wk3_ = wk3.p, *nr_ = nr.p, *nh_ = nh.p; const u32 *kIdx_ = kIdx.p;
				parallel_for<S_word>(c, kN, PV_LAMBDA(size_t i) { perm3_[i] = t0 + kIdx_[i]; });
				const u32 *keys3[3] = {eAlt_, eLen_, eRefStart_};
				for (int w = 0; w < 3; w++) {
					const u32 *src = keys3[w];
					parallel_for<S_word>(c, kN, PV_LAMBDA(size_t i) { wk3_[i] = src[perm3_[i]]; });
					radix_sort_pairs_inplace(c, wk3.p, perm3.p, kN, w == 0 ? bits_for(g.H) : 32);
				}
				parallel_for<S_item>(c, (size_t)kN + 1, PV_LAMBDA(size_t i) {
					if (i >= kN) { nr_[i] = 0; nh_[i] = 0; return; }
					u32 t = perm3_[i]; bool r = i == 0, h = i == 0;
					if (i > 0) { u32 p = perm3_[i - 1]; r = eRefStart_[p] != eRefStart_[t] || eLen_[p] != eLen_[t]; h = r || eAlt_[p] != eAlt_[t]; }
					nr_[i] = r ? 1u : 0u; nh_[i] = h ? 1u : 0u;
				});
				exclusive_scan(c, nr.p, nrS.p, (size_t)kN + 1, OpAddU32(), 0u);
				exclusive_scan(c, nh.p, nhS.p, (size_t)kN + 1, OpAddU32(), 0u);
				const u32 *nrS_ = nrS.p, *nhS_ = nhS.p;
				parallel_for<S_item>(c, kN, PV_LAMBDA(size_t i) {
					u32 t = perm3_[i];
					oSl_[2 * i] = eAlt_[t]; oSl_[2 * i + 1] = eAltStart_[t];
					if (nh_[i]) oHap_[nhS_[i]] = eAlt_[t];
					if (nr_[i]) { u32 *o = oRec_ + 6 * (size_t)nrS_[i]; o[0] = eRefStart_[t]; o[1] = eLen_[t]; o[2] = (u32)i; o[4] = nhS_[i]; }
					if (i == kN - 1) { oCnt_[0] = nrS_[kN]; oCnt_[1] = kN; oCnt_[2] = nhS_[kN]; }
				});
				parallel_for<S_item>(c, kN, PV_LAMBDA(size_t i) {   // counts = distance to the next record's first slice / path
					if (!nr_[i]) return;
					u32 r = nrS_[i]; u32 *o = oRec_ + 6 * (size_t)r;
					u32 j = (u32)i + 1; while (j < kN && !nr_[j]) j++;
					o[3] = j - (u32)i; o[5] = (j < kN ? nhS_[j] : nhS_[kN]) - nhS_[i];
				});
			}
		}
		if (!fast) {
		u32 hcap = 4; while (hcap < 2 * n + 2) hcap <<= 1;
		Buf<u32> vA(c, (size_t)n * 7), hT(c, (size_t)hcap * 2), aA(c, (size_t)n * 5), kA(c, (size_t)n * 2), lA(c, (size_t)n * 8), ord(c, n);
		c.memset_ff(hT.p, (size_t)hcap * 4);
		u32 *vA_ = vA.p, *hT_ = hT.p, *aA_ = aA.p, *kA_ = kA.p, *lA_ = lA.p, *ord_ = ord.p;
		out.n_slice_tree_replays++;
		parallel_for<S_tree>(c, 1, PV_LAMBDA(size_t) {
			SliceTree T;
			T.vStart = vA_; T.vLeft = vA_ + n; T.vRight = vA_ + 2 * (size_t)n; T.vAltHead = vA_ + 3 * (size_t)n; T.vAltTail = vA_ + 4 * (size_t)n; T.vKeyHead = vA_ + 5 * (size_t)n; T.vLenHead = vA_ + 6 * (size_t)n; T.nV = 0;
			T.hKey = hT_; T.hVal = hT_ + hcap; T.hMask = hcap - 1;
			T.aHap = aA_; T.aStart = aA_ + n; T.aLen = aA_ + 2 * (size_t)n; T.aNext = aA_ + 3 * (size_t)n; T.aLive = aA_ + 4 * (size_t)n; T.nA = 0;
			T.kHap = kA_; T.kNext = kA_ + n; T.nK = 0;
			T.lLen = lA_; T.lHap = lA_ + 2 * (size_t)n; T.lNext = lA_ + 4 * (size_t)n; T.lLive = lA_ + 6 * (size_t)n; T.nL = 0;
			T.root = NONE;
			for (u32 t = t0; t < t1; t++) {
				if (!eOk_[t]) continue;
				if (!T.add(eRefStart_[t], eAlt_[t], eAltStart_[t], eLen_[t])) { atomic_add(errs + 0, 1u); break; }
			}
			// records: vertices by ascending ref start (std::map order), per vertex the lengths ascending (vcf.cpp:147-150)
			u32 nv = T.nV;
			for (u32 v = 0; v < nv; v++) ord_[v] = v;
			{   // heap sort by ref start (starts are distinct: one vertex per start)
				auto sift = [&](u32 root, u32 end) { while (2 * root + 1 < end) { u32 ch = 2 * root + 1; if (ch + 1 < end && T.vStart[ord_[ch]] < T.vStart[ord_[ch + 1]]) ch++; if (T.vStart[ord_[root]] >= T.vStart[ord_[ch]]) return; u32 t = ord_[root]; ord_[root] = ord_[ch]; ord_[ch] = t; root = ch; } };
				for (u32 st = nv / 2; st-- > 0;) sift(st, nv);
				for (u32 end = nv; end-- > 1;) { u32 t = ord_[0]; ord_[0] = ord_[end]; ord_[end] = t; sift(0, end); }
			}
			u32 nRec = 0, nSl = 0, nHp = 0;
			for (u32 q = 0; q < nv; q++) {
				u32 v = ord_[q];
				u32 lastLen = 0; bool first = true;
				while (true) {   // next distinct live length above lastLen
					u32 best = NONE;
					for (u32 l = T.vLenHead[v]; l != NONE; l = T.lNext[l]) if (T.lLive[l] && (first || T.lLen[l] > lastLen) && T.lLen[l] < best) best = T.lLen[l];
					if (best == NONE) break;
					first = false; lastLen = best;
					// alt set: haps of the length class ascending, their alts of that length in insertion order
					u32 sl0 = nSl, hp0 = nHp;
					u32 lastHap = 0; bool firstH = true;
					while (true) {
						u32 bh = NONE;
						for (u32 l = T.vLenHead[v]; l != NONE; l = T.lNext[l]) if (T.lLive[l] && T.lLen[l] == best && (firstH || T.lHap[l] > lastHap) && T.lHap[l] < bh) bh = T.lHap[l];
						if (bh == NONE) break;
						firstH = false; lastHap = bh;
						oHap_[nHp++] = bh;
						for (u32 a = T.vAltHead[v]; a != NONE; a = T.aNext[a]) if (T.aLive[a] && T.aHap[a] == bh && T.aLen[a] == best) { oSl_[2 * (size_t)nSl] = bh; oSl_[2 * (size_t)nSl + 1] = T.aStart[a]; nSl++; }
					}
					if (nSl == sl0) { nHp = hp0; continue; }   // no alts at that length: no record (vcf.cpp:214-215)
					u32 *o = oRec_ + 6 * (size_t)nRec;
					o[0] = T.vStart[v]; o[1] = best; o[2] = sl0; o[3] = nSl - sl0; o[4] = hp0; o[5] = nHp - hp0;
					nRec++;
				}
			}
			oCnt_[0] = nRec; oCnt_[1] = nSl; oCnt_[2] = nHp;
		});
		c.sync();   // (the replay's scratch dies with this block)
		}
		u32 hc[4]; c.d2h(hc, oCnt.p, 16);
		const u32 nRec = hc[0], nSl = hc[1];
		t0 = t1;
		if (!nRec) continue;
		// SUBR records of this ref: one pseudo-region after the real ones carries the component (for the genotype columns)
		const u32 comp = G.pathComp[refH]; const u32 nh = G.cpOff[comp + 1] - G.cpOff[comp];
		const u32 pseudo = (u32)out.rov_fl.size();
		out.rov_fl.push_back(NONE); out.rov_comp.push_back(comp);
		const u32 grpBase = (u32)out.nGrpTot; const u64 gtBase = out.nGtTot;
		Buf<u32> oRov(c, nRec), oRef(c, nRec), oL(c, nRec), oR2(c, nRec), oVt(c, nRec), oPos(c, nRec), oRS(c, nRec), oRL(c, nRec), oGrp(c, (size_t)nRec + 1);
		Buf<u32> gHap(c, nRec), gStart(c, nRec), gLen(c, nRec), gCnt(c, nRec);
		Buf<u64> gtOff(c, (size_t)nRec + 1);
		Buf<u16> gt(c, (size_t)nRec * nh);
		c.memset_ff(gt.p, (size_t)nRec * nh * 2);
		u32 *oRov_ = oRov.p, *oRef_ = oRef.p, *oL_ = oL.p, *oR2_ = oR2.p, *oVt_ = oVt.p, *oPos_ = oPos.p, *oRS_ = oRS.p, *oRL_ = oRL.p, *oGrp_ = oGrp.p;
		u32 *gHap_ = gHap.p, *gStart_ = gStart.p, *gLen_ = gLen.p, *gCnt_ = gCnt.p; u64 *gtOff_ = gtOff.p; u16 *gt_ = gt.p;
		const u32 rb = G.h_stepOff[refH], rn = G.h_stepOff[refH + 1] - rb;
		parallel_for<S_gt>(c, (size_t)nRec + 1, PV_LAMBDA(size_t q) {
			gtOff_[q] = gtBase + (u64)q * nh;
			if (q >= nRec) { oGrp_[q] = grpBase + nRec; return; }
			const u32 *o = oRec_ + 6 * q;
			u32 s = o[0], len = o[1];
			// a merged length is measured on the ALT walk (slice_tree.hpp:378-391) and can run past the end of the ref or
			// alt walk; the reference then reads beyond its arrays (undefined).  Clamp and count.
			u32 ah = oSl_[2 * (size_t)o[2]], as0 = oSl_[2 * (size_t)o[2] + 1];
			u32 an = g.stepOff[ah + 1] - g.stepOff[ah];
			if (s >= rn || as0 >= an) { atomic_add(errs + 0, 1u); s = 0; as0 = 0; len = 1; }
			if (s + len > rn || s + len < s) { len = rn - s; atomic_add(errs + 4, 1u); }
			if (as0 + len > an || as0 + len < as0) { len = an - as0; atomic_add(errs + 4, 1u); }
			oRov_[q] = pseudo; oRef_[q] = refH; oVt_[q] = (u32)VT_SUBR; oRS_[q] = s; oRL_[q] = len; oGrp_[q] = grpBase + (u32)q;
			u32 a = g.steps[rb + s], b = g.steps[rb + s + len - 1];
			if ((a & 1u) && (b & 1u)) { a &= ~1u; b &= ~1u; }   // both reverse: printed forward (vcf.cpp:163-168)
			oL_[q] = a; oR2_[q] = b;
			if (s + 1 < rn) oPos_[q] = g.loci[rb + s + 1]; else { oPos_[q] = 0; atomic_add(errs + 4, 1u); }   // comp_pos reads loci[start + 1]
			gHap_[q] = ah; gStart_[q] = as0; gLen_[q] = len; gCnt_[q] = o[3];
			u16 *row = gt_ + q * nh;
			row[g.hapLocal[refH]] = 0;
			for (u32 x = 0; x < o[5]; x++) { u32 h = oHap_[o[4] + x]; if (g.pathComp[h] == comp) row[g.hapLocal[h]] = 1; }
		});
		RecPiece pc;
		pc.rov = std::move(oRov); pc.ref = std::move(oRef); pc.cxtL = std::move(oL); pc.cxtR = std::move(oR2); pc.vt = std::move(oVt); pc.pos = std::move(oPos);
		pc.refStart = std::move(oRS); pc.refLen = std::move(oRL); pc.grpOff = std::move(oGrp); pc.gHap = std::move(gHap); pc.gStart = std::move(gStart); pc.gLen = std::move(gLen);
		pc.gCnt = std::move(gCnt); pc.gtOff = std::move(gtOff); pc.gt = std::move(gt); pc.nRec = nRec; pc.nGrp = nRec; pc.nGt = (u64)nRec * nh;
		out.add_piece(std::move(pc));
		(void)nSl;
	}
	tr.mark("S slice trees");
}

PV_KERNEL(T_rows, 12);
PV_KERNEL(T_count, 24);
PV_KERNEL(T_fill, 24);
PV_KERNEL(T_laps, 24);
PV_KERNEL(T_sizes, 24);
PV_KERNEL(T_align, 64);
PV_KERNEL(T_mismatch, 16);
PV_KERNEL(T_inst, 24);
PV_KERNEL(T_matrix, 32);

// Tangled regions (a boundary column visited more than once): ita::untangle::untangle
// (/root/reference/src/ita/genomics/untangle.cpp:29-247) + ita::align (/root/reference/src/ita/align/align.cpp:71-253).
// Per tangled region: every path's race (all laps = runs of consecutive steps on region vertices), affine-gap global
// alignment of the first ref's lap sequence against every other path's (whole laps are the alphabet; u32 arithmetic
// with INVALID as +inf that wraps, quirk Q12), one unrolled depth matrix per alignment column that holds an 'X', then
// the ordinary overlay over those matrices with TANGLED = T (overlay.cpp:616-626).  Only the first ref's matrices are
// used (untangle.cpp:245-246, Q11).
inline void tangled_pass(Ctx &c, const DevGraph &G, const GraphView &g, const CallIn &in, const RovView &rv, const ChunkView &cv, u32 Rc,
			 const u32 *tangledFlags, u32 *errs, Records &out, Trace &tr, PinStore &pins) {
	Buf<u32> tList;
	const u32 nT = compact_flags(c, tangledFlags, Rc, tList);
	if (!nT) return;
	const u32 ref0Local = G.hapLocal[in.refs[0]];
	const u32 *tList_ = tList.p;
	// rows and columns of the tangled regions, flattened
	Buf<u32> tRows(c, (size_t)nT + 1), tRowPre(c, (size_t)nT + 1), tCols(c, (size_t)nT + 1), tColPre(c, (size_t)nT + 1);
	u32 *tRows_ = tRows.p, *tCols_ = tCols.p;
	parallel_for<T_rows>(c, (size_t)nT + 1, PV_LAMBDA(size_t t) {
		if (t >= nT) { tRows_[t] = 0; tCols_[t] = 0; return; }
		u32 r = tList_[t];
		tRows_[t] = rv.rowBase(r + 1) - rv.rowBase(r); tCols_[t] = rv.J(r);
	});
	exclusive_scan(c, tRows.p, tRowPre.p, (size_t)nT + 1, OpAddU32(), 0u);
	exclusive_scan(c, tCols.p, tColPre.p, (size_t)nT + 1, OpAddU32(), 0u);
	const u32 nTR = tRowPre.read(nT), nTC = tColPre.read(nT);
	const u32 *tRowPre_ = tRowPre.p, *tColPre_ = tColPre.p;
	// ---- races: steps of every (region, path) on region vertices, sorted, cut into laps
	Buf<u32> sCnt(c, (size_t)nTR + 1), sOff(c, (size_t)nTR + 1), cursor(c, nTR ? nTR : 1);
	sCnt.zero(); cursor.zero();
	u32 *sCnt_ = sCnt.p, *cursor_ = cursor.p;
	parallel_for<T_count>(c, nTC, PV_LAMBDA(size_t x) {
		u32 t = ub_u32(tColPre_, 0, nT + 1, (u32)x) - 1; u32 r = tList_[t]; u32 j = (u32)x - tColPre_[t];
		u32 v = rv.vert(r, j); u32 comp = rv.comp(r);
		for (u32 e = g.vsOff[v], ee = g.vsOff[v + 1]; e < ee; e++) {
			u32 h = g.vsHap[e] >> 1;
			if (g.pathComp[h] == comp) atomic_add(sCnt_ + tRowPre_[t] + g.hapLocal[h], 1u);
		}
	});
	exclusive_scan(c, sCnt.p, sOff.p, (size_t)nTR + 1, OpAddU32(), 0u);
	const u32 TS = sOff.read(nTR);
	Buf<u32> sList(c, TS ? TS : 1), lapS(c, TS ? TS : 1), lapL(c, TS ? TS : 1), lapCnt(c, nTR ? nTR : 1);
	u32 *sList_ = sList.p, *lapS_ = lapS.p, *lapL_ = lapL.p, *lapCnt_ = lapCnt.p; const u32 *sOff_ = sOff.p;
	parallel_for<T_fill>(c, nTC, PV_LAMBDA(size_t x) {
		u32 t = ub_u32(tColPre_, 0, nT + 1, (u32)x) - 1; u32 r = tList_[t]; u32 j = (u32)x - tColPre_[t];
		u32 v = rv.vert(r, j); u32 comp = rv.comp(r);
		for (u32 e = g.vsOff[v], ee = g.vsOff[v + 1]; e < ee; e++) {
			u32 h = g.vsHap[e] >> 1;
			if (g.pathComp[h] != comp) continue;
			u32 row = tRowPre_[t] + g.hapLocal[h];
			sList_[sOff_[row] + atomic_add(cursor_ + row, 1u)] = g.vsStep[e];
		}
	});
	parallel_for<T_laps>(c, nTR, PV_LAMBDA(size_t row) {
		u32 *a = sList_ + sOff_[row]; u32 n = sOff_[row + 1] - sOff_[row];
		// heap sort ascending (the lists are filled in arbitrary order)
		for (u32 st = n / 2; st-- > 0;) { u32 root = st; while (2 * root + 1 < n) { u32 ch = 2 * root + 1; if (ch + 1 < n && a[ch] < a[ch + 1]) ch++; if (a[root] >= a[ch]) break; u32 tt = a[root]; a[root] = a[ch]; a[ch] = tt; root = ch; } }
		for (u32 end = n; end-- > 1;) { u32 tt = a[0]; a[0] = a[end]; a[end] = tt; u32 root = 0; while (2 * root + 1 < end) { u32 ch = 2 * root + 1; if (ch + 1 < end && a[ch] < a[ch + 1]) ch++; if (a[root] >= a[ch]) break; u32 t2 = a[root]; a[root] = a[ch]; a[ch] = t2; root = ch; } }
		u32 nl = 0, o = sOff_[row];
		for (u32 k = 0; k < n;) { u32 e = k + 1; while (e < n && a[e] == a[e - 1] + 1) e++; lapS_[o + nl] = a[k]; lapL_[o + nl] = e - k; nl++; k = e; }
		lapCnt_[row] = nl;
	});
	// ---- alignments of the first ref's laps against every other path's
	Buf<u64> dpSz(c, (size_t)nTR + 1), dpOff(c, (size_t)nTR + 1);
	Buf<u32> etCap(c, (size_t)nTR + 1), etOff(c, (size_t)nTR + 1), etLen(c, nTR ? nTR : 1);
	u64 *dpSz_ = dpSz.p; u32 *etCap_ = etCap.p, *etLen_ = etLen.p;
	parallel_for<T_sizes>(c, (size_t)nTR + 1, PV_LAMBDA(size_t row) {
		if (row >= nTR) { dpSz_[row] = 0; etCap_[row] = 0; return; }
		u32 t = ub_u32(tRowPre_, 0, nT + 1, (u32)row) - 1; u32 i = (u32)row - tRowPre_[t];
		u32 Lr = lapCnt_[tRowPre_[t] + ref0Local], Lh = lapCnt_[row];
		if (i == ref0Local) { dpSz_[row] = 0; etCap_[row] = 0; return; }
		dpSz_[row] = (Lr == 1 && Lh == 1) ? 0ull : (u64)(Lr + 1) * (Lh + 1);
		etCap_[row] = Lr + Lh + 1;
	});
	exclusive_scan(c, dpSz.p, dpOff.p, (size_t)nTR + 1, OpAddU64(), (u64)0);
	exclusive_scan(c, etCap.p, etOff.p, (size_t)nTR + 1, OpAddU32(), 0u);
	u64 totDP; c.d2h(&totDP, dpOff.p + nTR, 8);
	const u32 totEt = etOff.read(nTR);
	if (totDP > (1ull << 31)) throw std::runtime_error("call: tangled regions need more than 24 GB of alignment matrices");
	Buf<u32> dpM(c, totDP ? totDP : 1), dpI(c, totDP ? totDP : 1), dpD(c, totDP ? totDP : 1);
	Buf<u8> et(c, totEt ? totEt : 1);
	u32 *dpM_ = dpM.p, *dpI_ = dpI.p, *dpD_ = dpD.p; u8 *et_ = et.p; const u64 *dpOff_ = dpOff.p; const u32 *etOff_ = etOff.p;
	parallel_for<T_align>(c, nTR, PV_LAMBDA(size_t row) {
		u32 t = ub_u32(tRowPre_, 0, nT + 1, (u32)row) - 1; u32 i = (u32)row - tRowPre_[t];
		if (i == ref0Local) { etLen_[row] = 0; return; }
		u32 rrow = tRowPre_[t] + ref0Local;
		u32 Lr = lapCnt_[rrow], Lh = lapCnt_[row];
		const u32 *rS = lapS_ + sOff_[rrow], *rL = lapL_ + sOff_[rrow], *hS = lapS_ + sOff_[row], *hL = lapL_ + sOff_[row];
		u8 *e = et_ + etOff_[row];
		auto eq = [&](u32 a, u32 b) -> bool { return rL[a] == hL[b] && walk_equal(g.steps, rS[a], hS[b], rL[a]); };   // eq_at (align.cpp:37-51)
		if (Lr == 1 && Lh == 1) { e[0] = eq(0, 0) ? 'M' : 'X'; etLen_[row] = 1; return; }
		const u32 INV = NONE, ca = 0, cx = 1, co = 2, ce = 1;   // match, mismatch, gap open, gap extend (align.cpp:231)
		const u32 rows = Lr + 1, cols = Lh + 1;
		u32 *M = dpM_ + dpOff_[row], *I = dpI_ + dpOff_[row], *D = dpD_ + dpOff_[row];
#define AT(X, a, b) X[(size_t)(a) * cols + (b)]
		for (u32 j = 1; j < cols; j++) { AT(D, 0, j) = INV; AT(I, 0, j) = j * ce + co; }
		for (u32 a = 1; a < rows; a++) { AT(I, a, 0) = INV; AT(D, a, 0) = a * ce + co; }
		AT(M, 0, 0) = 0; AT(D, 0, 0) = INV; AT(I, 0, 0) = INV;
		for (u32 a = 1; a < rows; a++) AT(M, a, 0) = INV;
		for (u32 j = 1; j < cols; j++) AT(M, 0, j) = INV;
		auto mn = [](u32 x, u32 y) -> u32 { return x < y ? x : y; };
		for (u32 a = 1; a < rows; a++)
			for (u32 j = 1; j < cols; j++) {
				AT(I, a, j) = mn(AT(M, a - 1, j) + co + ce, AT(I, a - 1, j) + ce);      // u32 wrap-around is the reference's behaviour
				AT(D, a, j) = mn(AT(M, a, j - 1) + co + ce, AT(D, a, j - 1) + ce);
				bool m = eq(a - 1, j - 1);
				AT(M, a, j) = mn(AT(M, a - 1, j - 1) + (m ? ca : cx), mn(AT(I, a, j), AT(D, a, j)));
			}
		u32 a = Lr, j = Lh, n = 0;
		while (a > 0 && j > 0) {
			bool m = eq(a - 1, j - 1);
			u32 up = AT(M, a - 1, j), lf = AT(M, a, j - 1), dg = AT(M, a - 1, j - 1) + (m ? ca : cx), cur = AT(M, a, j);
			if (cur == mn(up, mn(lf, dg))) { e[n++] = m ? 'M' : 'X'; a--; j--; }
			else if (up == mn(up, mn(lf, cur))) { e[n++] = 'I'; a--; }
			else if (lf == mn(up, mn(lf, cur))) { e[n++] = 'D'; j--; }
			else { atomic_add(errs + 0, 1u); break; }   // the reference exit(1)s here
		}
#undef AT
		while (a > 0) { e[n++] = 'I'; a--; }
		while (j > 0) { e[n++] = 'D'; j--; }
		for (u32 x = 0; x < n / 2; x++) { u8 tt = e[x]; e[x] = e[n - 1 - x]; e[n - 1 - x] = tt; }
		etLen_[row] = n;
	});
	// ---- one unrolled matrix per (region, ref lap j) whose alignment column j holds an 'X' (untangle.cpp:178-209)
	Buf<u32> LrOf(c, (size_t)nT + 1), jPre(c, (size_t)nT + 1);
	u32 *LrOf_ = LrOf.p;
	parallel_for<T_rows>(c, (size_t)nT + 1, PV_LAMBDA(size_t t) { LrOf_[t] = t < nT ? lapCnt_[tRowPre_[t] + ref0Local] : 0u; });
	exclusive_scan(c, LrOf.p, jPre.p, (size_t)nT + 1, OpAddU32(), 0u);
	const u32 nJ = jPre.read(nT);
	if (!nJ) return;
	Buf<u32> mm(c, nJ), instIdx;
	u32 *mm_ = mm.p; const u32 *jPre_ = jPre.p;
	parallel_for<T_mismatch>(c, nJ, PV_LAMBDA(size_t x) {
		u32 t = ub_u32(jPre_, 0, nT + 1, (u32)x) - 1; u32 j = (u32)x - jPre_[t];
		u32 hit = 0;
		for (u32 row = tRowPre_[t]; row < tRowPre_[t + 1] && !hit; row++)
			if (etLen_[row] > j && et_[etOff_[row] + j] == 'X') hit = 1;
		mm_[x] = hit;
	});
	const u32 nI = compact_flags(c, mm.p, nJ, instIdx);
	if (!nI) return;
	Buf<u32> instT(c, nI), instJ(c, nI), instRegion(c, nI), instTangled(c, nI), hasData(c, nI);
	Buf<u64> iRows(c, (size_t)nI + 1), iCells(c, (size_t)nI + 1), iRowPre(c, (size_t)nI + 1), iCellPre(c, (size_t)nI + 1);
	instTangled.zero(); hasData.zero();
	u32 *instT_ = instT.p, *instJ_ = instJ.p, *instRegion_ = instRegion.p; u64 *iRows_ = iRows.p, *iCells_ = iCells.p; const u32 *instIdx_ = instIdx.p;
	parallel_for<T_inst>(c, (size_t)nI + 1, PV_LAMBDA(size_t n) {
		if (n >= nI) { iRows_[n] = 0; iCells_[n] = 0; return; }
		u32 x = instIdx_[n]; u32 t = ub_u32(jPre_, 0, nT + 1, x) - 1;
		instT_[n] = t; instJ_[n] = x - jPre_[t]; instRegion_[n] = rv.region(tList_[t]);
		u32 nh = tRowPre_[t + 1] - tRowPre_[t];
		iRows_[n] = nh; iCells_[n] = (u64)nh * rv.J(tList_[t]);
	});
	exclusive_scan(c, iRows.p, iRowPre.p, (size_t)nI + 1, OpAddU64(), (u64)0);
	exclusive_scan(c, iCells.p, iCellPre.p, (size_t)nI + 1, OpAddU64(), (u64)0);
	u64 totRows, totCells; c.d2h(&totRows, iRowPre.p + nI, 8); c.d2h(&totCells, iCellPre.p + nI, 8);
	if (totRows > 0x7FFFFFF0ull) throw std::runtime_error("call: too many unrolled matrix rows in one chunk");
	Buf<u16> imat(c, totCells ? totCells : 1);
	Buf<u32> ilapStart(c, totRows ? totRows : 1), ilapLen(c, totRows ? totRows : 1);
	c.memset0(imat.p, totCells * 2); ilapStart.fill_ff(); ilapLen.zero();
	const RovView irv{rv.rovOff, rv.rovVerts, rv.rovComp, iRowPre.p, iCellPre.p, rv.hkey, rv.hval, rv.hmask, nI, rv.uniformNh, 0, instJ.p, instRegion.p};
	const ChunkView icv{g, irv, imat.p, ilapStart.p, ilapLen.p, instTangled.p, hasData.p, errs};
	parallel_for<T_matrix>(c, (size_t)totRows, PV_LAMBDA(size_t row) {
		u32 n = irv.row2rov((u32)row); u32 i = (u32)row - irv.rowBase(n);
		u32 t = instT_[n], j = instJ_[n];
		u32 x = tRowPre_[t] + i;
		u32 lapIdx;
		if (i == ref0Local) lapIdx = j;
		else {
			if (etLen_[x] <= j) return;
			u8 ch = et_[etOff_[x] + j];
			if (ch == 'D' || ch == 'I') return;
			u32 ln = 0;
			for (u32 q = 0; q < j; q++) if (et_[etOff_[x] + q] != 'I') ln++;   // comp_loop_no (untangle.cpp:165-176)
			lapIdx = ln;
		}
		if (lapIdx >= lapCnt_[x]) { atomic_add(errs + 0, 1u); return; }   // out-of-range race index in the reference
		u32 s0 = lapS_[sOff_[x] + lapIdx], len = lapL_[sOff_[x] + lapIdx];
		icv.lapStart[row] = s0; icv.lapLen[row] = len;
		u32 J = irv.J(n); u16 *cells = icv.mat + irv.matBase(n) + (u64)i * J;
		for (u32 q = 0; q < len; q++) {   // fill_row (untangle.cpp:126-143): later steps overwrite earlier ones
			u32 st = g.steps[s0 + q]; u32 col = irv.find(n, st >> 1);
			if (col != NONE) cells[col] = (st & 1u) ? 2 : 1;
		}
	});
	tr.mark("T untangle");
	overlay_pass(c, G, g, in, irv, icv, nI, (u32)totRows, out, tr, pins);
}


inline void call_device(Ctx &c, const DevGraph &G, const CallIn &in, const CallOpts &opt, Records &out) {
	out.reset();
	out.rec_grpOff.push_back(0); out.rec_gtOff.push_back(0);
	out.cpOff = G.cpOff; out.cpList = G.cpList;
	const u32 nFl = in.nFl, nRefs = in.nRefs, H = G.H; const u32 flOff = in.fl_offset;
	if (nFl == 0 || nRefs == 0 || H == 0) return;
	const GraphView g = view(G);
	Trace tr(c, "call");

	// ------------------------------------------------------------------ C1: region selection
	Buf<u32> fa(c, nFl), fz(c, nFl), fpar(c, nFl), fbase(c, nFl), refs(c, nRefs), colored(c, nFl), kid(c, nFl), called(c, nFl);
	{
		Buf<u8> af(c, nFl), zf(c, nFl), ex(c, nFl);
		if (in.on_device) {
			c.d2d(fa.p, in.a_v, (size_t)nFl * 4); c.d2d(fz.p, in.z_v, (size_t)nFl * 4); c.d2d(af.p, in.a_fwd, nFl); c.d2d(zf.p, in.z_fwd, nFl);
			c.d2d(fpar.p, in.parent, (size_t)nFl * 4); c.d2d(fbase.p, in.flBase, (size_t)nFl * 4);
		} else {
			c.h2d(fa.p, in.a_v, (size_t)nFl * 4); c.h2d(fz.p, in.z_v, (size_t)nFl * 4); c.h2d(af.p, in.a_fwd, nFl); c.h2d(zf.p, in.z_fwd, nFl);
			c.h2d(fpar.p, in.parent, (size_t)nFl * 4); c.h2d(fbase.p, in.flBase, (size_t)nFl * 4);
		}
		c.h2d(refs.p, in.refs, (size_t)nRefs * 4);
		if (in.excluded) c.h2d(ex.p, in.excluded, nFl);
		u32 *fa_w = fa.p, *fz_w = fz.p; const u8 *af_ = af.p, *zf_ = zf.p, *ex_ = in.excluded ? ex.p : nullptr;
		u32 *fbase_w = fbase.p; const u32 flOff0 = in.fl_offset;
		parallel_for<C_pack>(c, nFl, PV_LAMBDA(size_t f) {   // (vertex << 1) | reverse; NONE hides an excluded flubble
			fa_w[f] = (ex_ && ex_[f]) ? NONE : ((fa_w[f] << 1) | (af_[f] ? 0u : 1u));
			fz_w[f] = (fz_w[f] << 1) | (zf_[f] ? 0u : 1u);
			fbase_w[f] -= flOff0;   // slice-relative from here on
		});
	}
	const u32 *fa_ = fa.p, *fz_ = fz.p, *fpar_ = fpar.p, *fbase_ = fbase.p, *refs_ = refs.p;
	u32 *colored_ = colored.p, *kid_ = kid.p, *called_ = called.p;
	kid.zero();
	parallel_for<C_color>(c, nFl, PV_LAMBDA(size_t f) {   // color.cpp:15-30
		if (fa_[f] == NONE) { colored_[f] = 0; return; }
		u32 a = fa_[f] >> 1, z = fz_[f] >> 1; u32 ok = 1;
		for (u32 k = 0; k < nRefs && ok; k++) { u32 h = refs_[k]; if (!g.has_hap(a, h) || !g.has_hap(z, h)) ok = 0; }
		colored_[f] = ok;
	});
	parallel_for<C_child>(c, nFl, PV_LAMBDA(size_t f) { u32 p = fpar_[f]; if (colored_[f] && p != 0) kid_[fbase_[f] + p - 1] = 1; });
	parallel_for<C_child>(c, nFl, PV_LAMBDA(size_t f) { called_[f] = (colored_[f] && !kid_[f]) ? 1u : 0u; });   // color.cpp:141-148
	if (in.restrict_on) {
		const u32 rh = in.restrict_ref; const u64 lo = in.restrict_start, hi = in.restrict_end;
		parallel_for<C_restrict>(c, nFl, PV_LAMBDA(size_t f) {
			if (!called_[f]) return;
			bool hit = false;
			if (rh != NONE) {
				u32 s0 = g.stepOff[rh], s1 = g.stepOff[rh + 1];
				for (int side = 0; side < 2 && !hit; side++) {
					u32 v = (side ? fz_[f] : fa_[f]) >> 1;
					u32 a = g.vsOff[v], b = g.vsOff[v + 1];
					for (u32 k = lb_u32(g.vsStep, a, b, s0); k < b && g.vsStep[k] < s1; k++) { u64 pos = g.loci[g.vsStep[k]]; if (pos >= lo && pos < hi) { hit = true; break; } }
				}
			}
			if (!hit) called_[f] = 0;
		});
	}
	// (the number of called regions R0 is not read back here: tier 0 below runs over the upper bound nFl and looks the count up
	// on the device, so that R0 and the tier-0 counters share one host wait)
	Buf<u32> rov0, calledPos;
	u32 R0 = 0;
	MultiRead mrSel(c);
	compact_flags(c, called.p, nFl, rov0, &calledPos, &mrSel, &R0);
	tr.mark("C1 select");

	// ------------------------------------------------------------------ C2: vertex order per region
	// tier 0: thread-private scratch; overflowing regions retry with 16x larger slices of a global arena
	Buf<u32> J0(c, nFl), vtmp(c, (size_t)nFl * 12), cnt(c, 8);
	cnt.zero();
	u32 *J0_ = J0.p, *vtmp_ = vtmp.p, *cnt_ = cnt.p; const u32 *rov0_ = rov0.p, *R0dev = calledPos.p + nFl;
	parallel_for<C_bfs>(c, nFl, PV_LAMBDA(size_t r) {
		if (r >= *R0dev) return;
		u32 tv[12], es[13], et[32], dg[12], od[12], ov[12];
		BfsScratch s{12, 32, tv, es, et, dg, od, nullptr, nullptr, 0};
		u32 f = rov0_[r];
		u32 n = bfs_kahn_sort(g, fa_[f] >> 1, !(fa_[f] & 1u), fz_[f] >> 1, !(fz_[f] & 1u), s, ov);
		J0_[r] = n;
		if (n == NONE) { atomic_add(cnt_ + 0, 1u); return; }
		if (n == 0) atomic_add(cnt_ + 3, 1u);
		for (u32 k = 0; k < n; k++) vtmp_[r * 12 + k] = ov[k];
	});
	u32 hc[8]; mrSel.add(hc, cnt.p, 32); mrSel.wait();
	out.n_called = R0;
	if (R0 == 0) return;
	// regions that did not fit tier 0 retry in 16x larger slices of a global arena.  The orders stay where the
	// kernel wrote them (the arenas live until the gather below); only the per-slice result counts go to the host.
	std::vector<Buf<u32>> arenas;
	std::vector<u64> sliceTab;   // per slice: arena base pointer, u32 words per region, offset of the order inside a region's words
	Buf<u32> bigSlice(c, R0), bigSlot(c, R0);
	const bool haveBig = hc[0] != 0;
	if (haveBig) {
		bigSlice.fill_ff();
		u32 *bigSlice_ = bigSlice.p, *bigSlot_ = bigSlot.p;
		Buf<u32> ovf(c, R0), todoBuf;
		u32 *ovf_ = ovf.p;
		parallel_for<C_keep>(c, R0, PV_LAMBDA(size_t r) { ovf_[r] = J0_[r] == NONE ? 1u : 0u; });
		u32 nTodo = compact_flags(c, ovf.p, R0, todoBuf);
		std::vector<u32> todo = todoBuf.to_host(nTodo);
		u64 capV64 = 256;
		while (!todo.empty()) {
			const bool last = capV64 >= (u64)G.V;
			const u32 capV = (u32)std::min<u64>(capV64, (u64)G.V + 1);
			const u32 capE = (u32)std::min<u64>(capV64 * 4, 2 * G.E + 1);
			u32 hcap = 1; while (hcap < 2 * capV) hcap <<= 1;
			const u64 per = (u64)capV * 5 + 1 + capE + 2ull * hcap;
			const u64 ovOff = 4 * (u64)capV + 1 + capE;
			const u32 slice = (u32)std::max<u64>(1, std::min<u64>(todo.size(), (1ull << 28) / per));
			std::vector<u32> next;
			for (size_t b0 = 0; b0 < todo.size(); b0 += slice) {
				const u32 nb = (u32)std::min<size_t>(slice, todo.size() - b0);
				Buf<u32> list(c, nb), rs(c, nb);
				arenas.emplace_back(c, (size_t)nb * per);
				const u32 sliceIdx = (u32)(sliceTab.size() / 3);
				sliceTab.push_back((u64)(uintptr_t)arenas.back().p); sliceTab.push_back(per); sliceTab.push_back(ovOff);
				c.h2d(list.p, todo.data() + b0, (size_t)nb * 4);
				u32 *arena_ = arenas.back().p, *res_ = rs.p; const u32 *list_ = list.p;
				parallel_for<C_bfs>(c, nb, PV_LAMBDA(size_t i) {
					u32 *base = arena_ + i * per;
					const size_t cv = capV, ce = capE;
					BfsScratch s{capV, capE, base, base + cv, base + 2 * cv + 1, base + 2 * cv + 1 + ce, base + 3 * cv + 1 + ce,
						     base + 5 * cv + 1 + ce, base + 5 * cv + 1 + ce + hcap, hcap};
					u32 r = list_[i]; u32 f = rov0_[r];
					u32 n = bfs_kahn_sort(g, fa_[f] >> 1, !(fa_[f] & 1u), fz_[f] >> 1, !(fz_[f] & 1u), s, base + ovOff);
					res_[i] = n;
					if (n == 0) atomic_add(cnt_ + 3, 1u);
					if (n != NONE) { J0_[r] = n; bigSlice_[r] = sliceIdx; bigSlot_[r] = (u32)i; }
				});
				std::vector<u32> hres = rs.to_host();
				for (u32 i = 0; i < nb; i++) { if (hres[i] == NONE) next.push_back(todo[b0 + i]); else out.n_big_bfs++; }
			}
			if (!next.empty() && last) throw std::runtime_error("call: region BFS overflow at full-graph capacity");
			todo.swap(next);
			capV64 *= 16;
		}
	}
	Buf<u64> sliceTabD(c, sliceTab.size() ? sliceTab.size() : 1);
	c.h2d(sliceTabD.p, sliceTab.data(), sliceTab.size() * 8);
	if (haveBig) c.d2h(hc, cnt.p, 32);   // the arena tiers may have counted more failed sorts; without them the first read still holds
	tr.mark("C2 bfs");
	// regions whose BFS/Kahn sort failed (a cycle inside the region): the reference's haplotype-lap order
	// (graph.cpp:194-395), one thread per region over an arena sized by the region's total lap length
	Buf<u32> gsArena, gsSrc(c, hc[3] ? R0 : 1);
	Buf<u64> gsMeta;
	const bool haveGs = hc[3] != 0;
	if (haveGs) {
		gsSrc.fill_ff();
		Buf<u32> failF(c, R0), failList;
		u32 *failF_ = failF.p;
		parallel_for<C_keep>(c, R0, PV_LAMBDA(size_t r) { failF_[r] = J0_[r] == 0 ? 1u : 0u; });
		const u32 nGs = compact_flags(c, failF.p, R0, failList);
		Buf<u64> capD(c, nGs);
		u64 *capD_ = capD.p; const u32 *failList_ = failList.p;
		parallel_for<G_count>(c, nGs, PV_LAMBDA(size_t i) {
			u32 f = rov0_[failList_[i]]; u64 tot = 0;
			for (u32 h = 0; h < g.H; h++) { GsLaps L(g, fa_[f] >> 1, fz_[f] >> 1, h); for (u32 li = 0; li < L.n; li++) { u32 st, ln; L.lap(g, li, st, ln); tot += ln; } }
			capD_[i] = tot;
		});
		std::vector<u64> caps = capD.to_host();
		std::vector<u64> offs(nGs + 1, 0), meta((size_t)nGs * 3);
		for (u32 i = 0; i < nGs; i++) {
			u64 cap = std::min<u64>(caps[i], (u64)G.V) + 1;
			u64 hcap = 2; while (hcap < 2 * cap) hcap <<= 1;
			meta[3 * (size_t)i] = offs[i]; meta[3 * (size_t)i + 1] = cap; meta[3 * (size_t)i + 2] = hcap;
			offs[i + 1] = offs[i] + 4 * cap + 6 * hcap;   // out + val + nxt + prv, then the six table columns
		}
		if (offs[nGs] > (1ull << 31)) throw std::runtime_error("call: the haplotype-lap vertex order of the cyclic regions needs more than 8 GB of scratch");
		gsArena = Buf<u32>(c, offs[nGs] ? offs[nGs] : 1);
		Buf<u64> metaD(c, meta.size() ? meta.size() : 1);
		c.h2d(metaD.p, meta.data(), meta.size() * 8);
		u32 *arena_ = gsArena.p, *gsSrc_ = gsSrc.p; const u64 *meta_ = metaD.p;
		parallel_for<G_sort>(c, nGs, PV_LAMBDA(size_t i) {
			u32 r = failList_[i]; u32 f = rov0_[r];
			u32 *b = arena_ + meta_[3 * i]; const u32 cap = (u32)meta_[3 * i + 1], hcap = (u32)meta_[3 * i + 2];
			u32 *t = b + 4 * (size_t)cap;
			GsScratch s{cap, hcap, b + cap, b + 2 * (size_t)cap, b + 3 * (size_t)cap, t, t + hcap, t + 2 * (size_t)hcap, t + 3 * (size_t)hcap, t + 4 * (size_t)hcap, t + 5 * (size_t)hcap};
			u32 n = gen_sort_laps(g, fa_[f] >> 1, fz_[f] >> 1, s, b);
			if (n == GS_CRASH) { atomic_add(cnt_ + 4, 1u); n = 0; }
			J0_[r] = n; gsSrc_[r] = (u32)i;
		});
		c.sync();   // meta / caps live on the host stack of this block
		gsMeta = std::move(metaD);
		tr.mark("C2 lap order");
	}
	// keep regions with >= 3 sorted vertices (rov.cpp:208-212); an empty lap order too is skipped there
	Buf<u32> keepF(c, R0), rovSrc;
	u32 *keepF_ = keepF.p;
	parallel_for<C_keep>(c, R0, PV_LAMBDA(size_t r) {
		u32 n = J0_[r];
		keepF_[r] = n >= 3 ? 1u : 0u;
		if (n == 0) atomic_add(cnt_ + 1, 1u); else if (n < 3) atomic_add(cnt_ + 2, 1u);
	});
	const u32 R = compact_flags(c, keepF.p, R0, rovSrc);
	if (R == 0) { c.d2h(hc, cnt.p, 32); out.n_sort_failed = hc[1]; out.n_too_small = hc[2]; out.n_hard_errors = hc[4]; return; }
	Buf<u32> rovJ(c, (size_t)R + 1), rovOff(c, (size_t)R + 1), rovFl(c, R), rovComp(c, R), rovTangled(c, R);
	Buf<u64> rowsOf(c, (size_t)R + 1), cellsOf(c, (size_t)R + 1), rowPre(c, (size_t)R + 1), cellPre(c, (size_t)R + 1), tot(c, 4);
	tot.zero(); rovTangled.zero();
	u32 *rovJ_ = rovJ.p, *rovFl_ = rovFl.p, *rovComp_ = rovComp.p, *rovTangled_ = rovTangled.p; const u32 *rovSrc_ = rovSrc.p;
	u64 *rowsOf_ = rowsOf.p, *cellsOf_ = cellsOf.p, *tot_ = tot.p;
	parallel_for<C_sizes>(c, (size_t)R + 1, PV_LAMBDA(size_t r) {
		if (r >= R) { rovJ_[r] = 0; rowsOf_[r] = 0; cellsOf_[r] = 0; return; }
		u32 s = rovSrc_[r]; u32 n = J0_[s]; u32 f = rov0_[s];
		rovJ_[r] = n; rovFl_[r] = f + flOff;   // global flubble index
		// component of the region = component of any path through its start boundary (a selected region is
		// touched by every reference path)
		u32 a = fa_[f] >> 1;
		u32 comp = g.pathComp[g.vsHap[g.vsOff[a]] >> 1];
		rovComp_[r] = comp;
		u32 nh = g.cpOff[comp + 1] - g.cpOff[comp];
		rowsOf_[r] = nh; cellsOf_[r] = (u64)nh * n;
		if ((u64)n > *(volatile u64 *)(tot_ + 0)) atomic_max(tot_ + 0, (u64)n);   // test first: one contended address otherwise serialises every thread
		if (n > 16) atomic_add(tot_ + 1, (u64)n);
	});
	exclusive_scan(c, rovJ.p, rovOff.p, (size_t)R + 1, OpAddU32(), 0u);
	exclusive_scan(c, rowsOf.p, rowPre.p, (size_t)R + 1, OpAddU64(), (u64)0);
	exclusive_scan(c, cellsOf.p, cellPre.p, (size_t)R + 1, OpAddU64(), (u64)0);
	u64 ht[4]; u64 totRows, totCells; u32 sumJ;
	{ MultiRead mr(c); mr.add(ht, tot.p, 32); mr.add(&totRows, rowPre.p + R, 8); mr.add(&totCells, cellPre.p + R, 8); mr.add(&sumJ, rovOff.p + R, 4); mr.wait(); }
	out.max_J = ht[0];
	Buf<u32> rovVerts(c, sumJ);
	u32 *rovVerts_ = rovVerts.p; const u32 *rovOff_ = rovOff.p;
	{
		const u32 *bigSlice_ = bigSlice.p, *bigSlot_ = bigSlot.p; const u64 *tab = sliceTabD.p;
		const u32 *gsSrc_ = gsSrc.p, *gsArena_ = gsArena.p; const u64 *gsMeta_ = gsMeta.p;
		parallel_for<C_gather>(c, R, PV_LAMBDA(size_t r) {
			u32 s = rovSrc_[r]; u32 n = rovJ_[r]; u32 o = rovOff_[r];
			const u32 *src = vtmp_ + (size_t)s * 12;
			if (haveBig && bigSlice_[s] != NONE) { const u64 *t = tab + 3 * (size_t)bigSlice_[s]; src = (const u32 *)(uintptr_t)t[0] + (size_t)bigSlot_[s] * t[1] + t[2]; }
			if (haveGs && gsSrc_[s] != NONE) src = gsArena_ + gsMeta_[3 * (size_t)gsSrc_[s]];
			for (u32 k = 0; k < n; k++) rovVerts_[o + k] = src[k];
		});
	}
	rovFl.to_hvec(out.rov_fl, R); rovComp.to_hvec(out.rov_comp, R);
	tr.mark("C2 gather");

	// hash (region, vertex) -> column for regions with more than 16 vertices
	const u64 bigJ = ht[1];
	u64 hsize = 2; while (hsize < 2 * bigJ + 2) hsize <<= 1;
	Buf<u64> hkey(c, hsize); Buf<u32> hval(c, hsize);
	hkey.fill_ff();
	if (bigJ) {
		u64 *hkey_ = hkey.p; u32 *hval_ = hval.p; const u64 hmask = hsize - 1;
		parallel_for<C_hash>(c, sumJ, PV_LAMBDA(size_t x) {
			u32 r = ub_u32(rovOff_, 0, R + 1, (u32)x) - 1;
			if (rovOff_[r + 1] - rovOff_[r] <= 16) return;
			u64 key = ((u64)r << 32) | rovVerts_[x]; u64 i = splitmix64(key) & hmask;
			while (atomic_cas64(hkey_ + i, ~0ull, key) != ~0ull) i = (i + 1) & hmask;
			hval_[i] = (u32)x - rovOff_[r];
		});
	}

	// ------------------------------------------------------------------ chunks of regions bounded by matrix cells
	u32 uniformNh = 0;
	{
		bool uni = G.cpOff.size() > 1;
		for (size_t k = 1; k + 1 < G.cpOff.size(); k++) if (G.cpOff[k + 1] - G.cpOff[k] != G.cpOff[1] - G.cpOff[0]) uni = false;
		if (uni) uniformNh = G.cpOff[1] - G.cpOff[0];
	}
	std::vector<u32> chunkEnd;   // exclusive region index of every chunk
	std::vector<u64> hRowPre, hCellPre;
	if (totCells <= opt.cell_budget && totRows <= 0x7FFFFFF0ull) chunkEnd.push_back(R);
	else {
		hRowPre.resize((size_t)R + 1); hCellPre.resize((size_t)R + 1);
		c.d2h(hRowPre.data(), rowPre.p, ((size_t)R + 1) * 8); c.d2h(hCellPre.data(), cellPre.p, ((size_t)R + 1) * 8);
		for (u32 r0 = 0; r0 < R;) {
			u32 r1 = r0 + 1;
			while (r1 < R && hCellPre[r1 + 1] - hCellPre[r0] <= opt.cell_budget && hRowPre[r1 + 1] - hRowPre[r0] <= 0x7FFFFFF0ull) r1++;
			chunkEnd.push_back(r1); r0 = r1;
		}
	}
	PinStore pins;
	Buf<u32> errs(c, 8);   // [0] hard errors, [2] depth overflow, [3] tangled regions, [4] SUBR slices clamped to the walk
	errs.zero();
	u32 seenTangled = 0;

	u32 r0 = 0;
	for (u32 r1 : chunkEnd) {
		const u32 Rc = r1 - r0;
		u64 rows, cells;
		if (chunkEnd.size() == 1) { rows = totRows; cells = totCells; } else { rows = hRowPre[r1] - hRowPre[r0]; cells = hCellPre[r1] - hCellPre[r0]; }
		const u32 nRows = (u32)rows;
		out.n_chunks++;
		Buf<u16> mat(c, cells ? cells : 1);
		Buf<u32> lapStart(c, nRows ? nRows : 1), lapLen(c, nRows ? nRows : 1), hasData(c, Rc);
		c.memset0(mat.p, cells * 2); lapStart.fill_ff(); hasData.zero();
		const RovView rv{rovOff.p, rovVerts.p, rovComp.p, rowPre.p, cellPre.p, hkey.p, hval.p, hsize - 1, Rc, uniformNh, r0, nullptr, nullptr};
		const ChunkView cv{g, rv, mat.p, lapStart.p, lapLen.p, rovTangled_ + r0, hasData.p, errs.p};
		tr.mark("chunk setup");
		// ---- C3 + C4 + the first pass of C5 as ONE kernel, one warp per region (GPU build): the lanes stride over the step lists of
		// the region's vertices (depth cells, first steps), then over its rows (lap 0), then over its rows again (contexts of every
		// path against the first reference path) -- the three kernels below run back to back on data the same SM just wrote,
		// and the per-row index arithmetic (region of a row, matrix base, row base) is done once per region.
		Buf<u32> pairCnt0;
		bool fused = false;
#ifndef PV_EMU
		// Measured: with 16 paths a warp per region idles half its lanes and loses (703 vs 405 us for the three kernels at cfg2);
		// with 464 paths it wins (4.7 vs 7.7 ms at 200 k bubbles).  Regions with >= 48 rows on average take the fused kernel.
		fused = (getenv("POVU_B200_FUSED_REGION") != nullptr || (u64)nRows >= 48ull * Rc) && getenv("POVU_B200_NO_FUSED_REGION") == nullptr && nRows > 0;
		if (fused) {
			pairCnt0 = Buf<u32>(c, (size_t)nRows + 1);
			u32 *pc0 = pairCnt0.p;
			const u32 ref0Local = G.hapLocal[in.refs[0]];
			parallel_for<C_region>(c, (size_t)Rc * 32, PV_LAMBDA(size_t t) {
				const u32 r = (u32)(t >> 5), lane = (u32)(t & 31);
				const u32 J = rv.J(r), comp = rv.comp(r), rb = rv.rowBase(r), nh = rv.rowBase(r + 1) - rb;
				const u64 mb = rv.matBase(r);
				for (u32 j = 0; j < J; j++) {
					const u32 v = rv.vert(r, j);
					const u32 a = g.vsOff[v], ee = g.vsOff[v + 1];
					for (u32 e = a + lane; e < ee; e += 32) {   // (the neighbour tests hit the L1 lines the adjacent lanes just loaded: exchanging them by shuffle instead was measured 7 % slower)
						const u32 hs = g.vsHap[e]; u32 h = hs >> 1;
						if (e > a && (g.vsHap[e - 1] >> 1) == h) continue;   // not the head of its run
						u32 first = g.vsStep[e]; u32 n = 1;
						while (e + n < ee && (g.vsHap[e + n] >> 1) == h) n++;
						if (g.pathComp[h] != comp) continue;   // cannot happen for a connected walk; keeps indexing safe
						u32 i = g.hapLocal[h];
						if (n > 0xFFFEu) { atomic_add(cv.errs + 2, 1u); n = 0xFFFEu; }
						u32 code = n == 1 ? ((hs & 1u) ? 2u : 1u) : n;
						cv.mat[mb + (u64)i * J + j] = (u16)code;
						atomic_min(cv.lapStart + rb + i, first);
						if (n > 1 && (j == 0 || j == J - 1) && atomic_exch(cv.tangled + r, 1u) == 0) atomic_add(cv.errs + 3, 1u);
					}
				}
				__threadfence_block(); __syncwarp();
				for (u32 i = lane; i < nh; i += 32) {
					const u32 row = rb + i;
					u32 s0 = ld_cg(cv.lapStart + row);
					if (s0 == NONE) { cv.lapLen[row] = 0; continue; }
					u32 hh = ub_u32(g.stepOff, 0, g.H + 1, s0) - 1;
					u32 end = g.stepOff[hh + 1];
					u32 len = 1;
					while (s0 + len < end && rv.find(r, g.steps[s0 + len] >> 1) != NONE) len++;
					cv.lapLen[row] = len;
				}
				__threadfence_block(); __syncwarp();
				for (u32 i = lane; i < nh; i += 32) pc0[rb + i] = scan_pair(cv, ref0Local, 0, rb + i, false, CandView{}, 0);
				if (r == Rc - 1 && lane == 0) pc0[nRows] = 0;
			});
		}
#endif
		// ---- C3: depth matrix (allele.hpp:804-862), one thread per (region, column)
		u32 colBase, colEnd;
		if (chunkEnd.size() == 1) { colBase = 0; colEnd = sumJ; } else { colBase = rovOff.read(r0); colEnd = rovOff.read(r1); }
		const u32 nCols = colEnd - colBase;
		// LW logical lanes share one column and stride over the vertex's step list (coalesced when LW = 32); a lane owns the
		// runs (consecutive steps of one path) whose first entry it meets.  LW follows the mean list length: with 16 paths
		// a list has ~12 entries (one lane: 189 us at cfg2, two lanes would take 384 us), with 464 paths ~350 (32 lanes).
		u32 LW = 1;
		{ const u64 meanList = G.V ? (u64)G.S / G.V : 0; LW = meanList >= 64 ? 32u : (meanList >= 32 ? 8u : 1u); }   // (every lane repeats the column setup: short lists keep one lane)
		if (const char *e = getenv("POVU_B200_MATFILL_LANES")) { u32 v = (u32)strtoul(e, nullptr, 10); if (v >= 1 && v <= 32) LW = v; }   // tests: force the multi-lane path
		if (fused) { }
		else if (LW == 1) parallel_for<C_matfill>(c, nCols, PV_LAMBDA(size_t x) {   // short lists: one thread walks the list run by run
			u32 r = ub_u32(rovOff_ + r0, 0, Rc + 1, colBase + (u32)x) - 1;
			u32 J = rv.J(r); u32 j = colBase + (u32)x - rovOff_[r0 + r];
			u32 v = rovVerts_[colBase + x];
			u32 comp = rv.comp(r);
			u32 e = g.vsOff[v], ee = g.vsOff[v + 1];
			while (e < ee) {
				const u32 hs = g.vsHap[e]; u32 h = hs >> 1; u32 first = g.vsStep[e]; u32 n = 1;
				while (e + n < ee && (g.vsHap[e + n] >> 1) == h) n++;
				e += n;
				if (g.pathComp[h] != comp) continue;   // cannot happen for a connected walk; keeps indexing safe
				u32 i = g.hapLocal[h];
				if (n > 0xFFFEu) { atomic_add(cv.errs + 2, 1u); n = 0xFFFEu; }
				u32 code = n == 1 ? ((hs & 1u) ? 2u : 1u) : n;
				cv.mat[rv.matBase(r) + (u64)i * J + j] = (u16)code;
				atomic_min(cv.lapStart + rv.rowBase(r) + i, first);
				if (n > 1 && (j == 0 || j == J - 1) && atomic_exch(cv.tangled + r, 1u) == 0) atomic_add(cv.errs + 3, 1u);
			}
		});
		else parallel_for<C_matfill>(c, (size_t)nCols * LW, PV_LAMBDA(size_t t) {
			const u32 x = (u32)(t / LW), lane = (u32)(t % LW);
			u32 r = ub_u32(rovOff_ + r0, 0, Rc + 1, colBase + x) - 1;
			u32 J = rv.J(r); u32 j = colBase + x - rovOff_[r0 + r];
			u32 v = rovVerts_[colBase + x];
			u32 comp = rv.comp(r);
			const u32 a = g.vsOff[v], ee = g.vsOff[v + 1];
			for (u32 e = a + lane; e < ee; e += LW) {
				const u32 hs = g.vsHap[e]; u32 h = hs >> 1;
				if (e > a && (g.vsHap[e - 1] >> 1) == h) continue;   // not the head of its run
				u32 first = g.vsStep[e]; u32 n = 1;
				while (e + n < ee && (g.vsHap[e + n] >> 1) == h) n++;
				if (g.pathComp[h] != comp) continue;   // cannot happen for a connected walk; keeps indexing safe
				u32 i = g.hapLocal[h];
				if (n > 0xFFFEu) { atomic_add(cv.errs + 2, 1u); n = 0xFFFEu; }
				u32 code = n == 1 ? ((hs & 1u) ? 2u : 1u) : n;
				cv.mat[rv.matBase(r) + (u64)i * J + j] = (u16)code;
				atomic_min(cv.lapStart + rv.rowBase(r) + i, first);
				if (n > 1 && (j == 0 || j == J - 1) && atomic_exch(cv.tangled + r, 1u) == 0) atomic_add(cv.errs + 3, 1u);
			}
		});
		// ---- C4: lap 0 of every (region, path): the first run of consecutive steps on region vertices (untangle.cpp:29-103)
		if (!fused) parallel_for<C_lap>(c, nRows, PV_LAMBDA(size_t row) {
			u32 s = cv.lapStart[row];
			if (s == NONE) { cv.lapLen[row] = 0; return; }
			u32 r = rv.row2rov((u32)row);
			u32 hh = ub_u32(g.stepOff, 0, g.H + 1, s) - 1;
			u32 end = g.stepOff[hh + 1];
			u32 len = 1;
			while (s + len < end && rv.find(r, g.steps[s + len] >> 1) != NONE) len++;
			cv.lapLen[row] = len;
		});
		tr.mark("C3 C4 matrix+laps");

		u32 he3 = seenTangled;
		overlay_pass(c, G, g, in, rv, cv, Rc, nRows, out, tr, pins, &he3, fused ? &pairCnt0 : nullptr);      // (the tangled flags were set by C_matfill above)
		if (he3 != seenTangled) { seenTangled = he3; tangled_pass(c, G, g, in, rv, cv, Rc, rovTangled_ + r0, errs.p, out, tr, pins); }
		r0 = r1;
	}
	rovTangled.to_hvec(out.rov_tangled, R);      // (asynchronous: complete after the wait below)
	out.n_regions = R;
	out.n_pin_pairs = pins.total;
	if (pins.total) sne_pass(c, G, g, in, pins, R, errs.p, out, tr);
	u32 he[8];
	{ MultiRead mr(c); mr.add(hc, cnt.p, 32); mr.add(he, errs.p, 32); mr.wait(); }      // one wait: counters + every copy queued above
	out.rov_tangled.resize(out.rov_fl.size(), 0);
	out.n_sort_failed = hc[1]; out.n_too_small = hc[2];
	out.n_hard_errors = he[0] + hc[4]; out.n_subr_clamped = he[4];
	if (he[2]) throw std::runtime_error("call: a path visits one vertex more than 65534 times inside a region (u16 depth cell overflow)");
	for (u32 t : out.rov_tangled) if (t) out.n_tangled++;
}

} // namespace pv
