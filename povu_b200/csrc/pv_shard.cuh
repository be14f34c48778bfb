// povu_b200 -- multi-GPU partitioning on the device: from the WHOLE graph resident in HBM to this rank's share
// of the weakly connected components, without the graph ever visiting the host.
//   component labels (union-find)  ->  per-component weights (links + path steps + segments; shared-memory histograms)
//   ->  [host: greedy LPT over the handful of component weights, the only host step]  ->  keep flags, scans, compaction of
//   segments / sequences / links / walks with remapped vertex indices.
// Vertex order, link order, path ids and the order of steps inside a path are preserved, so every result of a shard is what the
// whole-graph run produces for those components (decompose.cpp:102,116-153 treats components independently as well).
#pragma once
#include "pv_call.cuh"

namespace pv {

PV_KERNEL(P_root, 8);
PV_KERNEL(P_weight, 4);
PV_KERNEL(P_keep, 8);
PV_KERNEL(P_vertex, 24);
PV_KERNEL(P_label, 8);
PV_KERNEL(P_link, 24);
PV_KERNEL(P_step, 12);
PV_KERNEL(P_check, 12);

struct ShardIn {      // the whole graph in HBM
	u32 V = 0, H = 0; u64 E = 0, S = 0;
	const u32 *ep1 = nullptr, *ep2 = nullptr, *steps = nullptr, *segLen = nullptr, *ids = nullptr;
	const u64 *lblOff = nullptr; const u8 *lbl = nullptr;      // optional (sessions opened with labels)
	const u32 *h_stepOff = nullptr;                            // host, H + 1
	// walks still on the host (page-locked): with `steps` == nullptr only the walks of the paths this rank ends up owning
	// are uploaded -- a path lies inside one component, and the walks are the bulk of a pangenome (S >> E)
	const u32 *h_steps = nullptr;
};
struct ShardOut {     // this rank's share, in HBM (+ the small host tables)
	u32 V = 0; u64 E = 0, S = 0;
	Buf<u32> ep1, ep2, steps, segLen, ids; Buf<u64> lblOff; Buf<u8> lbl;
	std::vector<u32> h_stepOff, h_ids, h_segLen;
	std::vector<u32> comp_global;      // local component k (ascending lowest vertex) -> global component id (1-based)
	std::vector<u32> owner;            // per global component (0-based): owning rank
	u32 n_comp_total = 0;
	std::vector<u64> weight;           // per global component
};

// weights[k] += 1 for every item whose component is k.  One thread adds up a run of 256 consecutive items: consecutive items
// mostly share a component (a walk stays inside one, links follow the chain), so a thread issues one atomic per run of equal
// components instead of one per item -- the few hot counters (chromosomes) are not hammered.
struct CompOfPort { const u32 *port, *lab, *compIdx; PV_DEV u32 operator()(size_t i) const { return compIdx[lab[port[i] >> 1]]; } };
struct CompOfVertex { const u32 *lab, *compIdx; PV_DEV u32 operator()(size_t v) const { return compIdx[lab[v]]; } };
template <class CompOf> inline void shard_weights(Ctx &c, size_t n, CompOf comp_of, u64 *weights) {
	if (n == 0) return;
	parallel_for<P_weight>(c, (n + 255) / 256, PV_LAMBDA(size_t chunk) {
		size_t lo = chunk * 256, hi = lo + 256 < n ? lo + 256 : n;
		u32 cur = comp_of(lo); u64 cnt = 1;
		for (size_t i = lo + 1; i < hi; i++) { u32 k = comp_of(i); if (k == cur) cnt++; else { atomic_add(weights + cur, cnt); cur = k; cnt = 1; } }
		atomic_add(weights + cur, cnt);
	});
}

inline void shard_device(Ctx &c, const ShardIn &in, u32 rank, u32 world, ShardOut &out) {
	if (world == 0 || rank >= world) throw std::runtime_error("shard: bad rank / world size");
	const u32 V = in.V, H = in.H; const u64 E = in.E, S = in.S;
	if (V == 0) throw std::runtime_error("shard: empty graph");
	Trace tr(c, "shard");
	// ---- components: label = lowest vertex index; dense index = rank of the label among the roots (discovery order)
	Buf<u32> lab(c, V), isRoot(c, V), rootList, compIdx;
	component_labels(c, V, E, in.ep1, in.ep2, lab.p);
	u32 *lab_ = lab.p, *isRoot_ = isRoot.p;
	parallel_for<P_root>(c, V, PV_LAMBDA(size_t v) { isRoot_[v] = lab_[v] == (u32)v ? 1u : 0u; });
	const u32 nC = compact_flags(c, isRoot.p, V, rootList, &compIdx);      // compIdx[v] (exclusive scan of the root flags) at a root = its dense index
	const u32 *compIdx_ = compIdx.p;
	{	// every link must lie inside one component (the labels were computed from these very links: a cheap self-check)
		Buf<u32> bad(c, 1); bad.zero(); u32 *bad_ = bad.p;
		const u32 *ep1 = in.ep1, *ep2 = in.ep2;
		parallel_for<P_check>(c, E, PV_LAMBDA(size_t e) { if (lab_[ep1[e] >> 1] != lab_[ep2[e] >> 1]) atomic_add(bad_, 1u); });
		parallel_for<P_check>(c, V, PV_LAMBDA(size_t v) { if (lab_[lab_[v]] != lab_[v]) atomic_add(bad_, 1u); });
		const u32 nb = bad.read(0);
		if (nb) throw std::runtime_error("shard: inconsistent component labels (" + std::to_string(nb) + " links across components or labels that are not roots)");
	}
	// ---- weights: links + path steps + segments per component
	Buf<u64> w(c, nC); w.zero();
	u64 *w_ = w.p;
	{
		shard_weights(c, E, CompOfPort{in.ep1, lab_, compIdx_}, w_);
		if (in.steps) shard_weights(c, S, CompOfPort{in.steps, lab_, compIdx_}, w_);
		shard_weights(c, V, CompOfVertex{lab_, compIdx_}, w_);
	}
	out.weight.resize(nC);
	c.d2h(out.weight.data(), w.p, (size_t)nC * 8);
	// walks on the host: the component of a path is the component of its first step (one small gather), its weight its length
	std::vector<u32> pathCompH;
	if (!in.steps) {
		if (S && !in.h_steps) throw std::runtime_error("shard: no walks given");
		std::vector<u32> firstV((size_t)H, NONE);
		for (u32 h = 0; h < H; h++) if (in.h_stepOff[h + 1] > in.h_stepOff[h]) firstV[h] = in.h_steps[in.h_stepOff[h]] >> 1;
		Buf<u32> fv(c, H ? H : 1), pc(c, H ? H : 1);
		c.h2d(fv.p, firstV.data(), (size_t)H * 4);
		const u32 *fv_ = fv.p; u32 *pc_ = pc.p;
		parallel_for<P_step>(c, H, PV_LAMBDA(size_t h) { u32 v = fv_[h]; pc_[h] = v == NONE ? NONE : compIdx_[lab_[v]]; });
		pathCompH = pc.to_host(); pathCompH.resize(H);
		for (u32 h = 0; h < H; h++) if (pathCompH[h] != NONE) out.weight[pathCompH[h]] += in.h_stepOff[h + 1] - in.h_stepOff[h];
	}
	tr.mark("labels + weights");
	// ---- the one host step: greedy longest-processing-time over nC numbers
	std::vector<u32> order(nC);
	for (u32 k = 0; k < nC; k++) order[k] = k;
	std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return out.weight[a] > out.weight[b]; });
	std::vector<u64> load(world, 0);
	out.owner.assign(nC, 0);
	for (u32 k : order) { u32 best = 0; for (u32 r = 1; r < world; r++) if (load[r] < load[best]) best = r; out.owner[k] = best; load[best] += out.weight[k]; }
	out.n_comp_total = nC;
	out.comp_global.clear();
	std::vector<u32> mine(nC, 0);
	for (u32 k = 0; k < nC; k++) if (out.owner[k] == rank) { out.comp_global.push_back(k + 1); mine[k] = 1; }
	Buf<u32> mineD(c, nC);
	c.h2d(mineD.p, mine.data(), (size_t)nC * 4);
	const u32 *mine_ = mineD.p;
	// ---- segments
	Buf<u32> keepV(c, (size_t)V + 1), remap(c, (size_t)V + 1);
	u32 *keepV_ = keepV.p; const u32 *remap_ = remap.p;
	parallel_for<P_keep>(c, (size_t)V + 1, PV_LAMBDA(size_t v) { keepV_[v] = v < V ? mine_[compIdx_[lab_[v]]] : 0u; });
	exclusive_scan(c, keepV.p, remap.p, (size_t)V + 1, OpAddU32(), 0u);
	// ---- links and steps: keep flags, positions
	Buf<u32> keepE(c, E + 1), posE(c, E + 1), keepS, posS;
	u32 *keepE_ = keepE.p; const u32 *posE_ = posE.p;
	{ const u32 *ep1 = in.ep1; parallel_for<P_keep>(c, E + 1, PV_LAMBDA(size_t e) { keepE_[e] = e < E ? keepV_[ep1[e] >> 1] : 0u; }); }
	exclusive_scan(c, keepE.p, posE.p, E + 1, OpAddU32(), 0u);
	u32 nV = 0, nE = 0, nS = 0;
	out.h_stepOff.assign((size_t)H + 1, 0);
	Buf<u32> rawOwn;      // host-walk mode: the owned paths' steps as uploaded (whole-graph vertex indices)
	if (in.steps) {
		keepS = Buf<u32>(c, S + 1); posS = Buf<u32>(c, S + 1);
		u32 *keepS_ = keepS.p; const u32 *posS_ = posS.p; const u32 *steps = in.steps;
		parallel_for<P_keep>(c, S + 1, PV_LAMBDA(size_t s) { keepS_[s] = s < S ? keepV_[steps[s] >> 1] : 0u; });
		exclusive_scan(c, keepS.p, posS.p, S + 1, OpAddU32(), 0u);
		// sizes + the new path offsets in ONE host wait
		Buf<u32> offIn(c, (size_t)H + 1), offOut(c, (size_t)H + 1);
		c.h2d(offIn.p, in.h_stepOff, ((size_t)H + 1) * 4);
		{ const u32 *oi = offIn.p; u32 *oo = offOut.p; parallel_for<P_step>(c, (size_t)H + 1, PV_LAMBDA(size_t h) { oo[h] = posS_[oi[h]]; }); }
		{ MultiRead mr(c); mr.add(&nV, remap.p + V, 4); mr.add(&nE, posE.p + E, 4); mr.add(&nS, posS.p + S, 4); mr.wait(); }
		c.d2h(out.h_stepOff.data(), offOut.p, ((size_t)H + 1) * 4);
	} else {
		u64 tot = 0;
		for (u32 h = 0; h < H; h++) { out.h_stepOff[h] = (u32)tot; if (pathCompH[h] != NONE && mine[pathCompH[h]]) tot += in.h_stepOff[h + 1] - in.h_stepOff[h]; }
		out.h_stepOff[H] = (u32)tot; nS = (u32)tot;
		rawOwn = Buf<u32>(c, nS ? nS : 1);
		for (u32 h = 0; h < H; h++) {   // one copy per owned path, straight from the page-locked walks
			const u32 len = out.h_stepOff[h + 1] - out.h_stepOff[h];
			if (len) c.h2d(rawOwn.p + out.h_stepOff[h], in.h_steps + in.h_stepOff[h], (size_t)len * 4);
		}
		{ MultiRead mr(c); mr.add(&nV, remap.p + V, 4); mr.add(&nE, posE.p + E, 4); mr.wait(); }
	}
	out.V = nV; out.E = nE; out.S = nS;
	tr.mark("flags + scans");
	// ---- compaction
	out.ids = Buf<u32>(c, nV ? nV : 1); out.segLen = Buf<u32>(c, nV ? nV : 1); out.ep1 = Buf<u32>(c, nE ? nE : 1); out.ep2 = Buf<u32>(c, nE ? nE : 1);
	out.steps = Buf<u32>(c, nS ? nS : 1);
	{
		u32 *oi = out.ids.p, *ol = out.segLen.p; const u32 *ii = in.ids, *il = in.segLen;
		parallel_for<P_vertex>(c, V, PV_LAMBDA(size_t v) { if (keepV_[v]) { u32 nv = remap_[v]; oi[nv] = ii[v]; ol[nv] = il[v]; } });
		u32 *o1 = out.ep1.p, *o2 = out.ep2.p; const u32 *ep1 = in.ep1, *ep2 = in.ep2;
		parallel_for<P_link>(c, E, PV_LAMBDA(size_t e) {
			if (!keepE_[e]) return;
			u32 a = ep1[e], b = ep2[e]; u32 p = posE_[e];
			o1[p] = (remap_[a >> 1] << 1) | (a & 1u); o2[p] = (remap_[b >> 1] << 1) | (b & 1u);
		});
		u32 *os = out.steps.p;
		if (in.steps) {
			const u32 *steps = in.steps, *keepS_ = keepS.p, *posS_ = posS.p;
			parallel_for<P_step>(c, S, PV_LAMBDA(size_t s) { if (keepS_[s]) { u32 x = steps[s]; os[posS_[s]] = (remap_[x >> 1] << 1) | (x & 1u); } });
		} else {
			const u32 *raw = rawOwn.p; Buf<u32> bad(c, 1); bad.zero(); u32 *bad_ = bad.p;
			parallel_for<P_step>(c, nS, PV_LAMBDA(size_t s) { u32 x = raw[s]; if (!keepV_[x >> 1]) atomic_add(bad_, 1u); os[s] = (remap_[x >> 1] << 1) | (x & 1u); });
			if (bad.read(0)) throw std::runtime_error("shard: a path leaves the component of its first step");
		}
	}
	if (in.lblOff && in.lbl) {
		Buf<u64> len(c, (size_t)nV + 1);
		out.lblOff = Buf<u64>(c, (size_t)nV + 1);
		u64 *len_ = len.p; const u32 *ol = out.segLen.p;
		parallel_for<P_label>(c, (size_t)nV + 1, PV_LAMBDA(size_t v) { len_[v] = v < nV ? (u64)ol[v] : 0ull; });
		exclusive_scan(c, len.p, out.lblOff.p, (size_t)nV + 1, OpAddU64(), (u64)0);
		u64 nb = 0; c.d2h(&nb, out.lblOff.p + nV, 8);
		out.lbl = Buf<u8>(c, nb ? nb : 1);
		u8 *ob = out.lbl.p; const u64 *oo = out.lblOff.p, *io = in.lblOff; const u8 *ib = in.lbl;
		parallel_for<P_label>(c, V, PV_LAMBDA(size_t v) {
			if (!keepV_[v]) return;
			u64 src = io[v], dst = oo[remap_[v]]; u64 m = io[v + 1] - src;
			for (u64 k = 0; k < m; k++) ob[dst + k] = ib[src + k];
		});
	}
	out.h_ids.resize(nV); out.h_segLen.resize(nV);
	c.d2h_async(out.h_ids.data(), out.ids.p, (size_t)nV * 4); c.d2h_async(out.h_segLen.data(), out.segLen.p, (size_t)nV * 4);
	c.sync();
	tr.mark("compaction");
}

} // namespace pv
