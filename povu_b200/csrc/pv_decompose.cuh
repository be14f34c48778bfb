// povu_b200 -- decompose: bidirected graph -> flubbles + PVST nesting, every step a data-parallel
// kernel (no host graph traversal).  Replaces, with identical results, the reference's
//   VG::componetize            /root/reference/src/povu/graph/bidirected.cpp:477-602
//   pst::Tree::from_bd         /root/reference/src/povu/graph/spanning_tree.cpp:262-463
//   simple_cycle_equiv         /root/reference/src/povu/algorithms/flubbles.cpp:503-719
//   compute_eq_class_stack     flubbles.cpp:412-501 (+ tree_utils.cpp:19-157)
//   compute_eq_class_metadata  flubbles.cpp:375-410
//   add_flubbles               flubbles.cpp:295-367
//
// Pipeline (DESIGN.md has the full derivation):
//  D1  CSR of the bi-edged graph whose per-port adjacency order equals the reference's per-component
//      renumbered edge order (two stable radix sorts).
//  D2  The reference's lexicographic DFS tree, reproduced exactly but in parallel: rounds of
//      {union-find over unvisited segments -> per-component claim by the deepest visited neighbour
//       (64-bit atomicMax) -> "first unvisited edge" functional graph -> pointer-doubling path
//       marking with rho/self-intersection truncation}.  A chain of bubbles needs 2-3 rounds.
//  D3  Euler tour + list ranking -> preorder numbers and subtree sizes.
//  D4  back edges (parallel edges collapse, one back edge per self-looped segment, tips -> dummy root),
//      bridges -> "simplifying" edges, 128-bit XOR labels, prefix XOR -> class key per black edge,
//      exact bracket-set verification of every merged pair (count + range-minimum test).
//  D5  class-stack positions from subtree counts (one scan), radix sort by key -> next/prev seen.
//  D6  flubbles + PVST parents: laminar components via a max-plus function scan and previous-smaller
//      queries; non-laminar components replay the reference's stack machine (one thread each).
#pragma once
#include "pv_prims.cuh"
#include "pv_refnum.cuh"

namespace pv {

struct DecompOut {
	// per flubble, in creation order (PVST vertex idx = local position + 1)
	hvec<u32> comp;      // 0-based component index (component id = comp + 1)
	hvec<u32> a_v, z_v;  // vertex *indices*
	hvec<u8> a_fwd, z_fwd;
	hvec<u32> parent;    // PVST idx within the component (0 = dummy root)
	hvec<u32> depth;     // PVST depth (root = 0)
	hvec<u32> fl_base;   // index of the first flubble of the flubble's component
	// per component
	std::vector<u32> comp_root, comp_nv, comp_nfl, comp_has_tips, comp_nonlaminar;
	// optional debug / parity outputs
	std::vector<u32> stack_v;       // class stack: vertex idx per position (components concatenated)
	std::vector<u32> stack_class;   // dense class id per position
	std::vector<u32> stack_fwd;
	std::vector<u32> stack_next;    // next stack position of the same class (the position itself when there is none), components concatenated
	u32 dfs_rounds = 0;
	u32 arc_rounds = 0;      // rounds of the fired-arc fixpoint (0 = every component laminar; > 64 = the sequential replay ran)
	u32 collisions = 0;
	u32 faithful_classes = 0;   // > 0: the classes were re-derived the reference's way (vertices with >= 3 children reaching above them; pv_refnum.cuh)
	void reset() {   // keeps the (page-locked) capacity of the big arrays
		comp.clear(); a_v.clear(); z_v.clear(); a_fwd.clear(); z_fwd.clear(); parent.clear(); depth.clear(); fl_base.clear();
		comp_root.clear(); comp_nv.clear(); comp_nfl.clear(); comp_has_tips.clear(); comp_nonlaminar.clear();
		stack_v.clear(); stack_class.clear(); stack_fwd.clear(); stack_next.clear(); dfs_rounds = 0; arc_rounds = 0; collisions = 0; faithful_classes = 0;
	}
};

struct DecompOpts {
	bool want_stack = false; u64 seed = 0x706F7675ull;
	Buf<u32> *keep_labels = nullptr;   // receives label[v] = lowest vertex index of v's component (what pv_call's component_labels computes)
	struct DevForest *keep_forest = nullptr;   // receives the flubble table on the device (call reads it there instead of re-uploading)
	bool async_results = false;        // the big result arrays travel on the context's copy stream; the caller must c.bg_sync() before reading them
	RefNumOut *refnum = nullptr;       // also derive the reference's internal numbering (pv_refnum.cuh: --hairpins / --structure-export only)
};
// the flubble table as decompose left it in HBM
struct DevForest { Buf<u32> a_v, z_v, parent, depth, comp, fl_base; Buf<u8> a_fwd, z_fwd; u32 n = 0; bool valid = false; };

// kernel tags with their algorithmic bytes per logical item
PV_KERNEL(K_minport, 16);      // 2 endpoints in, 2 degree atomics
PV_KERNEL(K_csr_fill, 24);
PV_KERNEL(K_rev, 12);
PV_KERNEL(K_tip, 12);
PV_KERNEL(K_uf_init, 20);
PV_KERNEL(K_uf_union, 24);     // per port: offsets + ~2 neighbours + visited + uf words
PV_KERNEL(K_uf_flatten, 12);
PV_KERNEL(K_uf_min, 12);
PV_KERNEL(K_start, 12);
PV_KERNEL(K_claim, 28);
PV_KERNEL(K_entry, 24);
PV_KERNEL(K_first_edge, 28);
PV_KERNEL(K_skip, 40);          // per port: its candidates + a bounded walk of the first-edge chain from each
PV_KERNEL(K_skipcheck, 28);
PV_KERNEL(K_mark, 24);         // J[x], J[J[x]], rnk[x] in, J'[x] out (+ the marks)
PV_KERNEL(K_segconf, 12);
PV_KERNEL(K_finalize, 36);
PV_KERNEL(K_visit, 16);
PV_KERNEL(K_active, 12);
PV_KERNEL(K_pack, 24);
PV_KERNEL(K_compcount, 12);
PV_KERNEL(K_treeslot, 8);
PV_KERNEL(K_euler, 48);
PV_KERNEL(K_pre, 32);
PV_KERNEL(K_backedge, 64);
PV_KERNEL(K_bridge, 24);
PV_KERNEL(K_simp, 12);
PV_KERNEL(K_stackw, 40);
PV_KERNEL(K_stackpos, 72);
PV_KERNEL(K_seen, 44);
PV_KERNEL(K_verify, 32);
PV_KERNEL(K_arc, 16);
PV_KERNEL(K_lam, 16);
PV_KERNEL(K_fn, 24);
PV_KERNEL(K_flub, 28);
PV_KERNEL(K_parent, 16);
PV_KERNEL(K_seq, 16);
PV_KERNEL(K_iota, 4);
PV_KERNEL(K_word, 24);

// The reference's stage-cost trace (POVU_STAGE_COST_TRACE=1, /root/reference/src/povu/common/stage_cost.cpp:59-94): six named
// decompose stages with calls / input items / output items / elapsed ns, printed to stderr at exit in the same line format, so
// CPU and GPU stage times line up.  Here the stage boundaries are CUDA events on the pipeline's stream; the mapping:
//   componentDecomposition     D1 (CSR in reference edge order) + dense component tables   [VG::componetize]
//   tipDummyAugmentation       D2 + D3: the exact DFS tree, Euler tour, pre / post order    [pst::Tree::from_bd carries this stage]
//   cycleClassAssignment       D4: back edges, bridges, labels, class keys                  [simple_cycle_equiv]
//   traversalFrameConstruction D5a: class-stack positions                                    [compute_eq_class_stack]
//   boundaryEmission           D5b: sort by class key, next / previous seen, exact check    [compute_eq_class_metadata]
//   hierarchyConstruction      D6: flubbles, PVST depths and parents                        [add_flubbles]
// calls counts components (the reference runs every stage but the first once per component).
struct StageCost {
	static const int N = 6;
	u64 calls[N] = {0}, in_items[N] = {0}, out_items[N] = {0}; double ns[N] = {0};
	bool on = false, registered = false;
	static StageCost &inst() { static StageCost s; return s; }
	static bool enabled() { const char *v = getenv("POVU_STAGE_COST_TRACE"); return v && v[0] && v[0] != '0'; }
	static void report() {
		StageCost &s = inst();
		static const char *contract[N] = {"componentDecomposition", "tipDummyAugmentation", "traversalFrameConstruction", "cycleClassAssignment", "boundaryEmission", "hierarchyConstruction"};
		static const char *trace[N] = {"component-decomposition", "tip-dummy-augmentation", "traversal-frame-construction", "cycle-class-assignment", "flubble-boundary-extraction", "flubble-hierarchy-construction"};
		fprintf(stderr, "povu-stage-cost-trace schema=povu.stage-cost.v1\n");
		for (int i = 0; i < N; i++)
			fprintf(stderr, "povu-stage-cost contract=%s trace=%s calls=%llu input_items=%llu output_items=%llu elapsed_ns=%llu\n", contract[i], trace[i],
				(unsigned long long)s.calls[i], (unsigned long long)s.in_items[i], (unsigned long long)s.out_items[i], (unsigned long long)s.ns[i]);
	}
	void add(int stage, u64 c, u64 in, u64 out, double nanos) {
		if (!registered) { registered = true; atexit(report); }
		calls[stage] += c; in_items[stage] += in; out_items[stage] += out; ns[stage] += nanos;
	}
};
// stage boundaries of one decompose call (events only exist when the trace is on)
struct StageMarks {
	Ctx &c; bool on; int n = 0;
#ifndef PV_EMU
	cudaEvent_t ev[12];
#else
	double t[12];
#endif
	explicit StageMarks(Ctx &ctx) : c(ctx), on(StageCost::enabled()) {}
	void mark() {
		if (!on || n >= 12) return;
#ifndef PV_EMU
		PV_CUDA_CHECK(cudaEventCreate(&ev[n])); PV_CUDA_CHECK(cudaEventRecord(ev[n], c.stream));
#else
		t[n] = Trace::now();
#endif
		n++;
	}
	double ns_between(int a, int b) {   // after the stream was synchronised
#ifndef PV_EMU
		float ms = 0; cudaEventElapsedTime(&ms, ev[a], ev[b]); return (double)ms * 1e6;
#else
		return (t[b] - t[a]) * 1e6;
#endif
	}
	~StageMarks() {
#ifndef PV_EMU
		for (int i = 0; i < n; i++) cudaEventDestroy(ev[i]);
#endif
	}
};

struct MP { i32 a, b; };   // x -> max(x + a, b)
struct OpMP {
	PV_HD MP operator()(MP f, MP g) const {   // apply f first, then g
		const i32 BIG = 1 << 29;
		i32 a = f.a + g.a; if (a < -BIG) a = -BIG;
		i32 b1 = f.b + g.a; i32 b = b1 > g.b ? b1 : g.b;
		return MP{a, b};
	}
};

// lower (raw) port of the link that entry x = 2e + end belongs to: the major key of the reference's renumbered edge order
struct EntryMinPort { const u32 *ep1, *ep2; PV_DEV u32 operator()(size_t x) const { u32 a = ep1[x >> 1], b = ep2[x >> 1]; return a < b ? a : b; } };
// port of adjacency entry x: entries 2e / 2e+1 are the two ends of link e (file order)
struct EntryPort {
	const u32 *ep1, *ep2;
	PV_DEV u32 operator()(size_t j) const {
		u32 e = (u32)(j >> 1); u32 a = ep1[e], b = ep2[e];
		if ((a >> 1) == (b >> 1)) { a = a & ~1u; b = a | 1u; }   // any self-loop becomes left--right (bidirected.cpp:66-77,526-531)
		return (j & 1u) ? b : a;
	}
};

PV_DEV u32 uf_find(u32 *uf, u32 x) {
	u32 p = ld_volatile(uf + x);
	while (p != x) {
		u32 gp = ld_volatile(uf + p);
		if (gp != p) uf[x] = gp;   // path halving (benign race: always points to an ancestor)
		x = p; p = gp;
	}
	return x;
}
PV_DEV void uf_union(u32 *uf, u32 a, u32 b) {
	while (true) {
		a = uf_find(uf, a); b = uf_find(uf, b);
		if (a == b) return;
		if (a < b) { u32 t = a; a = b; b = t; }     // hook the larger root under the smaller: root = min index
		if (atomic_cas(uf + a, a, b) == a) return;
	}
}
// Hooking by a hashed priority instead of by index: on chain-shaped graphs (vertex ids ascending along the backbone) hooking
// by index first builds one long list that every later find has to halve; a random priority keeps the trees shallow.  The
// root is then an arbitrary member of the component (callers that need the lowest vertex index fix it up afterwards).
// (a bijective 32-bit mixer: priorities are unique without a tie-break word, and cost five integer instructions -- the hooking
// loops below are instruction-bound, the 64-bit splitmix here was most of their time)
PV_DEV u32 uf_priority(u32 x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
PV_DEV void uf_union_hashed(u32 *uf, u32 a, u32 b) {
	while (true) {
		a = uf_find(uf, a); b = uf_find(uf, b);
		if (a == b) return;
		if (uf_priority(a) > uf_priority(b)) { u32 t = a; a = b; b = t; }     // a = lower priority: it goes under b
		if (atomic_cas(uf + a, a, b) == a) return;
	}
}

// Round 0 of the DFS labels the components of the WHOLE graph.  Segment ids of real graphs follow the backbone, so most edges
// join two segments a few indices apart: a CTA first unites its own span of UF_SPAN consecutive segments in shared memory
// (the same hashed hooking, shared-memory atomics), writes every segment's local root, and the global pass that follows
// only sees the edges that leave a span.  Without locality the pre-pass finds nothing to join and costs one sweep.
PV_KERNEL(K_uf_local, 28);
#ifndef PV_EMU
static const int UF_SPAN = 2048, UF_THREADS = 512;
__device__ __forceinline__ u32 ufl_find(volatile u32 *s, u32 x) {
	u32 p = s[x];
	while (p != x) { u32 gp = s[p]; if (gp != p) s[x] = gp; x = p; p = gp; }
	return x;
}
__global__ void __launch_bounds__(UF_THREADS) uf_local_kernel(const u32 *adjOff, const u32 *adjNbr, u32 V, u32 *uf) {
	__shared__ u32 s[UF_SPAN];
	pdl_wait();
	const u32 base = blockIdx.x * UF_SPAN;
	for (u32 l = threadIdx.x; l < UF_SPAN; l += UF_THREADS) s[l] = l;
	__syncthreads();
	for (u32 t = threadIdx.x; t < 2 * UF_SPAN; t += UF_THREADS) {      // one port per step
		const u32 l = t >> 1, v = base + l;
		if (v >= V) continue;
		const u32 p = 2 * v + (t & 1u);
		for (u32 e = adjOff[p], ee = adjOff[p + 1]; e < ee; e++) {
			const u32 w = adjNbr[e] >> 1;
			if (w >= v || w < base) continue;      // each edge inside the span once, from its higher end
			u32 a = l, b = w - base;
			while (true) {
				a = ufl_find(s, a); b = ufl_find(s, b);
				if (a == b) break;
				if (uf_priority(base + a) > uf_priority(base + b)) { u32 x = a; a = b; b = x; }
				if (atomicCAS(s + a, a, b) == a) break;
			}
		}
	}
	__syncthreads();
	for (u32 l = threadIdx.x; l < UF_SPAN; l += UF_THREADS) if (base + l < V) uf[base + l] = base + ufl_find(s, l);
}
#endif

// Which class re-encounters still find their previous occurrence on the reference's stack?  (see D6 in decompose_device)
// prv[i] = previous stack position of i's class or NONE.  fire[i]: 1 = the arc (prv[i], i) fires, 2 = dead or no arc.
// Returns the number of fixpoint rounds after the laminarity pass (0 = no two arcs cross), -1 if undecided arcs remain after
// 64 rounds (fire[] then holds 0 for them).  segComp / lamBad (optional): component of every position -> "has crossing arcs".
// extra_src / extra_dst, extra2_src / extra2_dst (optional): more device words read with the first host wait (the caller's
// flubble count, its "reference class naming may differ" flag).
inline int fired_arcs(Ctx &c, const u32 *prv_, u32 nS, u32 *fire_, const u32 *segComp_, u32 *lamBad_, const u32 *extra_src, u32 *extra_dst, const u32 *extra2_src = nullptr, u32 *extra2_dst = nullptr) {
	Buf<u32> arcA(c, nS ? nS : 1), fl(c, 1);
	u32 *arcA_ = arcA.p, *fl_ = fl.p;
	fl.zero();
	parallel_for<K_arc>(c, nS, PV_LAMBDA(size_t i) { arcA_[i] = prv_[i]; });   // NONE = no arc = +inf
	{
		Rmq rmqA;
		rmq_build(c, rmqA, arcA.p, nS);
		RmqView rv = rmqA.view();
		parallel_for<K_lam>(c, nS, PV_LAMBDA(size_t i) {
			u32 p = prv_[i];
			if (p == NONE) { fire_[i] = 2; return; }
			bool crossed = (u32)i > p + 1 && rv.query(p + 1, (u32)i - 1) < p;
			fire_[i] = crossed ? 0u : 1u;
			if (crossed) { fl_[0] = 1; if (lamBad_) lamBad_[segComp_[i]] = 1; }
		});
	}
	u32 undecided = 0;
	{ MultiRead mr(c); if (extra_src) mr.add(extra_dst, extra_src, 4); if (extra2_src) mr.add(extra2_dst, extra2_src, 4); mr.add(&undecided, fl.p, 4); mr.wait(); }
	int rounds = 0;
	if (undecided) {
		Buf<u32> arcF(c, nS);
		u32 *arcF_ = arcF.p;
		while (undecided) {
			if (++rounds > 64) return -1;
			parallel_for<K_arc>(c, nS, PV_LAMBDA(size_t i) { u32 st = fire_[i]; u32 p = prv_[i]; arcA_[i] = st == 2 ? NONE : p; arcF_[i] = st == 1 ? p : NONE; });
			Rmq rmqA, rmqF;
			rmq_build(c, rmqA, arcA.p, nS); rmq_build(c, rmqF, arcF.p, nS);
			RmqView ra = rmqA.view(), rf = rmqF.view();
			fl.zero();
			parallel_for<K_lam>(c, nS, PV_LAMBDA(size_t i) {
				if (fire_[i] != 0) return;
				u32 p = prv_[i];
				if (ra.query(p + 1, (u32)i - 1) >= p) { fire_[i] = 1; return; }      // nothing left that could cross it
				if (rf.query(p + 1, (u32)i - 1) < p) { fire_[i] = 2; return; }       // crossed by an arc that fires: p was popped
				fl_[0] = 1;
			});
			undecided = fl.read(0);
		}
	}
	return rounds;
}

inline void decompose_device(Ctx &c, u32 V, u64 E, const u32 *ep1, const u32 *ep2, const DecompOpts &opt, DecompOut &out) {
	out.reset();
	if (V == 0) return;
	const u32 P = 2 * V;            // ports: 2*v + side (0 = left, 1 = right)
	const u64 M = 2 * E;            // adjacency entries
	if (M >= 0xFFFFFFF0ull) throw std::runtime_error("graph too large for 32-bit adjacency slots");

	Trace tr(c, "decompose");
	StageMarks sm(c);
	sm.mark();   // 0
	// ------------------------------------------------------------------ D1: CSR in reference edge order
	Buf<u32> adjOff(c, P + 1), adjNbr(c, M ? M : 1), adjOwner(c, M ? M : 1), rev(c, M ? M : 1), origDeg(c, P);
	{
		// The adjacency of a port lists its links in the reference's per-component RENUMBERED edge order = (lower port of the
		// link, file order).  Entry x = 2e + end of link e therefore sorts, inside its port's group, by (min port of e, x): one
		// grouping of the 2E entries with that secondary key (count / slot / per-group insertion sort; hub ports take the two-key
		// radix sort) -- the renumbered edge index itself is never needed.
		origDeg.zero();
		u32 *origDeg_ = origDeg.p;
		parallel_for<K_minport>(c, E, PV_LAMBDA(size_t e) { atomic_add(origDeg_ + ep1[e], 1u); atomic_add(origDeg_ + ep2[e], 1u); });
		Buf<u32> list(c, M ? M : 1), slotOf(c, M ? M : 1);
		const EntryPort EP{ep1, ep2};
		group_indices_ord<EntryPort, EntryMinPort, true>(c, EP, EntryMinPort{ep1, ep2}, M, P, adjOff.p, list.p);
		u32 *slotOf_ = slotOf.p, *adjNbr_ = adjNbr.p, *adjOwner_ = adjOwner.p, *rev_ = rev.p; const u32 *list_ = list.p;
		parallel_for<K_csr_fill>(c, M, PV_LAMBDA(size_t s) {
			u32 j = list_[s];
			adjNbr_[s] = EP(j ^ 1u); adjOwner_[s] = EP(j); slotOf_[j] = (u32)s;
		});
		parallel_for<K_rev>(c, M, PV_LAMBDA(size_t s) { rev_[s] = slotOf_[list_[s] ^ 1u]; });
	}
	const u32 *adjOff_ = adjOff.p, *adjNbr_ = adjNbr.p, *adjOwner_ = adjOwner.p, *rev_ = rev.p, *origDeg_ = origDeg.p;

	tr.mark("D1 csr");
	sm.mark();   // 1
	// ------------------------------------------------------------------ D2: exact DFS tree by rounds
	Buf<u32> visited(c, V), uf(c, V), comp(c, V), gcomp(c, V), entrySide(c, V);
	Buf<u32> parent(c, P), pslot(c, P), depth(c, P), F(c, P), Fslot(c, P), J(c, P), J2(c, P), rnk(c, P);
	Buf<u64> claim(c, V);
	Buf<u32> trunc(c, V), entDepth(c, V), entV(c, V), tipPort(c, V), compStart(c, V), flag(c, 4);
	visited.zero(); parent.fill_ff(); pslot.fill_ff(); depth.zero();
	u32 *visited_ = visited.p, *uf_ = uf.p, *comp_ = comp.p, *gcomp_ = gcomp.p, *entrySide_ = entrySide.p;
	u32 *parent_ = parent.p, *pslot_ = pslot.p, *depth_ = depth.p, *F_ = F.p, *Fslot_ = Fslot.p, *rnk_ = rnk.p;
	u64 *claim_ = claim.p;
	u32 *trunc_ = trunc.p, *entDepth_ = entDepth.p, *entV_ = entV.p, *tipPort_ = tipPort.p, *compStart_ = compStart.p, *flag_ = flag.p;

	// tips of the WHOLE graph (from_gfa.cpp:262-277), decided on the original (un-rewritten) endpoints
	parallel_for<K_tip>(c, V, PV_LAMBDA(size_t v) {
		u32 dl = origDeg_[2 * v], dr = origDeg_[2 * v + 1];
		tipPort_[v] = dl == 0 ? (u32)(2 * v) : (dr == 0 ? (u32)(2 * v + 1) : NONE);
	});

	Buf<u32> active(c, V);
	u32 nActive = V;
	{ u32 *a = active.p; parallel_for<K_iota>(c, V, PV_LAMBDA(size_t i) { a[i] = (u32)i; }); }

	// dense component tables: the flags and their compaction are queued in round 0 (as soon as the labels exist) so that the
	// component count arrives with that round's own host wait
	Buf<u32> isRoot(c, V), compIdxOfRoot(c, V + 1), nvOfRoot(c, V), roots;
	u32 nComp = 0;
	u32 round = 0;
	bool repair = false;   // set by a round whose stripped path was cut short (see f')
	static const u32 skip_walk = getenv("POVU_B200_SKIP_WALK") ? (u32)strtoul(getenv("POVU_B200_SKIP_WALK"), nullptr, 10) : 32u;   // 0 = never predict
	while (nActive > 0) {
		const u32 *act = active.p;
		const u32 nA = nActive;
		MultiRead mrRound(c);
#ifdef PV_EMU
		if (getenv("PV_DEBUG")) { fprintf(stderr, "[dfs] round %u active %u:", round, nA); for (u32 i = 0; i < nA && i < 32; i++) fprintf(stderr, " %u", act[i]); fprintf(stderr, "\n"); }
#endif
		// (a,b,c) connected components of the still-unvisited part (segment granularity)
		parallel_for<K_uf_init>(c, nA, PV_LAMBDA(size_t i) { u32 v = act[i]; uf_[v] = v; claim_[v] = 0; trunc_[v] = NONE; });
		u32 span = 0;      // > 0: segments of one span of `span` consecutive indices are already united (round 0, see uf_local_kernel)
#ifndef PV_EMU
		static const bool no_local = getenv("POVU_B200_UF_NO_LOCAL") != nullptr;
		if (round == 0 && !no_local && V >= 4u * UF_SPAN) {
			span = UF_SPAN;
			c.prof_begin(K_uf_local::name(), V, (size_t)V * K_uf_local::bytes);
			PV_CUDA_CHECK(launch_pdl(uf_local_kernel, dim3((unsigned)((V + UF_SPAN - 1) / UF_SPAN)), dim3(UF_THREADS), 0, c.stream, adjOff_, adjNbr_, V, uf_));
			c.prof_end();
			c.launches++;
		}
#endif
		parallel_for<K_uf_union>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
			u32 v = act[i >> 1]; u32 p = 2 * v + (u32)(i & 1);
			for (u32 s = adjOff_[p], e = adjOff_[p + 1]; s < e; s++) {
				u32 w = adjNbr_[s] >> 1;
				if (span && w / span == v / span) continue;
				if (w < v && !visited_[w]) uf_union_hashed(uf_, v, w);   // an edge between two unvisited segments is seen from both ends: join once
			}
		});
		// a few rounds of plain pointer doubling first (every store is an ancestor, so they can run in place): the finds of the
		// flatten then start from short chains instead of chasing them one dependent load at a time
		for (int jmp = 0; jmp < 3; jmp++)
			parallel_for<K_uf_flatten>(c, nA, PV_LAMBDA(size_t i) { u32 v = act[i]; u32 p = ld_volatile(uf_ + v); u32 gp = ld_volatile(uf_ + p); if (gp != p) uf_[v] = gp; });
		parallel_for<K_uf_flatten>(c, nA, PV_LAMBDA(size_t i) { u32 v = act[i]; comp_[v] = uf_find(uf_, v); });
		if (round == 0) {
			// round 0's labels name the whole-graph components and must be the LOWEST vertex index of each (component order =
			// ascending lowest vertex, bidirected.cpp:585-596): representative -> minimum.  Runs of 32 consecutive vertices mostly
			// share a representative, so a thread issues one atomicMin per run.
			Buf<u32> minIdx(c, V);
			minIdx.fill_ff();
			u32 *minIdx_ = minIdx.p;
			parallel_for<K_uf_min>(c, ((size_t)V + 31) / 32, PV_LAMBDA(size_t t) {
				u32 v0 = (u32)t * 32, v1 = v0 + 32 < V ? v0 + 32 : V;
				u32 rep = comp_[v0], lo = v0;
				for (u32 v = v0 + 1; v < v1; v++) { u32 r = comp_[v]; if (r != rep) { atomic_min(minIdx_ + rep, lo); rep = r; lo = v; } }
				atomic_min(minIdx_ + rep, lo);
			});
			parallel_for<K_uf_min>(c, V, PV_LAMBDA(size_t v) { comp_[v] = minIdx_[comp_[v]]; });
			c.d2d(gcomp.p, comp.p, (size_t)V * 4);
			{
				nvOfRoot.zero();
				u32 *isRoot_ = isRoot.p, *nvOfRoot_ = nvOfRoot.p;
				parallel_for<K_compcount>(c, V, PV_LAMBDA(size_t v) { isRoot_[v] = comp_[v] == v ? 1u : 0u; atomic_inc_aggregated(nvOfRoot_, comp_[v]); });
				compact_flags(c, isRoot.p, V, roots, &compIdxOfRoot, &mrRound, &nComp);   // (the compaction's own scan is the dense component index)
			}
			// start port: smallest tip (v_id, end) of the component (types.cpp:60-68), else (lowest vertex, left)
			parallel_for<K_start>(c, V, PV_LAMBDA(size_t v) { compStart_[v] = NONE; rnk_[2 * v] = NONE; rnk_[2 * v + 1] = NONE; });
			parallel_for<K_start>(c, V, PV_LAMBDA(size_t v) { u32 t = tipPort_[v]; if (t != NONE) atomic_min(compStart_ + comp_[v], t); });
			parallel_for<K_entry>(c, V, PV_LAMBDA(size_t v) {
				if (comp_[v] != v) return;
				u32 q0 = compStart_[v];
				if (q0 == NONE) q0 = (u32)(2 * v);
				compStart_[v] = q0;
				rnk_[q0] = 0; entDepth_[v] = 1; entV_[v] = q0 >> 1;          // virtual root has depth 0
				parent_[q0] = NONE; pslot_[q0] = NONE;   // parent = the component's virtual root
			});
		} else {
			// (d) every unvisited component is claimed by its deepest visited neighbour, lowest edge first
			parallel_for<K_claim>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
				u32 v = act[i >> 1]; u32 q = 2 * v + (u32)(i & 1);
				rnk_[q] = NONE;
				u64 best = 0;
				for (u32 s = adjOff_[q], e = adjOff_[q + 1]; s < e; s++) {
					u32 p = adjNbr_[s];
					if (!visited_[p >> 1]) continue;
					u64 key = ((u64)depth_[p] << 32) | (u64)(0xFFFFFFFFu - rev_[s]);
					if (key > best) best = key;
				}
				if (best) atomic_max(claim_ + comp_[v], best);
			});
			parallel_for<K_entry>(c, nA, PV_LAMBDA(size_t i) {
				u32 v = act[i];
				if (comp_[v] != v) return;
				u64 k = claim_[v];
				u32 s = 0xFFFFFFFFu - (u32)(k & 0xFFFFFFFFu);
				u32 p = adjOwner_[s], q0 = adjNbr_[s];
				rnk_[q0] = 0; entDepth_[v] = (u32)(k >> 32) + 1; entV_[v] = q0 >> 1;
				parent_[q0] = p; pslot_[q0] = s;
			});
		}
		// (f) "first unvisited edge" functional graph on entry ports
		parallel_for<K_first_edge>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
			u32 v = act[i >> 1]; u32 x = 2 * v + (u32)(i & 1); u32 xb = x ^ 1u;
			u32 f = NONE, fs = NONE;
			for (u32 s = adjOff_[xb], e = adjOff_[xb + 1]; s < e; s++) {
				u32 y = adjNbr_[s]; u32 w = y >> 1;
				if (w != v && !visited_[w]) { f = y; fs = s; break; }
			}
			F_[x] = f; Fslot_[x] = fs;
		});
		// (f') repair mode.  The DFS takes the first edge to a segment that is unvisited WHEN IT GETS THERE; F above only knows
		// what was unvisited when the round began, so an edge back to a segment earlier on the very path being stripped (the back
		// link of a repeat unit, the far link of an inversion) makes the stripped path run into itself and the round ends there.
		// A graph with thousands of such structures along one chain then needs thousands of rounds, each O(remaining graph).
		// After a round that was cut short this way, the next rounds PREDICT those edges: candidate w of port x is skipped when a
		// bounded walk of the first-edge chain from w reaches x (w comes before x on that chain) and w's index lies between the
		// path's entry vertex and x (chains of several ports can merge: a walk from the far side of an inversion also reaches x;
		// segment ids of real graphs follow the backbone, so the side the path came from is the one between entry and x).  The
		// prediction only decides how far a round gets: every skipped candidate of every stripped node is verified afterwards (K_skipcheck: it
		// really is earlier on the path), a failed check cuts the path right after that node, and an edge that should have been
		// skipped but was not still ends the path at the self-intersection, exactly as before.
		const u32 *Fuse_ = F_, *Fsuse_ = Fslot_;
		Buf<u32> F1, Fslot1;
		if (repair) {
			F1 = Buf<u32>(c, P); Fslot1 = Buf<u32>(c, P);
			u32 *F1_ = F1.p, *Fslot1_ = Fslot1.p;
			const u32 WALK = skip_walk;
			parallel_for<K_skip>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
				u32 v = act[i >> 1]; u32 x = 2 * v + (u32)(i & 1); u32 xb = x ^ 1u;
				u32 f = NONE, fs = NONE;
				const u32 q = entV_[comp_[v]]; const u32 lo = q < v ? q : v, hi = q < v ? v : q;
				for (u32 s = adjOff_[xb], e = adjOff_[xb + 1]; s < e; s++) {
					u32 y = adjNbr_[s]; u32 w = y >> 1;
					if (w == v || visited_[w]) continue;
					bool before = false;
					for (u32 side = 0; side < 2 && !before && w >= lo && w <= hi; side++) {
						u32 z = 2 * w + side;
						for (u32 k = 0; k < WALK; k++) { z = F_[z]; if (z == NONE) break; if (z == x) { before = true; break; } }
					}
					if (!before) { f = y; fs = s; break; }
				}
				F1_[x] = f; Fslot1_[x] = fs;
			});
			Fuse_ = F1_; Fsuse_ = Fslot1_;
		}
		// (g) pointer-doubling path marking from every entry
		{
			c.d2d(J.p, Fuse_, (size_t)P * 4);
			u32 *Ja = J.p, *Jb = J2.p;
			int maxIt = bits_for((u64)nA * 2) + 1;
			// itFlag[it] = "iteration it marked something".  Later iterations test the previous flag on the device and fall
			// through when the marking has converged; the host looks only every 8th iteration (one sync instead of eight).
			// (flags sit 128 bytes apart: a word that one iteration polls must not share a sector with the word the same
			// iteration's markers keep storing to.)  One launch per doubling step: all steps in one cooperative launch with
			// grid barriers (parallel_iter) was measured 3x SLOWER here -- the sweeps must then read the jump pointers and ranks
			// from L2 (another SM wrote them one sweep earlier), and this kernel lives on L1 hits (947 vs 312 us at cfg2).
			Buf<u32> itFlag(c, ((size_t)maxIt + 1) * 32);
			itFlag.zero();
			u32 *itFlag_ = itFlag.p;
			for (int it = 0; it < maxIt; it++) {
				const u32 lim = 1u << it;
				// mark the J^(2^it)-successors of the nodes ranked below 2^it, and square the jump pointers for the next step
				// (the squaring runs one step past convergence; its result is then unused)
				parallel_for<K_mark>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
					if (it > 0 && itFlag_[(it - 1) * 32] == 0) return;
					u32 v = act[i >> 1]; u32 x = 2 * v + (u32)(i & 1);
					u32 y = Ja[x];
					Jb[x] = y == NONE ? NONE : Ja[y];
					u32 r = rnk_[x];
					if (r >= lim || y == NONE) return;          // NONE or marked in this very iteration
					u32 rho = r + lim;
					u32 old = atomic_min(rnk_ + y, rho);
					if (old != NONE && old != rho) atomic_min(trunc_ + comp_[v], old > rho ? old : rho);
					itFlag_[it * 32] = 1;
				});
				// from round 1 on the host looks at the convergence flag after iteration 1 (later rounds strip short paths and stop
				// here) and every 8th iteration; round 0 follows a path as long as the graph and runs to the end unobserved
				if (round > 0 && (it == 1 || (it & 7) == 7) && !itFlag.read((size_t)it * 32)) break;
				std::swap(Ja, Jb);
			}
		}
		// (h) truncate at the first self-intersection, then fix parents / depths of the stripped paths
		parallel_for<K_segconf>(c, nA, PV_LAMBDA(size_t i) {
			u32 v = act[i]; u32 r0 = rnk_[2 * v], r1 = rnk_[2 * v + 1];
			if (r0 != NONE && r1 != NONE) atomic_min(trunc_ + comp_[v], r0 > r1 ? r0 : r1);
		});
		if (repair) {   // every candidate that a stripped node skipped must lie earlier on the stripped path
			parallel_for<K_skipcheck>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
				u32 v = act[i >> 1]; u32 x = 2 * v + (u32)(i & 1); u32 xb = x ^ 1u;
				u32 r = rnk_[x]; u32 root = comp_[v];
				if (r == NONE || r >= trunc_[root]) return;
				const u32 fs = Fsuse_[x];
				for (u32 s = adjOff_[xb], e = adjOff_[xb + 1]; s < e && s != fs; s++) {
					u32 w = adjNbr_[s] >> 1;
					if (w == v || visited_[w]) continue;
					u32 r0 = rnk_[2 * w], r1 = rnk_[2 * w + 1]; u32 m = r0 < r1 ? r0 : r1;
					if (!(m < r)) { atomic_min(trunc_ + root, r + 1); break; }   // wrong guess: x stays, its successor does not
				}
			});
		}
		parallel_for<K_finalize>(c, (size_t)nA * 2, PV_LAMBDA(size_t i) {
			u32 v = act[i >> 1]; u32 x = 2 * v + (u32)(i & 1);
			u32 r = rnk_[x]; u32 root = comp_[v];
			if (r == NONE || r >= trunc_[root]) return;
			u32 d = entDepth_[root] + 2 * r;
			entrySide_[v] = x & 1u;
			depth_[x] = d; depth_[x ^ 1u] = d + 1;
			parent_[x ^ 1u] = x; pslot_[x ^ 1u] = NONE;
			u32 y = Fuse_[x];
			if (y != NONE && rnk_[y] == r + 1 && r + 1 < trunc_[root]) { parent_[y] = x ^ 1u; pslot_[y] = Fsuse_[x]; }
		});
		Buf<u32> still(c, nA);
		u32 *still_ = still.p;
		c.memset0(flag.p + 2, 4);
		parallel_for<K_visit>(c, nA, PV_LAMBDA(size_t i) {
			u32 v = act[i]; u32 root = comp_[v]; u32 t = trunc_[root];
			u32 r0 = rnk_[2 * v], r1 = rnk_[2 * v + 1];
			bool vis = (r0 != NONE && r0 < t) || (r1 != NONE && r1 < t);
			still_[i] = vis ? 0u : 1u;
			if (root == v && t != NONE) flag_[2] = 1;   // a stripped path was cut short
		});
		// visited[] is written only after every kernel of this round that reads it has finished
		parallel_for<K_visit>(c, nA, PV_LAMBDA(size_t i) { if (!still_[i]) visited_[act[i]] = 1; });
		Buf<u32> idx, idxPos;
		u32 nNext = 0, cut = 0;
		compact_flags(c, still.p, nA, idx, &idxPos, &mrRound, &nNext); mrRound.add(&cut, flag.p + 2, 4); mrRound.wait();
#ifdef PV_EMU
		if (getenv("PV_DEBUG")) {
			fprintf(stderr, "[dfs] round %u: active %u -> %u, repair %d, cut %u;", round, nA, nNext, (int)repair, cut);
			for (u32 i = 0; i < nA && i < 48; i++) { u32 v = act[i]; if (comp_[v] == v && trunc_[v] != NONE) fprintf(stderr, " comp %u trunc %u", v, trunc_[v]); }
			fprintf(stderr, "\n");
			if (nA <= 40) for (u32 i = 0; i < nA; i++) { u32 v = act[i]; fprintf(stderr, "   v%u: rnkL %d rnkR %d F0L %d F0R %d FL %d FR %d comp %u\n", v, (int)rnk_[2 * v], (int)rnk_[2 * v + 1], (int)F_[2 * v], (int)F_[2 * v + 1], (int)Fuse_[2 * v], (int)Fuse_[2 * v + 1], comp_[v]); }
		}
#endif
		repair = cut != 0 && skip_walk != 0;
		Buf<u32> nextActive(c, nNext ? nNext : 1);
		{ u32 *na = nextActive.p; const u32 *ix = idx.p; parallel_for<K_active>(c, nNext, PV_LAMBDA(size_t i) { na[i] = act[ix[i]]; }); }
		active = std::move(nextActive);
		if (nNext >= nActive) throw std::runtime_error("decompose: DFS round made no progress");
		nActive = nNext;
		round++;
	}
	out.dfs_rounds = round;
	tr.mark("D2 dfs rounds");
	sm.mark();   // 2

	// ------------------------------------------------------------------ components (dense indices)
	u32 *compIdxOfRoot_ = compIdxOfRoot.p, *nvOfRoot_ = nvOfRoot.p;
	// per component: nv, node base (2*nv+1 nodes incl. virtual root), segment base
	Buf<u32> cNv(c, nComp + 1), cNodeBase(c, nComp + 1), cSegBase(c, nComp + 1), cNodes(c, nComp + 1), cHasTips(c, nComp);
	u32 *cNv_ = cNv.p, *cNodeBase_ = cNodeBase.p, *cSegBase_ = cSegBase.p, *cNodes_ = cNodes.p, *cHasTips_ = cHasTips.p;
	const u32 *roots_ = roots.p;
	cHasTips.zero();
	parallel_for<K_compcount>(c, (size_t)nComp + 1, PV_LAMBDA(size_t k) {
		u32 nv = k < nComp ? nvOfRoot_[roots_[k]] : 0u;
		cNv_[k] = nv; cNodes_[k] = k < nComp ? 2 * nv + 1 : 0u;
	});
	parallel_for<K_compcount>(c, V, PV_LAMBDA(size_t v) { if (tipPort_[v] != NONE) cHasTips_[compIdxOfRoot_[gcomp_[v]]] = 1; });
	if (nComp <= 2048) {   // a handful of components: one thread adds them up instead of two scan launches
		parallel_for<K_compcount>(c, 1, PV_LAMBDA(size_t) {
			u32 a = 0, b = 0;
			for (u32 k = 0; k <= nComp; k++) { u32 x = cNodes_[k], y = cNv_[k]; cNodeBase_[k] = a; cSegBase_[k] = b; a += x; b += y; }
		});
	} else {
		exclusive_scan(c, cNodes.p, cNodeBase.p, (size_t)nComp + 1, OpAddU32(), 0u);
		exclusive_scan(c, cNv.p, cSegBase.p, (size_t)nComp + 1, OpAddU32(), 0u);
	}
	const u32 N = P + nComp;   // nodes in global preorder space

	tr.mark("components");
	sm.mark();   // 3
	// ------------------------------------------------------------------ D3: Euler tour + list ranking
	Buf<u32> treeSlot(c, M ? M : 1);
	treeSlot.fill_ff();
	u32 *treeSlot_ = treeSlot.p;
	parallel_for<K_treeslot>(c, P, PV_LAMBDA(size_t y) { u32 s = pslot_[y]; if (s != NONE) treeSlot_[s] = (u32)y; });
	Buf<u32> enext(c, (size_t)2 * P), ew(c, (size_t)2 * P), esuf(c, (size_t)2 * P);
	u32 *enext_ = enext.p, *ew_ = ew.p;
	parallel_for<K_euler>(c, P, PV_LAMBDA(size_t yy) {
		u32 y = (u32)yy; u32 v = y >> 1;
		bool isEntry = (y & 1u) == entrySide_[v];
		auto firstGray = [&](u32 p, u32 from) -> u32 {
			for (u32 s = from, e = adjOff_[p + 1]; s < e; s++) { u32 ch = treeSlot_[s]; if (ch != NONE) return ch; }
			return NONE;
		};
		// down(y) = 2y, up(y) = 2y+1
		u32 fc = isEntry ? (y ^ 1u) : firstGray(y, adjOff_[y]);
		enext_[2 * y] = fc != NONE ? 2 * fc : 2 * y + 1;
		u32 par = parent_[y];
		u32 ns;
		if (!isEntry) ns = firstGray(par, adjOff_[par]);            // black child: siblings are the entry port's gray children
		else ns = par == NONE ? NONE : firstGray(par, pslot_[y] + 1);
		enext_[2 * y + 1] = ns != NONE ? 2 * ns : (par == NONE ? NONE : 2 * par + 1);
		ew_[2 * y] = 1; ew_[2 * y + 1] = 0;
	});
	list_rank_suffix(c, enext.p, ew.p, esuf.p, (size_t)2 * P);
	Buf<u32> pre(c, P), sz(c, P), pre2node(c, N);
	u32 *pre_ = pre.p, *sz_ = sz.p, *pre2node_ = pre2node.p;
	const u32 *esuf_ = esuf.p;
	parallel_for<K_pre>(c, P, PV_LAMBDA(size_t y) {
		u32 k = compIdxOfRoot_[gcomp_[y >> 1]];
		u32 T = 2 * cNv_[k];
		u32 lp = T - esuf_[2 * y];                   // local preorder, virtual root = 0
		u32 g = cNodeBase_[k] + lp;
		pre_[y] = g; sz_[y] = esuf_[2 * y] - esuf_[2 * y + 1] + 1;
		pre2node_[g] = (u32)y;
	});
	parallel_for<K_pre>(c, nComp, PV_LAMBDA(size_t k) { pre2node_[cNodeBase_[k]] = P + (u32)k; });

	tr.mark("D3 euler+rank");
	sm.mark();   // 4
	// ------------------------------------------------------------------ D4: back edges, bridges, labels, keys
	Buf<u64> X(c, (size_t)2 * (N + 1));      // 128-bit label accumulators, (lo, hi) interleaved
	Buf<i32> cd(c, N + 1);
	Buf<u32> minT(c, N);
	X.zero(); cd.zero(); minT.fill_ff();
	u64 *X_ = X.p; i32 *cd_ = cd.p; u32 *minT_ = minT.p;
	const u64 seed = opt.seed;
	parallel_for<K_backedge>(c, P, PV_LAMBDA(size_t pp) {
		u32 p = (u32)pp; u32 v = p >> 1;
		u32 k = compIdxOfRoot_[gcomp_[v]];
		u32 myPre = pre_[p];
		for_each_backedge(p, adjOff_, adjNbr_, pre_, parent_, entrySide_, cNodeBase_[k], M, [&](u32 tgtPre, u64 id, u32) {
			u64 lo = splitmix64(seed ^ (id * 2)), hi = splitmix64((seed * 0x9E3779B97F4A7C15ull) ^ (id * 2 + 1));
			atomic_xor(X_ + 2 * (size_t)myPre, lo); atomic_xor(X_ + 2 * (size_t)myPre + 1, hi);
			atomic_xor(X_ + 2 * (size_t)tgtPre, lo); atomic_xor(X_ + 2 * (size_t)tgtPre + 1, hi);
			atomic_add((u32 *)cd_ + myPre, 1u); atomic_add((u32 *)cd_ + tgtPre, 0xFFFFFFFFu);
			atomic_min(minT_ + myPre, tgtPre);
		});
	});
	// bridges -> simplifying back edges (flubbles.cpp:621-643): v gets one iff no bracket leaves subtree(v)
	// upwards and no vertex below v got one
	Rmq rmq1;
	rmq_build(c, rmq1, minT.p, N);
	Buf<u32> bridge(c, N + 1), bsum(c, N + 1);
	u32 *bridge_ = bridge.p, *bsum_ = bsum.p;
	{
		RmqView rv = rmq1.view();
		parallel_for<K_bridge>(c, (size_t)N + 1, PV_LAMBDA(size_t g) {
			u32 b = 0;
			if (g < N) {
				u32 node = pre2node_[g];
				if (node < P) {
					u32 k = compIdxOfRoot_[gcomp_[node >> 1]];
					bool refRoot = !cHasTips_[k] && g == cNodeBase_[k] + 1;   // without tips the start port is the tree root
					if (!refRoot) { u32 hi = rv.query((u32)g, (u32)g + sz_[node] - 1); b = hi >= (u32)g ? 1u : 0u; }
				}
			}
			bridge_[g] = b;
		});
	}
	exclusive_scan(c, bridge.p, bsum.p, (size_t)N + 1, OpAddU32(), 0u);
	parallel_for<K_simp>(c, N, PV_LAMBDA(size_t g) {
		if (!bridge_[g]) return;
		u32 node = pre2node_[g];
		u32 inSub = bsum_[g + sz_[node]] - bsum_[g];
		if (inSub != 1) return;
		u32 k = compIdxOfRoot_[gcomp_[node >> 1]];
		u32 tgt = cHasTips_[k] ? cNodeBase_[k] : cNodeBase_[k] + 1;
		u64 id = (u64)M + P + node;
		u64 lo = splitmix64(seed ^ (id * 2)), hi = splitmix64((seed * 0x9E3779B97F4A7C15ull) ^ (id * 2 + 1));
		atomic_xor(X_ + 2 * g, lo); atomic_xor(X_ + 2 * g + 1, hi);
		atomic_xor(X_ + 2 * (size_t)tgt, lo); atomic_xor(X_ + 2 * (size_t)tgt + 1, hi);
		atomic_add((u32 *)cd_ + g, 1u); atomic_add((u32 *)cd_ + tgt, 0xFFFFFFFFu);
		atomic_min(minT_ + g, tgt);
	});
	Rmq rmq2;
	rmq_build(c, rmq2, minT.p, N);
	Buf<u128> PX(c, N + 1);
	Buf<i32> PC(c, N + 1);
	exclusive_scan(c, (const u128 *)X.p, PX.p, (size_t)N + 1, OpXor128(), u128{0, 0});
	exclusive_scan(c, cd.p, PC.p, (size_t)N + 1, OpAddI32(), 0);
	const u128 *PX_ = PX.p; const i32 *PC_ = PC.p;

	tr.mark("D4 labels");
	sm.mark();   // 5
	// ------------------------------------------------------------------ D5: class stack positions
	// pos(seg) = #entry ports before + root-path sum of w + G(entry)   (see DESIGN.md "class stack without a second traversal")
	Buf<u32> isEnt(c, N + 1), S(c, N + 1);
	Buf<i32> Dw(c, N + 1), PW(c, N + 1);
	Dw.zero();
	u32 *isEnt_ = isEnt.p, *S_ = S.p; i32 *Dw_ = Dw.p, *PW_ = PW.p;
	parallel_for<K_stackw>(c, (size_t)N + 1, PV_LAMBDA(size_t g) {
		u32 e = 0;
		if (g < N) { u32 node = pre2node_[g]; if (node < P && (node & 1u) == entrySide_[node >> 1]) e = 1; }
		isEnt_[g] = e;
	});
	exclusive_scan(c, isEnt.p, S.p, (size_t)N + 1, OpAddU32(), 0u);
	parallel_for<K_stackw>(c, P, PV_LAMBDA(size_t yy) {
		u32 y = (u32)yy; u32 par = parent_[y];
		if (par == NONE) return;
		u32 pv_ = par >> 1;
		bool parIsEntry = (par & 1u) == entrySide_[pv_];
		if (!parIsEntry) return;
		u32 a = par, ab = par ^ 1u;
		u32 ga = pre_[a], gab = pre_[ab];
		i32 nsegA = (i32)(S_[ga + sz_[a]] - S_[ga]);
		i32 nsegAb = (i32)(S_[gab + sz_[ab]] - S_[gab]);
		i32 w = (y == ab) ? (nsegA - 1 - nsegAb) : -(1 + nsegAb);
		u32 gy = pre_[y];
		atomic_add((u32 *)Dw_ + gy, (u32)w); atomic_add((u32 *)Dw_ + gy + sz_[y], (u32)(-w));
	});
	inclusive_scan(c, Dw.p, PW.p, (size_t)N + 1, OpAddI32(), 0);
	const u32 nS = V;   // one stack entry per segment
	Buf<u32> stV(c, nS), stPre(c, nS), stCnt(c, nS);
	Buf<u128> stKey(c, nS);
	u32 *stV_ = stV.p, *stPre_ = stPre.p, *stCnt_ = stCnt.p; u128 *stKey_ = stKey.p;
	parallel_for<K_stackpos>(c, V, PV_LAMBDA(size_t v) {
		u32 u = (u32)(2 * v) + entrySide_[v], cx = u ^ 1u;
		u32 k = compIdxOfRoot_[gcomp_[v]];
		u32 gu = pre_[u], gc = pre_[cx];
		i32 segsBefore = (i32)(S_[gu] - S_[cNodeBase_[k]]);
		i32 nsegU = (i32)(S_[gu + sz_[u]] - S_[gu]);
		i32 nsegC = (i32)(S_[gc + sz_[cx]] - S_[gc]);
		i32 G = nsegU - 1 - nsegC;
		i32 pos = segsBefore + PW_[gu] + G;
		u32 gp = cSegBase_[k] + (u32)pos;
		stV_[gp] = (u32)v; stPre_[gp] = gc;
		u128 a = PX_[gc + sz_[cx]], b = PX_[gc];
		stKey_[gp] = u128{a.lo ^ b.lo, a.hi ^ b.hi};
		stCnt_[gp] = (u32)(PC_[gc + sz_[cx]] - PC_[gc]);
	});
	sm.mark();   // 6
	// group equal keys: ONE stable radix sort by a 32-bit word of the 128-bit key (4 onesweep passes instead of 8-16).  Keys
	// that agree on that word but differ elsewhere (a few dozen pairs per million classes, by the birthday bound) interleave
	// inside a run of equal words; the neighbour search below steps over them, which costs the length of the foreign stretch.
	// The sort is stable in stack position, so "next / previous in the run with my key" is next / previous seen in the stack.
	Buf<u32> ord(c, nS), wkey(c, nS);
	u32 *ord_ = ord.p, *wkey_ = wkey.p;
	parallel_for<K_word>(c, nS, PV_LAMBDA(size_t i) { ord_[i] = (u32)i; wkey_[i] = (u32)stKey_[i].lo; });
	radix_sort_pairs(c, wkey, ord, nS, 32);
	wkey_ = wkey.p; ord_ = ord.p;
	Buf<u32> nxt(c, nS), prv(c, nS);
	u32 *nxt_ = nxt.p, *prv_ = prv.p;
	parallel_for<K_seen>(c, nS, PV_LAMBDA(size_t i) {
		const u32 me = ord_[i], w = wkey_[i]; const u128 km = stKey_[me];
		u32 nx = me, pv_ = NONE;
		for (size_t j = i + 1; j < nS && wkey_[j] == w; j++) { u32 o = ord_[j]; u128 k = stKey_[o]; if (k.lo == km.lo && k.hi == km.hi) { nx = o; break; } }
		for (size_t j = i; j-- > 0 && wkey_[j] == w;) { u32 o = ord_[j]; u128 k = stKey_[o]; if (k.lo == km.lo && k.hi == km.hi) { pv_ = o; break; } }
		nxt_[me] = nx; prv_[me] = pv_;
	});

	// exact verification of every merged pair (bracket sets equal <=> same count, ancestor chain, and no
	// bracket from subtree(v1) \ subtree(v2) leaves subtree(v1))
	c.memset0(flag.p, 16);
	// the reference names classes by (top bracket, list size); with >= 3 children reaching above a vertex its capping edge can be
	// too short and that naming merges classes that cycle equivalence keeps apart (pv_refnum.cuh: refnum_quirk_flag).  The count
	// travels with the first host wait of D6; only then are the classes re-derived the reference's way.
	RefNumIn rn{V, P, N, nComp, nS, M, adjOff_, adjNbr_, treeSlot_, parent_, entrySide_, pre_, sz_, depth_, pre2node_, compIdxOfRoot_, gcomp_, cNodeBase_, cNv_, cHasTips_, cSegBase_,
		    bridge_, bsum_, PC_, rmq2.view(), minT_, stV_, nullptr};
	static const bool force_faithful = getenv("POVU_B200_FAITHFUL_CLASSES") != nullptr, never_faithful = getenv("POVU_B200_XOR_CLASSES") != nullptr;
	if (!never_faithful) refnum_quirk_flag(c, rn, flag.p + 3);
	{
		RmqView rv = rmq2.view();
		parallel_for<K_verify>(c, nS, PV_LAMBDA(size_t i) {
			u32 j = nxt_[i];
			if (j == (u32)i) return;
			u32 c1 = stPre_[i], c2 = stPre_[j];
			u32 z1 = sz_[pre2node_[c1]], z2 = sz_[pre2node_[c2]];
			bool ok = c1 < c2 && c2 + z2 <= c1 + z1 && stCnt_[i] == stCnt_[j];
			if (ok && c2 > c1) ok = rv.query(c1, c2 - 1) >= c1;
			if (ok && c2 + z2 < c1 + z1) ok = rv.query(c2 + z2, c1 + z1 - 1) >= c1;
			if (!ok) atomic_add(flag_ + 1, 1u);
		});
	}

	tr.mark("D5 stack+sort");
	sm.mark();   // 7
	// ------------------------------------------------------------------ D6: flubbles + PVST parents
	// A class met again at stack position i closes what was opened since its previous occurrence p = prv[i] -- IF p is still on
	// the reference's stack: an earlier re-encounter (p', j) with p' < p < j < i has popped p with everything above p'
	// (add_flubbles, flubbles.cpp:295-367).  "Arc (p, i) fires" is therefore: no FIRED arc crosses it from the left.  In a
	// laminar component no arcs cross at all and every arc fires (one range-minimum pass shows it); otherwise the fired set is the
	// fixpoint of: an arc with no possible crosser fires, an arc crossed by a fired arc is dead -- the undecided arc with the
	// smallest right end is always decidable, and real graphs (LPA's repeat loop) settle in a few rounds.  Every round is one
	// range-minimum structure per status plus one kernel; the stack machine itself (one thread walking a whole chromosome) is
	// only the fallback for inputs that do not settle within 64 rounds.
	Buf<u32> segComp(c, nS), fire(c, nS), lamBad(c, nComp ? nComp : 1);
	u32 *segComp_ = segComp.p, *fire_ = fire.p, *lamBad_ = lamBad.p;
	lamBad.zero();
	parallel_for<K_arc>(c, nS, PV_LAMBDA(size_t i) { segComp_[i] = compIdxOfRoot_[gcomp_[stV_[i]]]; });
	Buf<u32> open(c, nS + 1), flIdx(c, nS + 1);
	Buf<MP> fn(c, nS), fnS(c, nS);
	u32 *open_ = open.p, *flIdx_ = flIdx.p; MP *fn_ = fn.p, *fnS_ = fnS.p;
	parallel_for<K_fn>(c, (size_t)nS + 1, PV_LAMBDA(size_t i) { open_[i] = i < nS && nxt_[i] > (u32)i + 1 ? 1u : 0u; });
	exclusive_scan(c, open.p, flIdx.p, (size_t)nS + 1, OpAddU32(), 0u);
	u32 nFl = 0, quirk = 0;
	int arcRounds = fired_arcs(c, prv.p, nS, fire.p, segComp.p, lamBad.p, flIdx.p + nS, &nFl, flag.p + 3, &quirk);
	out.faithful_classes = 0;
	if ((quirk || force_faithful) && !never_faithful) {
		// classes the reference's way (capping edges, top bracket + list size: pv_refnum.cuh), then next / previous seen by one stable
		// sort of the stack positions on the class label; everything after this point is unchanged
		Buf<u32> sc;
		refnum_device(c, rn, nullptr, &sc);
		Buf<u32> ckey(c, nS), cord(c, nS);
		{ u32 *k_ = ckey.p, *o_ = cord.p; const u32 *sc_ = sc.p; parallel_for<K_word>(c, nS, PV_LAMBDA(size_t i) { k_[i] = sc_[i]; o_[i] = (u32)i; }); }
		radix_sort_pairs(c, ckey, cord, nS, bits_for((u64)N + 1));
		{
			const u32 *k_ = ckey.p, *o_ = cord.p;
			parallel_for<K_seen>(c, nS, PV_LAMBDA(size_t i) {
				const u32 me = o_[i], k = k_[i];
				nxt_[me] = (i + 1 < nS && k_[i + 1] == k) ? o_[i + 1] : me;
				prv_[me] = (i > 0 && k_[i - 1] == k) ? o_[i - 1] : NONE;
			});
		}
		lamBad.zero();
		parallel_for<K_fn>(c, (size_t)nS + 1, PV_LAMBDA(size_t i) { open_[i] = i < nS && nxt_[i] > (u32)i + 1 ? 1u : 0u; });
		exclusive_scan(c, open.p, flIdx.p, (size_t)nS + 1, OpAddU32(), 0u);
		arcRounds = fired_arcs(c, prv.p, nS, fire.p, segComp.p, lamBad.p, flIdx.p + nS, &nFl);
		out.faithful_classes = quirk ? quirk : 1;
	}
	out.arc_rounds = arcRounds < 0 ? 65u : (u32)arcRounds;
	const bool replay = arcRounds < 0;      // the sequential fallback
	parallel_for<K_fn>(c, nS, PV_LAMBDA(size_t i) {
		u32 o = open_[i];
		u32 u = fire_[i] == 1 ? 1u : 0u;
		bool first = i == 0 || segComp_[i - 1] != segComp_[i];
		fn_[i] = MP{first ? -(1 << 29) : (i32)o - (i32)u, (i32)o};
	});
	inclusive_scan(c, fn.p, fnS.p, nS, OpMP(), MP{0, -(1 << 29)});
	Buf<u8> fAf(c, nFl ? nFl : 1), fZf(c, nFl ? nFl : 1);
	u8 *fAf_ = fAf.p, *fZf_ = fZf.p;
	Buf<u32> fBase(c, nFl ? nFl : 1);
	u32 *fBase_ = fBase.p;
	Buf<u32> fA(c, nFl ? nFl : 1), fZ(c, nFl ? nFl : 1), fOr(c, nFl ? nFl : 1), fDepth(c, nFl ? nFl : 1), fPar(c, nFl ? nFl : 1), fComp(c, nFl ? nFl : 1);
	Buf<u32> cFlBase(c, nComp + 1);
	u32 *fA_ = fA.p, *fZ_ = fZ.p, *fOr_ = fOr.p, *fDepth_ = fDepth.p, *fPar_ = fPar.p, *fComp_ = fComp.p, *cFlBase_ = cFlBase.p;
	parallel_for<K_flub>(c, (size_t)nComp + 1, PV_LAMBDA(size_t k) { cFlBase_[k] = flIdx_[cSegBase_[k]]; });
	parallel_for<K_flub>(c, nS, PV_LAMBDA(size_t i) {
		if (!open_[i]) return;
		u32 f = flIdx_[i]; u32 j = nxt_[i];
		u32 va = stV_[i], vz = stV_[j];
		u32 oa = entrySide_[va] == 0 ? 1u : 0u, oz = entrySide_[vz] == 0 ? 1u : 0u;   // forward iff the child side is the right one
		if (!oa && !oz) { u32 t = va; va = vz; vz = t; oa = 1; oz = 1; }             // normalize_endpoints (flubbles.cpp:233-243)
		fA_[f] = va; fZ_[f] = vz; fOr_[f] = oa | (oz << 1); fAf_[f] = (u8)oa; fZf_[f] = (u8)oz;
		fBase_[f] = cFlBase_[segComp_[i]];
		MP m = fnS_[i];
		fDepth_[f] = (u32)(m.a > m.b ? m.a : m.b);
		fComp_[f] = segComp_[i];
	});
	Rmq rmq4;
	if (nFl) {
		rmq_build(c, rmq4, fDepth.p, nFl);
		RmqView rv = rmq4.view();
		parallel_for<K_parent>(c, nFl, PV_LAMBDA(size_t f) {
			u32 d = fDepth_[f]; u32 k = fComp_[f];
			if (replay && lamBad_[k]) return;
			u32 par = 0;
			if (d > 1 && f > 0) {
				u32 j = rv.prev_smaller((u32)f - 1, d);
				if (j != NONE && j >= cFlBase_[k]) par = j - cFlBase_[k] + 1;
			}
			fPar_[f] = par;
		});
	}
	// non-laminar components: replay the reference's stack machine, one thread per component.  The stack holds at most one
	// entry per class (a re-encountered class pops down to its previous entry first), so "is my class on the stack" is "is my
	// previous occurrence still on the stack": positions stand in for class ids.
	Buf<u32> inS(c, replay ? nS : 1), stk(c, replay ? nS : 1);
	if (replay) inS.zero();
	u32 *inS_ = inS.p, *stk_ = stk.p;
	if (replay) parallel_for<K_seq>(c, nComp, PV_LAMBDA(size_t k) {
		if (!lamBad_[k]) return;
		u32 b = cSegBase_[k], e = cSegBase_[k + 1], fb = cFlBase_[k];
		u32 top = b;           // stack occupies stk_[b .. top)
		u32 prt = 0;           // local PVST idx
		for (u32 i = b; i < e; i++) {
			u32 p = prv_[i];
			if (p != NONE && inS_[p]) {
				while (top > b) { u32 x = stk_[--top]; inS_[x] = 0; if (x == p) break; }
				if (prt != 0) prt = fPar_[fb + prt - 1];
			}
			if (open_[i]) {
				u32 f = flIdx_[i];
				fPar_[f] = prt;
				fDepth_[f] = prt == 0 ? 1u : fDepth_[fb + prt - 1] + 1;
				prt = f - fb + 1;
			}
			stk_[top++] = i; inS_[i] = 1;
		}
	});

	tr.mark("D6 flubbles");
	sm.mark();   // 8
	// ------------------------------------------------------------------ results to the host
	const bool bg = opt.async_results && opt.keep_forest;   // (the sources must outlive the copies: only when the caller keeps them)
	if (bg) {
		c.bg_begin();
		fA.to_hvec_bg(out.a_v, nFl); fZ.to_hvec_bg(out.z_v, nFl); fPar.to_hvec_bg(out.parent, nFl); fDepth.to_hvec_bg(out.depth, nFl); fComp.to_hvec_bg(out.comp, nFl);
		fAf.to_hvec_bg(out.a_fwd, nFl); fZf.to_hvec_bg(out.z_fwd, nFl); fBase.to_hvec_bg(out.fl_base, nFl);
	} else {
		fA.to_hvec(out.a_v, nFl); fZ.to_hvec(out.z_v, nFl); fPar.to_hvec(out.parent, nFl); fDepth.to_hvec(out.depth, nFl); fComp.to_hvec(out.comp, nFl);
		fAf.to_hvec(out.a_fwd, nFl); fZf.to_hvec(out.z_fwd, nFl); fBase.to_hvec(out.fl_base, nFl);
	}
	// the small per-component tables and the flags travel as ONE packed copy (one host wait instead of six)
	{
		const size_t PK = (size_t)nComp + 1;
		Buf<u32> pack(c, 4 + 5 * PK);
		u32 *pk = pack.p; const u32 *rt_ = roots.p, *nv_ = cNv.p, *tp_ = cHasTips.p, *lb_ = lamBad.p, *fb_ = cFlBase.p;
		parallel_for<K_pack>(c, PK > 4 ? PK : 4, PV_LAMBDA(size_t k) {
			if (k < 4) pk[k] = flag_[k];
			if (k >= PK) return;
			const bool in = k < nComp;
			pk[4 + k] = in ? rt_[k] : 0u; pk[4 + PK + k] = nv_[k]; pk[4 + 2 * PK + k] = in ? tp_[k] : 0u; pk[4 + 3 * PK + k] = in ? lb_[k] : 0u; pk[4 + 4 * PK + k] = fb_[k];
		});
		std::vector<u32> h = pack.to_host();   // synchronises the main stream
		out.collisions = h[1];
		out.comp_root.assign(h.begin() + 4, h.begin() + 4 + nComp);
		out.comp_nv.assign(h.begin() + 4 + PK, h.begin() + 4 + PK + nComp);
		out.comp_has_tips.assign(h.begin() + 4 + 2 * PK, h.begin() + 4 + 2 * PK + nComp);
		out.comp_nonlaminar.assign(h.begin() + 4 + 3 * PK, h.begin() + 4 + 3 * PK + nComp);
		out.comp_nfl.resize(nComp);
		for (u32 k = 0; k < nComp; k++) out.comp_nfl[k] = h[4 + 4 * PK + k + 1] - h[4 + 4 * PK + k];
	}
	if (opt.want_stack) {
		out.stack_v = stV.to_host();
		{   // dense class ids in order of first appearance on the stack (debug output only: derived from the previous-seen links)
			std::vector<u32> hp = prv.to_host(); out.stack_class.resize(nS); u32 next_id = 0;
			for (u32 i = 0; i < nS; i++) out.stack_class[i] = hp[i] == NONE ? next_id++ : out.stack_class[hp[i]];
		}
		std::vector<u32> es = entrySide.to_host();
		out.stack_fwd.resize(nS);
		for (u32 i = 0; i < nS; i++) out.stack_fwd[i] = es[out.stack_v[i]] == 0;
		out.stack_next = nxt.to_host();
	}
	if (opt.refnum) refnum_device(c, rn, opt.refnum);
	tr.mark("results d2h");   // (pack.to_host() above waited for the stream: every copy queued on it before has landed)
	if (sm.on && sm.n == 9) {
		StageCost &sc = StageCost::inst();
		u64 tips = 0; for (u32 k = 0; k < nComp; k++) tips += out.comp_has_tips[k] ? 1 : 0;   // (components with tips; the exact tip count stays on the device)
		const u64 treeE = (u64)N - nComp, nodes = N;
		sc.add(0, 1, (u64)V + E, nComp, sm.ns_between(0, 1) + sm.ns_between(2, 3));
		sc.add(1, nComp, (u64)V + E + tips, nodes + treeE, sm.ns_between(1, 2) + sm.ns_between(3, 4));
		sc.add(3, nComp, nodes + treeE, treeE, sm.ns_between(4, 5));
		sc.add(2, nComp, nodes + treeE, nS, sm.ns_between(5, 6));
		sc.add(4, nComp, nS, nFl, sm.ns_between(6, 7));
		sc.add(5, nComp, nS, (u64)nFl + nComp, sm.ns_between(7, 8));
	}
	if (opt.keep_labels) *opt.keep_labels = std::move(gcomp);
	if (opt.keep_forest) {
		DevForest &d = *opt.keep_forest;
		d.a_v = std::move(fA); d.z_v = std::move(fZ); d.parent = std::move(fPar); d.depth = std::move(fDepth); d.comp = std::move(fComp); d.fl_base = std::move(fBase);
		d.a_fwd = std::move(fAf); d.z_fwd = std::move(fZf); d.n = nFl; d.valid = true;
	}
}

} // namespace pv
