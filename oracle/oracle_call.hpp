/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into or called by the product library.
 * CPU restatement of the reference's `call` half for NON-TANGLED regions (the path every Lean fixture
 * except the hairpin, and every SNP/indel synthetic config, takes).  Tangled regions (a boundary column
 * of the depth matrix visited more than once -> untangle.cpp + align.cpp) and the SNE/SUBR inversion
 * pass (sne.cpp, slice_tree.hpp) are NOT restated yet: regions that would need them are counted in
 * CallStats and skipped, and the tests compare against the compiled reference (oracle/_ref) on the
 * records with TANGLED=F and VARTYPE != SUBR only.  "parity unpinned" for those two sub-paths.
 *
 * Follows:
 *   refs / genotype columns ... /root/reference/include/povu/refs/refs.hpp:278-301,359-412
 *   color_pvst ................ /root/reference/src/ita/variation/color.cpp:15-30,104-163
 *   find_pvst_rovs / gen_rov .. /root/reference/src/ita/variation/rov.cpp:185-268
 *   comp_bfs_tree ............. /root/reference/src/ita/graph/graph.cpp:75-163
 *   BfsTree::sort (Kahn) ...... /root/reference/include/ita/graph/bfs_tree.hpp:257-328
 *   find_laps / gen_sort ...... /root/reference/src/ita/graph/graph.cpp:165-395
 *   depth_matrix::fill ........ /root/reference/include/ita/genomics/allele.hpp:804-862
 *   gen_race (lineup+cluster) . /root/reference/src/ita/genomics/untangle.cpp:29-103
 *   comp_exps and helpers ..... /root/reference/src/ita/variation/overlay.cpp:84-588
 *   minimal_rov / alt_set ..... /root/reference/include/ita/genomics/allele.hpp:302-475
 *   gen_vcf_records ........... /root/reference/src/ita/genomics/vcf.cpp:20-312
 *   VcfRec fields ............. /root/reference/include/ita/genomics/vcf.hpp:33-393
 *   VCF text .................. /root/reference/src/mto/to_vcf.cpp:19-139
 *   chunking .................. /root/reference/src/ita/genomics/genomics.cpp:69-164
 */
#pragma once
#include <cmath>
#include <ctime>
#include <queue>
#include <sstream>
#include "oracle_decompose.hpp"

namespace oracle {

struct IdOr { u32 v_id; bool fwd; };
inline bool operator<(const IdOr &a, const IdOr &b) { return a.v_id < b.v_id || (a.v_id == b.v_id && (int)!a.fwd < (int)!b.fwd); }   /* types.cpp:112-120: forward < reverse */
inline bool operator==(const IdOr &a, const IdOr &b) { return a.v_id == b.v_id && a.fwd == b.fwd; }
inline bool operator!=(const IdOr &a, const IdOr &b) { return !(a == b); }
inline std::string str(const IdOr &x) { return std::string(x.fwd ? ">" : "<") + std::to_string(x.v_id); }
typedef std::vector<IdOr> Walk;

struct Pvst {            /* one .pvst: vertex 0 = dummy root */
	u32 component_id;
	std::vector<Flubble> fl;           /* fl[k] = PVST vertex k+1 */
	std::vector<u32> height;           /* per PVST vertex, root 0 (pvst.hpp:807-836) */
	std::vector<std::vector<u32>> children;
	void finish() {
		size_t n = fl.size() + 1;
		children.assign(n, {}); height.assign(n, 0);
		for (u32 k = 0; k < fl.size(); k++) children[fl[k].parent].push_back(k + 1);
		std::vector<u32> st{0};
		while (!st.empty()) { u32 v = st.back(); st.pop_back(); for (u32 c : children[v]) { height[c] = height[v] + 1; st.push_back(c); } }
	}
};

/* ---------- refs (refs.hpp) ---------- */
struct GtMeta {
	std::vector<std::string> col_names; std::vector<u32> ploidy;
	std::map<u32, std::pair<u32, u32>> ref2col;   /* ref id -> (column, phase) */
};
inline bool is_prefix(const std::string &p, const std::string &s) { return s.compare(0, p.size(), p) == 0 && p.size() <= s.size(); }
inline std::set<u32> refs_in_sample(const Graph &g, const std::string &sample) {   /* refs.hpp:278-301 */
	std::set<u32> out;
	for (u32 r = 0; r < g.paths.size(); r++) {
		const Path &p = g.paths[r];
		if (p.pansn) { if (p.sample == sample) out.insert(r); }
		else if (is_prefix(sample, p.sample)) out.insert(r);
		if (is_prefix(sample, p.tag)) out.insert(r);
	}
	return out;
}
inline GtMeta genotype_metadata(const Graph &g) {   /* refs.hpp:359-412 */
	GtMeta m; std::set<u32> handled;
	for (u32 r = 0; r < g.paths.size(); r++) {
		if (handled.count(r)) continue;
		handled.insert(r);
		std::set<u32> same = refs_in_sample(g, g.paths[r].sample);
		if (same.size() <= 1) { m.ref2col[r] = {(u32)m.col_names.size(), 0}; m.ploidy.push_back(1); m.col_names.push_back(g.paths[r].sample); }
		else {
			u32 hc = (u32)m.col_names.size(); std::set<u32> haps;
			for (u32 x : same) { m.ref2col[x] = {hc, g.paths[x].hap_id - 1}; haps.insert(g.paths[x].hap_id); handled.insert(x); }
			m.ploidy.push_back((u32)haps.size()); m.col_names.push_back(g.paths[r].sample);
		}
	}
	return m;
}

/* ---------- region selection (color.cpp:104-163) ---------- */
inline bool has_all_refs(const Graph &g, const std::set<u32> &refs, u32 s_idx, u32 t_idx) {
	for (u32 r : refs) if (g.v2steps[s_idx][r].empty() || g.v2steps[t_idx][r].empty()) return false;
	return true;
}
inline std::set<u32> color_pvst(const Graph &g, const Pvst &p, const std::set<u32> &refs) {
	std::set<u32> to_call; std::vector<u32> st{0};
	while (!st.empty()) {
		u32 cur = st.back(); st.pop_back();
		if (cur != 0) {
			const Flubble &f = p.fl[cur - 1];
			if (has_all_refs(g, refs, g.id2idx.at(f.a_id), g.id2idx.at(f.z_id))) to_call.insert(cur);
			u32 par = f.parent;
			if (to_call.count(par) && to_call.count(cur)) to_call.erase(par);
		}
		for (u32 c : p.children[cur]) st.push_back(c);
	}
	return to_call;
}

/* ---------- vertex order inside a region (graph.cpp:75-163, bfs_tree.hpp:257-328) ---------- */
struct BfsEdge { u32 from, to; bool tree; };
inline std::vector<u32> bfs_sort(const Graph &g, const Flubble &f) {
	typedef std::pair<u32, bool> IdxOr;   /* (v_idx, forward) */
	IdxOr start{g.id2idx.at(f.a_id), f.a_fwd}, end{g.id2idx.at(f.z_id), f.z_fwd};
	std::queue<IdxOr> q; q.push(start);
	std::set<u32> seen{start.first};
	std::vector<IdOr> tv; std::vector<BfsEdge> edges;
	std::map<std::pair<u32, int>, u32> parent_map;   /* key (v_idx, orientation: forward=0 < reverse=1) */
	auto key = [](const IdxOr &x) { return std::make_pair(x.first, x.second ? 0 : 1); };
	tv.push_back({g.ids[start.first], start.second}); parent_map[key(start)] = 0;
	const size_t MAX_UNBLOCK_CTR = 10000;
	while (!q.empty()) {
		IdxOr cur = q.front(); q.pop();
		if (q.size() > MAX_UNBLOCK_CTR) continue;
		u32 v = cur.first; u32 ve = cur.second ? SIDE_R : SIDE_L;     /* OUT end */
		const std::set<u32> &nb = ve == SIDE_L ? g.e_l[v] : g.e_r[v];
		for (u32 e : nb) {
			auto [side, alt] = g.other(e, v, ve);
			IdxOr nbr{alt, side == SIDE_L};                        /* get_or(side, IN) */
			if (seen.count(nbr.first)) { u32 from = parent_map[key(cur)]; u32 to = parent_map[key(nbr)]; edges.push_back({from, to, false}); continue; }
			if (cur == end) continue;
			u32 c = (u32)tv.size(); tv.push_back({g.ids[nbr.first], nbr.second});
			parent_map[key(nbr)] = c; edges.push_back({parent_map[key(cur)], c, true});
			seen.insert(nbr.first); q.push(nbr);
		}
	}
	/* Kahn */
	u32 N = (u32)tv.size();
	std::vector<u32> deg(N, 0), order;
	for (auto &e : edges) { if (e.tree) deg[e.to]++; else if (e.from != e.to) deg[e.to]++; }
	std::queue<u32> kq; kq.push(0); std::set<u32> done{0};
	while (!kq.empty()) {
		u32 cur = kq.front(); kq.pop(); order.push_back(cur);
		for (auto &e : edges) {
			if (!e.tree && e.to == e.from) continue;
			if (e.from != cur) continue;
			u32 u = e.to;
			if (done.count(u)) continue;
			deg[u]--;
			if (deg[u] > 0) continue;
			kq.push(u); done.insert(u);
		}
	}
	if (done.size() != N) return {};
	std::vector<u32> out; for (u32 t : order) out.push_back(tv[t].v_id);
	return out;
}

inline void find_laps(const Graph &g, u32 h, u32 u_id, u32 v_id, std::vector<std::pair<u32, u32>> &laps) {   /* graph.cpp:165-192 */
	std::vector<u32> a = g.v2steps[g.id2idx.at(u_id)][h], b = g.v2steps[g.id2idx.at(v_id)][h];
	if (a.empty() || b.empty()) return;
	size_t N = std::min(a.size(), b.size());
	for (size_t i = 0; i < N; i++) { u32 s = a[i], e = b[i]; if (s > e) std::swap(s, e); laps.push_back({s, e - s + 1}); }
}
/* fallback order when the BFS/Kahn sort fails (graph.cpp:194-395); throws where the reference would */
inline std::vector<u32> gen_sort(const Graph &g, const Flubble &f) {
	std::list<u32> sw; std::map<u32, std::list<u32>::iterator> w2it;
	for (u32 h = 0; h < g.paths.size(); h++) {
		std::vector<std::pair<u32, u32>> laps; find_laps(g, h, f.a_id, f.z_id, laps);
		const Path &rw = g.paths[h];
		for (auto [start, len] : laps) {
			if (sw.empty()) { for (u32 j = start; j < start + len; j++) { u32 v = rw.v_ids[j]; if (w2it.count(v)) continue; w2it[v] = sw.insert(sw.end(), v); } continue; }
			std::map<u32, std::list<u32>::iterator> lc, rc;
			bool hx = false, hy = false; std::list<u32>::iterator x, y;
			for (u32 j = start; j < start + len; j++) {
				u32 v = rw.v_ids[j];
				if (lc.count(v)) continue;
				if (w2it.count(v)) lc[v] = w2it.at(v);
				else if (j > start) { if (w2it.count(rw.v_ids[j - 1])) { x = w2it.at(rw.v_ids[j - 1]); hx = true; } if (!hx) throw std::runtime_error("bad optional access"); lc[v] = x; }
			}
			for (u32 j = start + len - 1;; j--) {
				u32 v = rw.v_ids[j];
				if (!rc.count(v)) {
					if (w2it.count(v)) rc[v] = w2it.at(v);
					else if (j + 1 < start + len) { if (w2it.count(rw.v_ids[j + 1])) { y = w2it.at(rw.v_ids[j + 1]); hy = true; } if (!hy) throw std::runtime_error("bad optional access"); rc[v] = y; }
				}
				if (j == start) break;   /* the reference's `j >= start` with u32 j relies on start > 0 */
			}
			for (u32 j = start; j < start + len; j++) {
				u32 v = rw.v_ids[j];
				if (w2it.count(v)) continue;
				auto l_it = lc.at(v), r_it = rc.at(v);
				auto ip = std::distance(sw.begin(), l_it) < std::distance(sw.begin(), r_it) ? r_it : l_it;
				w2it[v] = sw.insert(ip, v);
			}
		}
	}
	return std::vector<u32>(sw.begin(), sw.end());
}

struct RoV { u32 pvst_k; u32 pvst_v; const Flubble *fl; u32 height; std::vector<u32> verts; std::map<u32, u32> pos; u32 component_id; };

/* ---------- overlay (allele.hpp, overlay.cpp) ---------- */
struct Slice { u32 hap, start, len; };
struct ExtStep { u32 idx, v_id; bool fwd; };
typedef std::vector<ExtStep> Lap;

inline IdOr step_at(const Graph &g, u32 hap, u32 i) { return {g.paths[hap].v_ids[i], g.paths[hap].strands[i] == 0}; }
inline Walk to_walk(const Graph &g, const Slice &s) { Walk w; for (u32 i = s.start; i < s.start + s.len; i++) w.push_back(step_at(g, s.hap, i)); return w; }

inline std::vector<Lap> gen_race(const Graph &g, const std::vector<u32> &sorted, u32 h) {   /* untangle.cpp:29-103 */
	Lap un;
	for (u32 v_id : sorted)
		for (u32 i : g.v2steps[g.id2idx.at(v_id)][h]) {
			ExtStep es{i, v_id, g.paths[h].strands[i] == 0};
			auto it = std::upper_bound(un.begin(), un.end(), es, [](const ExtStep &a, const ExtStep &b) { return a.idx < b.idx; });
			un.insert(it, es);
		}
	std::vector<Lap> laps; Lap cur;
	for (size_t k = 0; k < un.size(); k++) {
		cur.push_back(un[k]);
		if (k + 1 == un.size() || un[k].idx + 1 != un[k + 1].idx) { laps.push_back(cur); cur.clear(); }
	}
	return laps;
}

struct Cxt { IdOr l, r; };
inline bool operator<(const Cxt &a, const Cxt &b) { return a.l < b.l || (!(b.l < a.l) && a.r < b.r); }
inline std::string str(const Cxt &c) { return str(c.l) + str(c.r); }

struct MinRov {
	Cxt cxt; Slice ref;
	std::set<u32> haps_matching_ref;
	std::map<Walk, std::vector<Slice>> ins, dels, subs;
	void add_alt(const Graph &g, const Slice &s) {   /* allele.hpp:462-475 */
		if (s.len == ref.len) subs[to_walk(g, s)].push_back(s);
		else if (s.len > ref.len) ins[to_walk(g, s)].push_back(s);
		else dels[to_walk(g, s)].push_back(s);
	}
};
struct Trek { const RoV *rov; bool tangled; std::map<u32, std::map<Cxt, MinRov>> data; };

struct CallStats { uint64_t rovs = 0, tangled_skipped = 0, pin_pairs = 0, too_small = 0, sort_failed = 0, hard_errors = 0; };

inline bool slice_bounds(const Cxt &c, const Lap &lap, u32 &start, u32 &len) {   /* overlay.cpp:189-246 */
	std::vector<u32> starts, ends;
	for (auto &s : lap) { if (s.v_id == c.l.v_id) starts.push_back(s.idx); if (s.v_id == c.r.v_id) ends.push_back(s.idx); }
	if (starts.empty() || ends.empty()) return false;
	std::sort(starts.begin(), starts.end()); std::sort(ends.begin(), ends.end());
	if (starts.front() > ends.front()) std::swap(starts, ends);
	start = starts.front(); len = ends.back() - start + 1;
	return true;
}
inline bool slice_match(const Graph &g, const Slice &a, const Slice &b) {
	if (a.len != b.len) return false;
	for (u32 i = 0; i < a.len; i++) if (step_at(g, a.hap, a.start + i) != step_at(g, b.hap, b.start + i)) return false;
	return true;
}
inline bool is_inv_local(const Graph &g, const Slice &a, const Slice &b) {   /* overlay.cpp:167-187 */
	if (a.len != b.len) return false;
	for (u32 i = 0, j = b.len - 1; i < a.len; i++, j--) {
		IdOr sa = step_at(g, a.hap, a.start + i), sb = step_at(g, b.hap, b.start + j);
		sb.fwd = !sb.fwd;
		if (sa != sb) return false;
	}
	return true;
}

/* non-tangled comp_exps (overlay.cpp:348-588).  Returns false if the reference would exit(EXIT_FAILURE). */
inline bool comp_exps(const Graph &g, const RoV &rov, const std::set<u32> &refs, const std::vector<std::vector<u32>> &dm, Trek &tk, CallStats &st) {
	const u32 I = (u32)dm.size(), J = (u32)rov.verts.size();
	std::map<u32, std::vector<Lap>> races;
	auto race = [&](u32 h) -> const std::vector<Lap> & { if (!races.count(h)) races[h] = gen_race(g, rov.verts, h); return races.at(h); };
	for (u32 ref_h : refs) {
		const std::vector<u32> &rr = dm[ref_h];
		for (u32 h = 0; h < I; h++) {
			if (h == ref_h) continue;
			const std::vector<u32> &ar = dm[h];
			bool empty = true; for (u32 x : ar) if (x) empty = false;
			if (empty) continue;
			/* comp_bounds (overlay.cpp:109-149) */
			std::vector<std::pair<u32, u32>> cxts;
			if (J != 1) {
				std::queue<u32> q;
				for (u32 j = 0; j < J; j++) {
					bool ok = rr[j] > 0 && (rr[j] == ar[j] || (rr[j] == 1 && ar[j] == 2) || (rr[j] == 2 && ar[j] == 1));
					if (ok) q.push(j);
					if (q.size() > 1) {
						u32 u = q.front(); q.pop(); u32 v = q.front();
						for (u32 i = u + 1; i < v; i++) if (rr[i] != ar[i]) { cxts.push_back({u, v}); break; }
					}
				}
			}
			bool inv = J != 1;
			if (inv) for (u32 j = 0; j < J; j++) { bool ci = (rr[j] == 1 && ar[j] == 2) || (rr[j] == 2 && ar[j] == 1); bool nul = rr[j] == 0 && ar[j] == 0; if (!ci && !nul) { inv = false; break; } }
			if (cxts.empty() && !inv) continue;
			const std::vector<Lap> &ref_race = race(ref_h), &alt_race = race(h);
			if (ref_race.empty() || alt_race.empty()) { st.hard_errors++; return false; }
			const Lap &ref_lap = ref_race[0], &alt_lap = alt_race[0];
			if (inv) { st.pin_pairs++; continue; }
			for (auto [bu, bv] : cxts) {
				u32 u = rov.verts[bu], v = rov.verts[bv];
				/* gen_cxt (overlay.cpp:307-327): first u / last v in the ref lap */
				bool hu = false, hv = false; Cxt c{{0, true}, {0, true}};
				for (auto &s : ref_lap) { if (s.v_id == u && !hu) { c.l = {s.v_id, s.fwd}; hu = true; } if (s.v_id == v) { c.r = {s.v_id, s.fwd}; hv = true; } }
				if (!hu || !hv) { st.hard_errors++; return false; }   /* reference: front()/back() on an empty vector */
				std::map<Cxt, MinRov> &m = tk.data[ref_h];
				Slice rs{ref_h, 0, 0}, as{h, 0, 0};
				if (!slice_bounds(c, ref_lap, rs.start, rs.len) || !slice_bounds(c, alt_lap, as.start, as.len)) { st.hard_errors++; return false; }
				if (is_inv_local(g, rs, as)) { st.pin_pairs++; continue; }
				if (slice_match(g, rs, as)) continue;
				if (!m.count(c)) m.insert({c, MinRov{c, rs, {}, {}, {}, {}}});
				m.at(c).add_alt(g, as);
			}
		}
		if (tk.data.empty()) return true;
		if (!tk.data.count(ref_h)) { st.hard_errors++; return false; }   /* reference: std::map::at throws */
		for (auto &[cxt, mr] : tk.data.at(ref_h)) {
			u32 ju_ = rov.pos.at(cxt.l.v_id), jv_ = rov.pos.at(cxt.r.v_id);
			u32 ju = std::min(ju_, jv_), jv = std::max(ju_, jv_);
			const Lap &ref_lap = race(ref_h)[0];
			Slice rs{ref_h, 0, 0};
			if (!slice_bounds(cxt, ref_lap, rs.start, rs.len)) { st.hard_errors++; return false; }
			for (u32 h = 0; h < I; h++) {
				bool same = true; for (u32 j = ju; j <= jv; j++) if (rr[j] != dm[h][j]) { same = false; break; }
				if (!same) continue;
				const std::vector<Lap> &ar = race(h);
				if (ar.empty()) continue;
				Slice as{h, 0, 0};
				if (!slice_bounds(cxt, ar[0], as.start, as.len)) continue;
				if (slice_match(g, rs, as)) mr.haps_matching_ref.insert(h);
			}
		}
	}
	return true;
}

/* ---------- records (vcf.cpp, vcf.hpp, to_vcf.cpp) ---------- */
enum VarType { VT_DEL = 0, VT_INS = 1, VT_SUB = 2, VT_SUBR = 3 };
struct VcfRec {
	u32 ref_h; u32 pos; std::string id, enc; Slice ref; u32 height; VarType vt; bool tangled;
	std::vector<std::vector<Slice>> alts; std::set<u32> ref_at_haps; u32 ns; std::vector<std::vector<std::string>> gt;
};
inline char comp_base(char c) { switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return c; } }
inline std::string step_label(const Graph &g, const IdOr &s) {
	const std::string &l = g.labels[g.id2idx.at(s.v_id)];
	if (s.fwd) return l;
	std::string r; for (auto it = l.rbegin(); it != l.rend(); ++it) r += comp_base(*it);
	return r;
}
inline std::string slice_str(const Graph &g, const Slice &s, VarType vt) {   /* allele.hpp:134-159 */
	u32 i = s.start, N = s.start + s.len;
	if (vt == VT_SUB) { i++; N--; } else if (vt == VT_INS || vt == VT_DEL) N--;
	std::string o; for (; i < N; i++) o += str(step_at(g, s.hap, i));
	return o;
}
inline std::string slice_dna(const Graph &g, const Slice &s, VarType vt) {   /* allele.hpp:182-209 */
	u32 i = s.start, N = s.start + s.len; std::string o;
	if (vt == VT_INS || vt == VT_DEL) o += step_label(g, step_at(g, s.hap, i)).back();
	if (vt != VT_SUBR) { i++; N--; }
	for (; i < N; i++) o += step_label(g, step_at(g, s.hap, i));
	return o;
}
inline u32 slice_pos(const Graph &g, const Slice &s, VarType vt) { u32 l = g.paths[s.hap].loci[s.start + 1]; return (vt == VT_DEL || vt == VT_INS) ? l - 1 : l; }

inline void append_record(const Graph &g, const GtMeta &gm, u32 ref_h, const Cxt &cxt, const RoV &rov, bool tangled, const MinRov &mr,
			  const std::map<Walk, std::vector<Slice>> &wta, VarType vt, std::vector<VcfRec> &recs) {
	if (wta.empty()) return;
	VcfRec r{ref_h, slice_pos(g, mr.ref, vt), str(cxt), flubble_label(*rov.fl), mr.ref, rov.height, vt, tangled, {}, mr.haps_matching_ref, 0, {}};
	for (auto &kv : wta) r.alts.push_back(kv.second);
	r.gt.resize(gm.ploidy.size());
	for (size_t c = 0; c < gm.ploidy.size(); c++) r.gt[c].assign(gm.ploidy[c], ".");
	std::set<u32> ns;
	for (u32 h : mr.haps_matching_ref) { auto [hc, pc] = gm.ref2col.at(h); ns.insert(hc); r.gt[hc][pc] = "0"; }
	u32 i = 1;
	for (auto &kv : wta) { for (auto &s : kv.second) { auto [hc, pc] = gm.ref2col.at(s.hap); r.gt[hc][pc] = std::to_string(i); ns.insert(hc); } i++; }
	r.ns = (u32)ns.size();
	recs.push_back(std::move(r));
}

inline std::string fmt_af(u32 ac, u32 an) {   /* vcf.hpp:310-332: fmt "{:.1f}" of the double quotient (round-half-even on the binary value) */
	char buf[64]; snprintf(buf, sizeof buf, "%.1f", (double)ac / (double)an); return buf;
}
inline std::string record_text(const Graph &g, const VcfRec &r) {   /* to_vcf.cpp:84-119 */
	std::string ac, af, alt_str, alt_dna; u32 an = (u32)r.ref_at_haps.size();
	for (auto &a : r.alts) an += (u32)a.size();
	for (size_t i = 0; i < r.alts.size(); i++) {
		if (i) { ac += ","; af += ","; alt_str += ","; alt_dna += ","; }
		ac += std::to_string(r.alts[i].size()); af += fmt_af((u32)r.alts[i].size(), an);
		alt_str += slice_str(g, r.alts[i].front(), r.vt); alt_dna += slice_dna(g, r.alts[i].front(), r.vt);
	}
	static const char *VT[] = {"DEL", "INS", "SUB", "SUBR"};
	std::ostringstream o;
	o << g.paths[r.ref_h].tag << "\t" << r.pos << "\t" << r.id << "\t" << slice_dna(g, r.ref, r.vt) << "\t" << alt_dna << "\t60\tPASS\t"
	  << "AC=" << ac << ";AF=" << af << ";AN=" << an << ";NS=" << r.ns << ";AT=" << slice_str(g, r.ref, r.vt) << "," << alt_str
	  << ";VARTYPE=" << VT[r.vt] << ";TANGLED=" << (r.tangled ? "T" : "F");
	if (r.vt != VT_SUBR) o << ";ES=" << r.enc << ";LV=" << (r.height - 1);
	o << "\tGT\t";
	for (size_t c = 0; c < r.gt.size(); c++) {
		if (c) o << "\t";
		bool all_missing = true; for (auto &s : r.gt[c]) if (s != ".") all_missing = false;
		if (all_missing) { o << "."; continue; }
		for (size_t k = 0; k < r.gt[c].size(); k++) { if (k) o << "|"; o << r.gt[c][k]; }
	}
	o << "\n";
	return o.str();
}
inline std::string vcf_header(const Graph &g, const GtMeta &gm, const std::vector<std::string> &prefixes, bool with_date) {   /* to_vcf.cpp:19-79 */
	std::string o = "##fileformat=VCFv4.2\n";
	if (with_date) { char d[16]; time_t t = time(0); strftime(d, sizeof d, "%Y%m%d", localtime(&t)); o += std::string("##fileDate=") + d + "\n"; }
	o += "##source=povu\n"
	     "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
	     "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Total number of alternate alleles in called genotypes\">\n"
	     "##INFO=<ID=AT,Number=R,Type=String,Description=\"Allele traversal path through the graph\">\n"
	     "##INFO=<ID=AN,Number=1,Type=String,Description=\"Total number of alleles in called genotypes\">\n"
	     "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele frequency in the population\">\n"
	     "##INFO=<ID=NS,Number=1,Type=Integer,Description=\"Number of samples with data\">\n"
	     "##INFO=<ID=VARTYPE,Number=1,Type=String,Description=\"Type of variation: INS (insertion), DEL (deletion), SUB (substitution), SUBR(substitution in reverse) \">\n"
	     "##INFO=<ID=TANGLED,Number=1,Type=String,Description=\"Variant lies in a tangled region of the graph: T or F\">\n"
	     "##INFO=<ID=LV,Number=1,Type=Integer,Description=\"Level in the PVST (0=top level)\">\n"
	     "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
	for (auto &p : prefixes) for (u32 r : refs_in_sample(g, p)) o += "##contig=<ID=" + g.paths[r].tag + ",length=" + std::to_string(g.paths[r].hap_len) + ">\n";
	o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
	for (auto &c : gm.col_names) o += "\t" + c;
	return o + "\n";
}

/* do_call (call.cpp:71-173) + gen_vcf_rec_map (genomics.cpp:69-164): chunks of `chunk_size` RoVs,
 * records grouped by ref id inside a chunk.  PVSTs are visited in ascending component id (the reference
 * uses unspecified directory order; tests canonicalise). */
inline std::string call_vcf(const Graph &g, const std::vector<Pvst> &pvsts, const std::vector<std::string> &prefixes, size_t chunk_size, bool with_date, CallStats &st) {
	GtMeta gm = genotype_metadata(g);
	std::set<u32> refs;
	for (auto &p : prefixes) for (u32 r : refs_in_sample(g, p)) refs.insert(r);
	std::string out = vcf_header(g, gm, prefixes, with_date);
	std::vector<RoV> rovs;
	for (u32 k = 0; k < pvsts.size(); k++) {
		const Pvst &p = pvsts[k];
		for (u32 pv : color_pvst(g, p, refs)) {
			RoV r; r.pvst_k = k; r.pvst_v = pv; r.fl = &p.fl[pv - 1]; r.height = p.height[pv]; r.component_id = p.component_id;
			std::vector<u32> sorted = bfs_sort(g, *r.fl);
			if (sorted.empty()) { try { sorted = gen_sort(g, *r.fl); } catch (std::exception &) { st.hard_errors++; sorted.clear(); } }
			for (u32 v : sorted) { r.pos[v] = (u32)r.verts.size(); r.verts.push_back(v); }
			if (r.verts.size() < 3) { st.too_small++; continue; }
			rovs.push_back(std::move(r));
		}
	}
	st.rovs = rovs.size();
	const u32 I = (u32)g.paths.size();
	for (size_t base = 0; base < rovs.size(); base += chunk_size) {
		size_t end = std::min(base + chunk_size, rovs.size());
		std::map<u32, std::vector<VcfRec>> by_ref;
		for (size_t i = base; i < end; i++) {
			const RoV &rov = rovs[i];
			const u32 J = (u32)rov.verts.size();
			std::vector<std::vector<u32>> dm(I, std::vector<u32>(J, 0));
			bool tangled = false;
			for (u32 h = 0; h < I; h++)
				for (u32 j = 0; j < J; j++) {   /* allele.hpp:804-862 */
					const std::vector<u32> &idxs = g.v2steps[g.id2idx.at(rov.verts[j])][h];
					u32 depth = (u32)idxs.size();
					if (depth == 0) continue;
					if (depth == 1) { dm[h][rov.pos.at(rov.verts[j])] = g.paths[h].strands[idxs[0]] == 0 ? 1 : 2; continue; }
					if (j == 0 || j == J - 1) tangled = true;
					dm[h][j] = depth;
				}
			if (tangled) { st.tangled_skipped++; continue; }
			Trek tk{&rov, false, {}};
			if (!comp_exps(g, rov, refs, dm, tk, st)) continue;
			for (auto &[ref_h, d] : tk.data) {   /* gen_exp_vcf_recs: INS, SUB, DEL per context (vcf.cpp:241-273) */
				std::vector<VcfRec> &recs = by_ref[ref_h];
				for (auto &[cxt, mr] : d) {
					append_record(g, gm, ref_h, cxt, rov, false, mr, mr.ins, VT_INS, recs);
					append_record(g, gm, ref_h, cxt, rov, false, mr, mr.subs, VT_SUB, recs);
					append_record(g, gm, ref_h, cxt, rov, false, mr, mr.dels, VT_DEL, recs);
				}
			}
		}
		for (auto &[ref_h, recs] : by_ref) for (auto &r : recs) out += record_text(g, r);
	}
	return out;
}

} // namespace oracle
