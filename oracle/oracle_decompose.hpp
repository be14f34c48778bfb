/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into or called by the product
 * library povu_b200/.  CPU restatement (sequential, data-structure-for-data-structure
 * simplified) of the reference's decompose half.  Pinned against oracle/_ref (the
 * unmodified reference compiled here) and the reference's golden vectors by
 * tests/test_oracle_*.py.
 *
 * Follows:
 *   componetize ................. /root/reference/src/povu/graph/bidirected.cpp:477-602
 *   spanning tree (from_bd) ..... /root/reference/src/povu/graph/spanning_tree.cpp:262-463
 *   bracket list ops ............ /root/reference/src/povu/graph/bracket_list.cpp:61-100,
 *                                 spanning_tree.cpp:821-881
 *   handle_vertex / cycle equiv . /root/reference/src/povu/algorithms/flubbles.cpp:503-719
 *   class stack (net effect) .... flubbles.cpp:412-501 + tree_utils.cpp:19-157
 *   next_seen ................... flubbles.cpp:375-410
 *   add_flubbles ................ flubbles.cpp:233-367
 *   .pvst writer ................ /root/reference/src/mto/to_pvst.cpp:23-109
 */
#pragma once
#include <list>
#include <unordered_map>
#include <unordered_set>
#include "oracle_graph.hpp"

namespace oracle {

static const u32 INVALID = 0xFFFFFFFFu;

/* ---------- components (bidirected.cpp:477-602) ---------- */
struct Component {
	Graph g;                 // vertices ascending original idx; edges renumbered
};

inline std::vector<Component> componetize(const Graph &g) {
	std::vector<Component> out;
	size_t V = g.ids.size();
	if (V == 0) return out;
	std::vector<char> visited(V, 0);
	size_t next_start = 0;
	while (true) {
		while (next_start < V && visited[next_start]) next_start++;
		if (next_start >= V) break;
		std::set<u32> comp;
		std::vector<u32> st{(u32)next_start};
		visited[next_start] = 1; comp.insert((u32)next_start);
		while (!st.empty()) {
			u32 v = st.back(); st.pop_back();
			for (int pass = 0; pass < 2; pass++)
				for (u32 e : (pass == 0 ? g.e_l[v] : g.e_r[v])) {
					u32 w = g.other_v(e, v);
					if (visited[w]) continue;
					visited[w] = 1; st.push_back(w); comp.insert(w);
				}
		}
		Component c;
		for (u32 v : comp) c.g.add_vertex(g.ids[v], g.labels[v]);
		std::unordered_set<u32> added;
		for (u32 v : comp)
			for (u32 side = 0; side < 2; side++)
				for (u32 e : (side == SIDE_L ? g.e_l[v] : g.e_r[v])) {
					if (added.count(e)) continue;
					added.insert(e);
					auto [os, ov] = g.other(e, v, side);   // self-loops become (side, complement)
					c.g.add_edge(g.ids[v], side, g.ids[ov], os);
				}
		for (auto &t : g.tips) if (comp.count(g.id2idx.at(t.first))) c.g.tips.insert(t);
		out.push_back(std::move(c));
	}
	return out;
}

/* ---------- spanning tree (spanning_tree.cpp:262-463) ---------- */
enum { T_L = 0, T_R = 1, T_DUMMY = 2 };
struct BackEdge { u32 id, src, tgt; int type; u32 cls; };   // type 0 back, 1 capping, 2 simplifying
struct TreeEdge { u32 id, parent, child; bool black; u32 cls; };
struct TVertex {
	u32 g_v_id; int type; u32 parent_e; std::vector<u32> child_e;
	std::vector<u32> obe, ibe; u32 hi; u32 pre, post;
};
struct STree {
	std::vector<TVertex> nodes; std::vector<TreeEdge> tedges; std::vector<BackEdge> bedges;
	u32 add_vertex(u32 gid, int type) { nodes.push_back({gid, type, INVALID, {}, {}, {}, INVALID, 0, 0}); return (u32)nodes.size() - 1; }
	void add_tree_edge(u32 f, u32 t, bool black) {
		u32 idx = (u32)tedges.size();
		tedges.push_back({idx + (u32)bedges.size(), f, t, black, INVALID});
		nodes[f].child_e.push_back(idx); nodes[t].parent_e = idx;
	}
	u32 add_be(u32 f, u32 t, int type) {
		u32 idx = (u32)bedges.size();
		bedges.push_back({idx + (u32)tedges.size(), f, t, type, INVALID});
		nodes[f].obe.push_back(idx); nodes[t].ibe.push_back(idx);
		return idx;
	}
	std::vector<u32> children(u32 v) const { std::vector<u32> c; for (u32 e : nodes[v].child_e) c.push_back(tedges[e].child); std::sort(c.begin(), c.end()); return c; }
};

inline STree spanning_tree(const Graph &g) {
	STree t;
	const bool has_tips = !g.tips.empty();
	size_t V = g.ids.size();
	std::vector<char> visited(V, 0);
	std::set<std::pair<u32, u32>> connected;
	std::unordered_set<u32> self_loops;
	std::vector<u32> be2ctr(2 * V + 1, 0);
	auto upair = [](u32 a, u32 b) { return std::make_pair(std::min(a, b), std::max(a, b)); };
	u32 p_idx = INVALID, order = 0;
	auto add_vertex_to_tree = [&](u32 side, u32 v) {
		u32 a = t.add_vertex(g.ids[v], side == SIDE_L ? T_L : T_R); t.nodes[a].pre = order++;
		u32 b = t.add_vertex(g.ids[v], side == SIDE_L ? T_R : T_L); t.nodes[b].pre = order++;
		be2ctr[2 * v + side] = a; be2ctr[2 * v + (side ^ 1)] = b;
		if (p_idx != INVALID) { t.add_tree_edge(p_idx, a, false); connected.insert(upair(p_idx, a)); }
		t.add_tree_edge(a, b, true); connected.insert(upair(a, b));
	};
	std::vector<u32> st;   // stack of bi-edged idxs
	if (has_tips) { p_idx = t.add_vertex(INVALID, T_DUMMY); t.nodes[p_idx].pre = order++; }
	u32 s_side, s_v;
	if (has_tips) { s_side = g.tips.begin()->second; s_v = g.id2idx.at(g.tips.begin()->first); }
	else { s_side = SIDE_L; s_v = 0; }
	st.push_back(2 * s_v + s_side); st.push_back(2 * s_v + (s_side ^ 1));
	visited[s_v] = 1;
	add_vertex_to_tree(s_side, s_v);
	while (!st.empty()) {
		u32 be = st.back(); p_idx = be2ctr[be];
		u32 v = be / 2, side = be % 2;
		const std::set<u32> &nb = side == SIDE_L ? g.e_l[v] : g.e_r[v];
		if (nb.empty() && !connected.count(upair(p_idx, 0))) { t.add_be(p_idx, 0, 0); connected.insert(upair(p_idx, 0)); }
		bool found = false;
		for (u32 e : nb) {
			auto [os, ov] = g.other(e, v, side);
			u32 o_be = 2 * ov + os;
			if (!visited[ov]) {
				add_vertex_to_tree(os, ov); visited[ov] = 1;
				st.push_back(2 * ov + os); st.push_back(2 * ov + (os ^ 1));
				found = true; break;
			} else if (!connected.count(upair(p_idx, be2ctr[o_be]))) {
				t.add_be(p_idx, be2ctr[o_be], 0); connected.insert(upair(p_idx, be2ctr[o_be]));
			} else if (v == ov && !self_loops.count(v)) {
				t.add_be(p_idx, be2ctr[o_be], 0); self_loops.insert(v);
			}
		}
		if (!found) { t.nodes[p_idx].post = order++; st.pop_back(); }
	}
	if (has_tips) t.nodes[0].post = order++;
	return t;
}

/* ---------- cycle equivalence (flubbles.cpp:503-719) ---------- */
struct Bracket { u32 be_id; u32 recent_size; u32 recent_class; };
struct BList { std::list<Bracket> l; };   /* iterators survive splice, so one global id->iterator map replaces the per-list maps */

inline void cycle_equiv(STree &t) {
	size_t N = t.nodes.size();
	std::vector<BList *> bl(N, nullptr);
	u32 n_classes = 0;
	std::unordered_map<u32, std::list<Bracket>::iterator> where;
	auto push = [&](u32 v, u32 be_idx) {
		if (!bl[v]) bl[v] = new BList();
		bl[v]->l.push_front({t.bedges[be_idx].id, INVALID, INVALID});
		where[t.bedges[be_idx].id] = bl[v]->l.begin();
	};
	std::unordered_map<u32, u32> beid2idx;
	for (u32 i = 0; i < t.bedges.size(); i++) beid2idx[t.bedges[i].id] = i;
	for (u32 v = (u32)N; v-- > 0;) {
		u32 hi_0 = INVALID;
		for (u32 b : t.nodes[v].obe) hi_0 = std::min(hi_0, t.bedges[b].tgt);
		std::vector<u32> ch = t.children(v);
		u32 hi_1 = INVALID;
		for (u32 c : ch) hi_1 = std::min(hi_1, t.nodes[c].hi);
		t.nodes[v].hi = std::min(hi_0, hi_1);
		u32 hi_child = INVALID;
		for (u32 c : ch) if (t.nodes[c].hi == hi_1) { hi_child = c; break; }
		u32 hi_2 = INVALID;
		for (u32 c : ch) if (c != hi_child && t.nodes[c].hi < v) { hi_2 = t.nodes[c].hi; break; }
		for (u32 c : ch) {                       /* concat: child list spliced to the front */
			if (!bl[v]) { bl[v] = bl[c]; bl[c] = nullptr; }
			else { bl[v]->l.splice(bl[v]->l.begin(), bl[c]->l); delete bl[c]; bl[c] = nullptr; }
		}
		std::vector<u32> ibe = t.nodes[v].ibe; std::sort(ibe.begin(), ibe.end());
		for (u32 b : ibe) {
			u32 id = t.bedges[b].id;
			if (bl[v] && where.count(id)) { bl[v]->l.erase(where[id]); where.erase(id); }
			if (t.bedges[b].type != 1 && t.bedges[b].cls == INVALID) t.bedges[b].cls = n_classes++;
		}
		std::vector<u32> obe = t.nodes[v].obe; std::sort(obe.begin(), obe.end());
		for (u32 b : obe) push(v, b);
		if (hi_2 < hi_0) { u32 b = t.add_be(v, hi_2, 1); beid2idx[t.bedges[b].id] = b; push(v, b); }
		if (!bl[v]) throw std::runtime_error("Bracket list is null");
		if (bl[v]->l.empty()) { u32 b = t.add_be(v, 0, 2); beid2idx[t.bedges[b].id] = b; push(v, b); t.nodes[v].hi = 0; }
		if (v != 0) {
			Bracket &b = bl[v]->l.front();
			if ((u32)bl[v]->l.size() != b.recent_size) { b.recent_size = (u32)bl[v]->l.size(); b.recent_class = n_classes++; }
			t.tedges[t.nodes[v].parent_e].cls = b.recent_class;
			if (b.recent_size == 1) t.bedges[beid2idx.at(b.be_id)].cls = b.recent_class;
		}
	}
	for (auto *p : bl) delete p;
}

/* ---------- class stack (flubbles.cpp:412-501; net effect: preorder listing of black
 * tree edges with, at every vertex, gray children in ascending index and the black child
 * last) ---------- */
struct StackEntry { bool forward; u32 v_id; u32 tree_edge; u32 cls; };

inline std::vector<StackEntry> class_stack(const STree &t) {
	std::vector<StackEntry> out;
	std::vector<u32> st{0};
	while (!st.empty()) {
		u32 v = st.back(); st.pop_back();
		if (t.nodes[v].parent_e != INVALID) {
			const TreeEdge &pe = t.tedges[t.nodes[v].parent_e];
			if (pe.black) out.push_back({t.nodes[v].type == T_R, t.nodes[v].g_v_id, t.nodes[v].parent_e, pe.cls});
		}
		std::vector<u32> gray, black;
		for (u32 e : t.nodes[v].child_e) (t.tedges[e].black ? black : gray).push_back(t.tedges[e].child);
		std::sort(gray.begin(), gray.end());
		/* visit order: gray ascending, then black -> push in reverse */
		for (u32 c : black) st.push_back(c);
		for (size_t i = gray.size(); i-- > 0;) st.push_back(gray[i]);
	}
	return out;
}

/* ---------- next_seen + add_flubbles (flubbles.cpp:233-410) ---------- */
struct Flubble { u32 a_id; bool a_fwd; u32 z_id; bool z_fwd; u32 parent; };   // parent = PVST idx (0 = dummy root)

inline std::vector<Flubble> add_flubbles(const std::vector<StackEntry> &s) {
	size_t n = s.size();
	std::vector<u32> next_seen(n);
	std::unordered_map<u32, u32> nm;
	for (size_t i = n; i-- > 0;) { next_seen[i] = nm.count(s[i].cls) ? nm[s[i].cls] : (u32)i; nm[s[i].cls] = (u32)i; }
	std::vector<Flubble> fl;             // fl[k] is PVST vertex k+1
	std::vector<u32> stk; std::unordered_set<u32> in_s;
	u32 prt = 0;
	for (u32 i = 0; i < n; i++) {
		u32 cl = s[i].cls;
		if (in_s.count(cl)) {
			while (!stk.empty()) { u32 c = stk.back(); stk.pop_back(); in_s.erase(c); if (c == cl) break; }
			if (prt != 0) prt = fl[prt - 1].parent;
		}
		if (i + 1 < next_seen[i]) {
			const StackEntry &a = s[i], &z = s[next_seen[i]];
			Flubble f;
			if (!a.forward && !z.forward) f = {z.v_id, true, a.v_id, true, prt};   /* normalize_endpoints :233-243 */
			else f = {a.v_id, a.forward, z.v_id, z.forward, prt};
			fl.push_back(f);
			prt = (u32)fl.size();
		}
		stk.push_back(cl); in_s.insert(cl);
	}
	return fl;
}

inline std::string flubble_label(const Flubble &f) {
	return std::string(f.a_fwd ? ">" : "<") + std::to_string(f.a_id) + (f.z_fwd ? ">" : "<") + std::to_string(f.z_id);
}

/* ---------- .pvst v0.0.3 text (to_pvst.cpp:23-109) ---------- */
inline std::string pvst_text(const std::vector<Flubble> &fl) {
	std::vector<std::vector<u32>> ch(fl.size() + 1);
	for (u32 k = 0; k < fl.size(); k++) ch[fl[k].parent].push_back(k + 1);
	std::string o = "H\t0.0.3\t.\t.\t.\n";
	for (u32 i = 0; i <= fl.size(); i++) {
		o += i == 0 ? "D\t" : "F\t";
		o += std::to_string(i) + "\t" + (i == 0 ? std::string(".") : flubble_label(fl[i - 1])) + "\t";
		if (ch[i].empty()) o += ".";
		else for (size_t k = 0; k < ch[i].size(); k++) { if (k) o += ", "; o += std::to_string(ch[i][k]); }
		o += i == 0 ? "\t.\n" : "\tL\n";
	}
	return o;
}

struct DecomposedComponent { u32 component_id; u32 n_vertices; std::vector<StackEntry> stack; std::vector<Flubble> flubbles; };

/* do_decompose (app/subcommand/decompose.cpp:94-160): components with < 3 vertices are skipped */
inline std::vector<DecomposedComponent> decompose(const Graph &g) {
	std::vector<DecomposedComponent> out;
	std::vector<Component> comps = componetize(g);
	for (size_t i = 0; i < comps.size(); i++) {
		if (comps[i].g.ids.size() < 3) continue;
		STree t = spanning_tree(comps[i].g);
		cycle_equiv(t);
		DecomposedComponent d; d.component_id = (u32)i + 1; d.n_vertices = (u32)comps[i].g.ids.size();
		d.stack = class_stack(t);
		d.flubbles = add_flubbles(d.stack);
		out.push_back(std::move(d));
	}
	return out;
}

} // namespace oracle
