// povu_b200 -- the reference's two debug surfaces, formatted on the host from what the device pipelines produced:
//   * `--hairpins`:          "Boundary: b1 b2" lines (flubbles.cpp:711-717, printed per component after its cycle-class pass)
//   * `--structure-export`:  the "povu.lean4.structure.v1" JSON document (/root/reference/src/mto/to_structure_export.cpp:148-640)
//                            and its "<path>.flubble-debug.jsonl" side file (flubbles.cpp:108-231), one frame per component.
// Text only: the numbers come from pv_refnum.cuh (tree-edge ids, class ids), the device .pvst writer and the device VCF writer --
// a record's JSON object is re-cut from its VCF line, so both outputs cannot disagree.
#pragma once
#include "pv_engine.hpp"
#include <fstream>
#include <unordered_map>

namespace pve {

inline void json_str(std::string &o, const char *p, size_t n) {
	o += '"';
	for (size_t i = 0; i < n; i++) {
		char ch = p[i];
		switch (ch) {
		case '"': o += "\\\""; break;
		case '\\': o += "\\\\"; break;
		case '\b': o += "\\b"; break;
		case '\f': o += "\\f"; break;
		case '\n': o += "\\n"; break;
		case '\r': o += "\\r"; break;
		case '\t': o += "\\t"; break;
		default: o += ch;
		}
	}
	o += '"';
}
inline void json_str(std::string &o, const std::string &s) { json_str(o, s.data(), s.size()); }
inline void json_key(std::string &o, const char *k) { o += '"'; o += k; o += "\":"; }

// "Boundary: b1 b2" lines of every component with at least three vertices, in component order
inline std::string hairpin_text(const Session &S) {
	if (!S.refnum.valid) throw std::runtime_error("hairpins: decompose did not run with the debug numbering on");
	std::string o;
	for (size_t i = 0; i < S.refnum.hp_comp.size(); i++) {
		const u32 k = S.refnum.hp_comp[i];
		if (S.dec.comp_nv[k] < 3) continue;      // decompose.cpp:135-142: too small, never reaches the cycle-class pass
		const u32 b1 = S.refnum.hp_b1[i], b2 = S.refnum.hp_b2[i];
		o += "Boundary: " + std::to_string(S.g.ids[b1]) + " " + (b2 == pv::NONE ? std::string("4294967295") : std::to_string(S.g.ids[b2])) + "\n";
	}
	return o;
}

// one "povu.flubble-debug.frame.v1" line per component with at least three vertices (flubbles.cpp:108-213)
inline std::string flubble_debug_frames(const Session &S) {
	if (!S.refnum.valid || S.dec.stack_v.size() != S.refnum.class_id.size()) throw std::runtime_error("structure export: decompose did not run with the debug numbering on");
	if (S.refnum.errors) throw std::runtime_error("structure export: inconsistent bracket lists (internal error)");
	std::string o;
	size_t base = 0;
	for (u32 k = 0; k < S.dec.comp_nv.size(); k++) {
		const size_t n = S.dec.comp_nv[k];
		if (n >= 3) {
			const size_t nvtx = 2 * n + (S.dec.comp_has_tips[k] ? 1 : 0);
			o += "{\"schema\":\"povu.flubble-debug.frame.v1\",\"tree_vertex_count\":" + std::to_string(nvtx) + ",\"tree_edge_count\":" + std::to_string(nvtx - 1) + ",\"stack_entries\":[";
			// expected_next_seen: the next later entry of the same class, else the entry itself (flubbles.cpp:94-106)
			std::vector<u32> expect(n);
			{ std::unordered_map<u32, u32> seen; seen.reserve(n * 2); for (size_t i = n; i-- > 0;) { u32 cl = S.refnum.class_id[base + i]; auto it = seen.find(cl); expect[i] = it == seen.end() ? (u32)i : it->second; seen[cl] = (u32)i; } }
			for (size_t i = 0; i < n; i++) {
				const size_t gi = base + i;
				const u32 cls = S.refnum.class_id[gi], eidx = S.refnum.edge_idx[gi], actual = S.dec.stack_next[gi] - (u32)base;
				const bool in_range = actual < n, same = in_range && S.refnum.class_id[base + actual] == cls, ok = actual == expect[i];
				if (i) o += ',';
				o += "{\"order\":" + std::to_string(i) + ",\"tree_edge_index\":" + std::to_string(eidx) + ",\"tree_edge_id\":" + std::to_string(S.refnum.edge_id[gi]) +
				     ",\"boundary_vertex_id\":" + std::to_string(S.g.ids[S.dec.stack_v[gi]]) + ",\"orientation\":\"" + (S.dec.stack_fwd[gi] ? ">" : "<") +
				     "\",\"provenance\":\"real\",\"color\":\"black\",\"class_id\":" + std::to_string(cls) + ",\"parent_tree_vertex\":" + std::to_string(eidx) +
				     ",\"child_tree_vertex\":" + std::to_string(eidx + 1) + ",\"next_seen\":" + std::to_string(actual) + ",\"expected_next_seen\":" + std::to_string(expect[i]) +
				     ",\"next_seen_in_range\":" + (in_range ? "true" : "false") + ",\"next_seen_same_class\":" + (same ? "true" : "false") + ",\"diagnostic\":\"" +
				     (ok ? "ok" : "mismatch: next_seen does not match the next later stack entry with the same class") + "\"}";
			}
			o += "],\"next_seen_table\":[";
			for (size_t i = 0; i < n; i++) {
				const u32 actual = S.dec.stack_next[base + i] - (u32)base;
				if (i) o += ',';
				o += "{\"stack_order\":" + std::to_string(i) + ",\"class_id\":" + std::to_string(S.refnum.class_id[base + i]) + ",\"next_seen\":" + std::to_string(actual) +
				     ",\"expected_next_seen\":" + std::to_string(expect[i]) + ",\"diagnostic\":\"" +
				     (actual == expect[i] ? "ok" : "mismatch: next_seen table entry is not the next same-class stack index") + "\"}";
			}
			o += "]}\n";
		}
		base += n;
	}
	return o;
}

struct PvstNode { char fam; std::string label, start, end; bool has_route; char route; std::vector<u32> children; u32 parent = pv::NONE, depth = 0; };
inline const char *family_name(char f) {
	switch (f) { case 'D': return "dummy"; case 'F': return "flubble"; case 'T': return "tiny"; case 'O': return "parallel"; case 'C': return "concealed"; case 'S': return "smothered"; case 'M': return "midi"; }
	return "unknown";
}
// the nodes of one .pvst text in file order (what from_pvst.cpp:162-302 rebuilds): family, label, children in listed order
inline std::vector<PvstNode> parse_pvst_nodes(const std::string &text) {
	std::vector<PvstNode> nodes; std::unordered_map<u32, u32> file_idx;
	std::vector<std::vector<u32>> kids_by_file;
	size_t p = 0;
	while (p < text.size()) {
		size_t e = text.find('\n', p); if (e == std::string::npos) e = text.size();
		std::vector<std::string> tok; size_t q = p;
		while (q <= e) { size_t t = text.find('\t', q); if (t == std::string::npos || t > e) t = e; tok.push_back(text.substr(q, t - q)); q = t + 1; }
		p = e + 1;
		if (tok.size() < 5 || tok[0] == "H") continue;
		PvstNode nd; nd.fam = tok[0][0]; nd.label = tok[2]; nd.has_route = tok[4] != "."; nd.route = nd.has_route ? tok[4][0] : '.';
		if (nd.fam != 'D') {   // label = two oriented ids
			size_t cut = nd.label.find_first_of("<>", 1);
			if (cut == std::string::npos) throw std::runtime_error("structure export: unexpected PVST label " + nd.label);
			nd.start = nd.label.substr(0, cut); nd.end = nd.label.substr(cut);
		}
		std::vector<u32> kids;
		if (tok[3] != ".") { size_t a = 0; while (a < tok[3].size()) { size_t b = tok[3].find(',', a); if (b == std::string::npos) b = tok[3].size(); kids.push_back((u32)strtoul(tok[3].substr(a, b - a).c_str(), nullptr, 10)); a = b + 1; } }
		file_idx[(u32)strtoul(tok[1].c_str(), nullptr, 10)] = (u32)nodes.size();
		kids_by_file.push_back(std::move(kids)); nodes.push_back(std::move(nd));
	}
	for (size_t i = 0; i < nodes.size(); i++) for (u32 cf : kids_by_file[i]) { u32 ci = file_idx.at(cf); nodes[i].children.push_back(ci); nodes[ci].parent = (u32)i; }
	// heights = distance from the root (pvst.hpp:807-836)
	std::vector<u32> order; for (u32 i = 0; i < nodes.size(); i++) if (nodes[i].parent == pv::NONE) order.push_back(i);
	for (size_t h = 0; h < order.size(); h++) for (u32 ci : nodes[order[h]].children) { nodes[ci].depth = nodes[order[h]].depth + 1; order.push_back(ci); }
	return nodes;
}

// the whole document.  Runs after call(); `input_name` = file name of the GFA; chunk_size = -c of the CLI (record order).
inline void structure_export(Session &S, const std::string &path, const std::string &input_name, size_t chunk_size, const std::string &frames_jsonl) {
	// forests written by the reference's `decompose -s` name vertices under several parents or under none; the tables below are
	// rebuilt from one parent per vertex and would not reproduce that -- refuse instead of writing something else
	for (char f : S.forest.family) if (f == 'C' || f == 'S' || f == 'M') throw std::runtime_error("structure export: forests with concealed / smothered / midi vertices (decompose -s) are not supported");
	if (!S.forest.orphan.empty()) throw std::runtime_error("structure export: the forest has vertices that no children list reaches (decompose -s); not supported");
	std::ofstream out(path, std::ios::binary);
	if (!out.is_open()) throw std::runtime_error("could not open structure export path: " + path);
	S.ensure_host(true, true, true);
	const pvh::HostGraph &g = S.g;
	std::string o;
	o.reserve(1 << 20);
	o += "{\"schema\":\"povu.lean4.structure.v1\",\"producer\":\"current-povu-cli\",\"accepted_gfa\":{\"input_name\":"; json_str(o, input_name);
	o += ",\"segments\":[";
	for (u32 v = 0; v < g.V(); v++) {
		if (v) o += ',';
		o += "{\"order\":" + std::to_string(v) + ",\"id\":" + std::to_string(g.ids[v]) + ",\"sequence\":";
		if (g.has_labels) json_str(o, g.lbl.data() + g.lblOff[v], (size_t)(g.lblOff[v + 1] - g.lblOff[v])); else o += "\"\"";
		o += '}';
		if (o.size() > (8u << 20)) { out << o; o.clear(); }
	}
	o += "],\"links\":[";
	for (u64 e = 0; e < g.E(); e++) {
		if (e) o += ',';
		const u32 a = g.ep1[e], b = g.ep2[e];
		o += "{\"order\":" + std::to_string(e) + ",\"from\":" + std::to_string(g.ids[a >> 1]) + ",\"from_orient\":\"" + ((a & 1u) ? "+" : "-") + "\",\"to\":" + std::to_string(g.ids[b >> 1]) +
		     ",\"to_orient\":\"" + ((b & 1u) ? "-" : "+") + "\",\"overlap\":\"0M\"}";
		if (o.size() > (8u << 20)) { out << o; o.clear(); }
	}
	o += "],\"paths\":[";
	for (u32 h = 0; h < g.H(); h++) {
		if (h) o += ',';
		o += "{\"order\":" + std::to_string(h) + ",\"name\":"; json_str(o, g.paths[h].tag); o += ",\"sample\":"; json_str(o, g.paths[h].sample); o += ",\"steps\":[";
		for (u32 s = g.stepOff[h]; s < g.stepOff[h + 1]; s++) { if (s > g.stepOff[h]) o += ','; u32 x = g.steps[s]; o += '"'; o += (x & 1u) ? '<' : '>'; o += std::to_string(g.ids[x >> 1]); o += '"'; }
		o += "]}";
		if (o.size() > (8u << 20)) { out << o; o.clear(); }
	}
	o += "]},\"reference_prefixes\":[";
	for (size_t i = 0; i < S.prefixes.size(); i++) { if (i) o += ','; json_str(o, S.prefixes[i]); }
	o += "],";

	// PVST nodes, trees in ascending component order (position in that order + 1 names the tree: to_structure_export.cpp:208-213)
	const u32 T = (u32)S.forest.comp_id.size();
	std::vector<std::vector<PvstNode>> trees(T);
	for (u32 t = 0; t < T; t++) trees[t] = parse_pvst_nodes(S.pvst_text(t));
	auto node_id = [](u32 t, u32 i) { return std::to_string(t + 1) + ":" + std::to_string(i); };
	auto route_name = [](char r) { return r == 'L' ? "s2e" : "e2s"; };
	o += "\"boundary_candidates\":[";
	{
		size_t order = 0;
		for (u32 t = 0; t < T; t++) for (u32 i = 0; i < trees[t].size(); i++) {
			const PvstNode &nd = trees[t][i];
			if (!nd.has_route) continue;
			if (order) o += ',';
			o += "{\"order\":" + std::to_string(order++) + ",\"kind\":\"pvst-route-boundary\",\"node_id\":\"" + node_id(t, i) + "\",\"family\":\"" + family_name(nd.fam) + "\",\"start\":";
			json_str(o, nd.start); o += ",\"end\":"; json_str(o, nd.end); o += std::string(",\"route\":\"") + route_name(nd.route) + "\",\"boundary\":"; json_str(o, nd.start + nd.end); o += '}';
		}
	}
	o += "],\"pvst_nodes\":[";
	{
		size_t order = 0;
		for (u32 t = 0; t < T; t++) for (u32 i = 0; i < trees[t].size(); i++) {
			const PvstNode &nd = trees[t][i];
			if (order) o += ',';
			o += "{\"order\":" + std::to_string(order++) + ",\"component\":" + std::to_string(t + 1) + ",\"local_index\":" + std::to_string(i) + ",\"node_id\":\"" + node_id(t, i) +
			     "\",\"family\":\"" + family_name(nd.fam) + "\",\"type\":\"" + family_name(nd.fam) + "\",\"label\":"; json_str(o, nd.label);
			if (nd.has_route) {
				o += ",\"start\":"; json_str(o, nd.start); o += ",\"end\":"; json_str(o, nd.end); o += std::string(",\"route\":\"") + route_name(nd.route) + "\",\"boundary\":"; json_str(o, nd.start + nd.end);
			} else o += ",\"start\":null,\"end\":null,\"route\":null,\"boundary\":null";
			o += ",\"parent\":";
			if (nd.parent == pv::NONE) o += "null"; else o += "\"" + node_id(t, nd.parent) + "\"";
			o += ",\"children\":[";
			for (size_t c = 0; c < nd.children.size(); c++) { if (c) o += ','; o += "\"" + node_id(t, nd.children[c]) + "\""; }
			o += "],\"tree_depth\":" + std::to_string(nd.depth) + "}";
			if (o.size() > (8u << 20)) { out << o; o.clear(); }
		}
	}
	o += "],\"flubble_debug\":{\"schema\":\"povu.flubble-debug.v1\",\"frames\":[";
	{
		bool first = true; size_t p = 0;
		while (p < frames_jsonl.size()) {
			size_t e = frames_jsonl.find('\n', p); if (e == std::string::npos) e = frames_jsonl.size();
			if (e > p) { if (!first) o += ','; first = false; o.append(frames_jsonl, p, e - p); }
			p = e + 1;
		}
	}
	o += "]},\"variant_calls\":[";
	out << o; o.clear();

	// variant calls: one object per VCF record line, in the order of the combined stream
	if (S.rec.n()) {
		pv::TextOut body;
		S.vcf_records_device(chunk_size, nullptr, body);
		S.rec.to_host(*S.c);
		const std::vector<size_t> ord = pvh::record_order(S.rec, chunk_size);
		const pvh::GtMeta &gm = S.gm;
		const char *txt = body.text.data();
		for (size_t n = 0; n < ord.size(); n++) {
			const size_t i = ord[n];
			const char *lb = txt + body.lineOff[n], *le = txt + body.lineOff[n + 1];
			if (le > lb && le[-1] == '\n') le--;
			std::vector<std::pair<const char *, const char *>> col;
			for (const char *q = lb; q <= le;) { const char *tb = (const char *)memchr(q, '\t', (size_t)(le - q)); if (!tb) tb = le; col.push_back({q, tb}); q = tb + 1; }
			if (col.size() < 9) throw std::runtime_error("structure export: malformed VCF line");
			auto cs = [&](size_t k) { return std::string(col[k].first, col[k].second); };
			auto split = [](const std::string &s, char d) { std::vector<std::string> r; size_t a = 0; while (a <= s.size()) { size_t b = s.find(d, a); if (b == std::string::npos) b = s.size(); r.push_back(s.substr(a, b - a)); a = b + 1; } return r; };
			std::unordered_map<std::string, std::string> info;
			for (auto &kv : split(cs(7), ';')) { size_t eq = kv.find('='); if (eq != std::string::npos) info[kv.substr(0, eq)] = kv.substr(eq + 1); }
			const bool subr = info["VARTYPE"] == "SUBR";
			const std::vector<std::string> ac = split(info["AC"], ','), at = split(info["AT"], ','), alts = split(cs(4), ',');
			const u64 an = strtoull(info["AN"].c_str(), nullptr, 10);
			u64 ac_sum = 0; for (auto &x : ac) ac_sum += strtoull(x.c_str(), nullptr, 10);
			if (n) o += ',';
			o += "{\"order\":" + std::to_string(n) + ",\"source\":{\"kind\":\"" + (subr ? "hairpin" : "flubble") + "\",\"node_id\":";
			if (subr) o += "null,\"family\":null";
			else {
				const u32 r = S.rec.rec_rov[i], fl = S.rec.rov_fl[r], t = S.tree_of_flubble(fl), li = fl - S.forest.treeOff[t] + 1;
				o += "\"" + node_id(t, li) + "\",\"family\":\"" + family_name(trees[t].at(li).fam) + "\"";
			}
			o += ",\"endpoint_id\":"; json_str(o, subr ? cs(2) : info["ES"]);
			o += ",\"enclosing_site\":"; if (subr) o += "null"; else json_str(o, info["ES"]);
			o += ",\"level\":"; if (subr) o += "null"; else o += info["LV"];
			o += "},\"chrom\":"; json_str(o, cs(0));
			o += ",\"contig_order\":" + std::to_string(S.rec.rec_ref[i]) + ",\"pos\":" + cs(1) + ",\"id\":"; json_str(o, cs(2));
			o += ",\"ref\":"; json_str(o, cs(3)); o += ",\"ref_traversal\":"; json_str(o, at.empty() ? std::string() : at[0]);
			o += ",\"alternates\":[";
			for (size_t a = 0; a < ac.size(); a++) {
				if (a) o += ',';
				o += "{\"index\":" + std::to_string(a + 1) + ",\"dna\":"; json_str(o, a < alts.size() ? alts[a] : std::string());
				o += ",\"traversal\":"; json_str(o, a + 1 < at.size() ? at[a + 1] : std::string()); o += ",\"count\":" + ac[a] + "}";
			}
			o += "],\"variant_type\":"; json_str(o, info["VARTYPE"]);
			o += std::string(",\"tangled\":") + (info["TANGLED"] == "T" ? "true" : "false") + ",\"qual\":"; json_str(o, cs(5));
			o += ",\"filter\":"; json_str(o, cs(6)); o += ",\"format\":"; json_str(o, cs(8));
			o += ",\"reference_allele_count\":" + std::to_string(an - ac_sum) + ",\"ac\":[";
			for (size_t a = 0; a < ac.size(); a++) { if (a) o += ','; o += ac[a]; }
			o += "],\"af\":[";
			for (size_t a = 0; a < ac.size(); a++) { if (a) o += ','; o += "\"" + ac[a] + "/" + std::to_string(an) + "\""; }
			o += "],\"an\":" + std::to_string(an) + ",\"ns\":" + info["NS"] + ",\"genotypes\":[";
			for (size_t k = 0; k < gm.col_names.size(); k++) {
				if (k) o += ',';
				o += "{\"sample\":"; json_str(o, gm.col_names[k]); o += ",\"value\":"; json_str(o, 9 + k < col.size() ? cs(9 + k) : std::string()); o += '}';
			}
			o += "]}";
			if (o.size() > (8u << 20)) { out << o; o.clear(); }
		}
	}
	o += "]}";
	out << o;
	out.close();
}

} // namespace pve
