/* TEST INFRASTRUCTURE ONLY (oracle) -- never linked into or called by the product
 * library povu_b200/.  CPU restatement of the reference's graph loading step.
 *
 * Follows:
 *   GFA pre-validation ............ /root/reference/src/mto/from_gfa.cpp:34-98
 *   liteseq parse rules ........... SURVEY.md Appendix A (liteseq itself is an absent
 *                                   third-party dependency, github.com/urbanslug/liteseq
 *                                   @3678c8b7, cmake/deps.cmake:26-33; "parity unpinned"
 *                                   beyond the rules the fixtures exercise)
 *   to_bd (vertices/edges/tips) ... /root/reference/src/mto/from_gfa.cpp:141-280
 *   VG::add_edge .................. /root/reference/src/povu/graph/bidirected.cpp:321-340
 *   set_vtx_ref_idx ............... /root/reference/src/povu/graph/bidirected.cpp:347-355
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

typedef uint32_t u32;
enum { SIDE_L = 0, SIDE_R = 1 };

struct Edge { u32 v1, s1, v2, s2; };            // vertex *indices* and sides
struct Path {
	std::string tag, sample; u32 hap_id; bool pansn; u32 hap_len;
	std::vector<u32> v_ids, strands, loci;        // strand 0 = fwd; loci 1-based start
};

struct Graph {
	std::vector<u32> ids;                          // v_idx -> v_id (ascending)
	std::vector<std::string> labels;
	std::map<u32, u32> id2idx;
	std::vector<Edge> edges;
	std::vector<std::set<u32>> e_l, e_r;           // per vertex: incident edge idxs (ascending)
	std::set<std::pair<u32, u32>> tips;            // (v_id, side), ordered like side_n_id_t (types.cpp:60-68)
	std::vector<Path> paths;
	std::vector<std::vector<std::vector<u32>>> v2steps; // [v_idx][ref] -> sorted step idxs

	u32 add_vertex(u32 id, const std::string &lab) {
		ids.push_back(id); labels.push_back(lab); e_l.emplace_back(); e_r.emplace_back();
		id2idx[id] = (u32)ids.size() - 1; return (u32)ids.size() - 1;
	}
	/* bidirected.cpp:321-340 */
	u32 add_edge(u32 v1_id, u32 s1, u32 v2_id, u32 s2) {
		u32 a = id2idx.at(v1_id), b = id2idx.at(v2_id);
		edges.push_back({a, s1, b, s2});
		u32 e = (u32)edges.size() - 1;
		(s1 == SIDE_L ? e_l[a] : e_r[a]).insert(e);
		(s2 == SIDE_L ? e_l[b] : e_r[b]).insert(e);
		return e;
	}
	/* Edge::get_other_vtx(v_idx, ve), bidirected.cpp:66-77: any self-loop answers
	 * (complement side, same vertex) */
	std::pair<u32, u32> other(u32 e_idx, u32 v_idx, u32 side) const {
		const Edge &e = edges[e_idx];
		if (e.v1 == e.v2) return {side ^ 1u, e.v1};
		return e.v1 == v_idx ? std::make_pair(e.s2, e.v2) : std::make_pair(e.s1, e.v1);
	}
	/* one-argument form, bidirected.cpp:60-64 */
	u32 other_v(u32 e_idx, u32 v_idx) const {
		const Edge &e = edges[e_idx];
		return e.v1 == v_idx ? e.v2 : e.v1;
	}
};

inline std::vector<std::string> split(const std::string &s, char d) {
	std::vector<std::string> out; size_t b = 0;
	for (;;) {
		size_t e = s.find(d, b);
		out.push_back(s.substr(b, e == std::string::npos ? e : e - b));
		if (e == std::string::npos) break;
		b = e + 1;
	}
	return out;
}

/* from_gfa.cpp:57-98 pre-scan + liteseq rules + from_gfa.cpp:141-280 */
inline Graph load_gfa(const std::string &fp, bool inc_labels, bool inc_refs) {
	auto bad = [&](const std::string &d) { return std::runtime_error("Invalid GFA '" + fp + "': " + d); };
	std::ifstream in(fp);
	if (!in) throw bad("could not open file");
	std::string line; size_t line_no = 1;
	std::vector<std::vector<std::string>> S, L, P;
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		if (line.empty()) { ++line_no; continue; }
		char t = line[0];
		if (t != 'H' && t != 'L' && t != 'P' && t != 'W' && t != 'S')
			throw bad("unsupported record type '" + std::string(1, t) + "' on line " + std::to_string(line_no));
		std::vector<std::string> f = split(line, '\t');
		if (t == 'S') {
			if (f.size() < 2) throw bad("S record on line " + std::to_string(line_no) + " is missing a segment id and sequence");
			if (f.size() < 3) throw bad("S record on line " + std::to_string(line_no) + " is missing a sequence");
			if (f[2].empty()) throw bad("S record on line " + std::to_string(line_no) + " has an empty sequence");
			S.push_back(f);
		} else if (t == 'L') L.push_back(f);
		else if (t == 'P') P.push_back(f);
		++line_no;
	}
	/* liteseq stand-in rules */
	for (auto &f : L) if (f.size() < 6 || (f[5] != "0M" && f[5] != "*")) throw bad("liteseq could not parse the file");
	for (auto &f : P) if (f.size() < 4) throw bad("liteseq could not parse the file");
	Graph g;
	std::map<u32, std::string> seg;               // ascending numeric id == liteseq's id-indexed table
	for (auto &f : S) { u32 id = (u32)strtoul(f[1].c_str(), 0, 10); if (!seg.count(id)) seg[id] = f[2]; }
	if (seg.empty()) throw bad("liteseq returned no vertices");
	for (auto &kv : seg) g.add_vertex(kv.first, inc_labels ? kv.second : std::string());
	for (auto &f : L) {
		u32 a = (u32)strtoul(f[1].c_str(), 0, 10), b = (u32)strtoul(f[3].c_str(), 0, 10);
		if (!seg.count(a) || !seg.count(b)) throw bad("liteseq could not parse the file");
		g.add_edge(a, f[2] == "+" ? SIDE_R : SIDE_L, b, f[4] == "+" ? SIDE_L : SIDE_R);
	}
	if (inc_refs) {
		g.v2steps.assign(g.ids.size(), std::vector<std::vector<u32>>(P.size()));
		for (size_t r = 0; r < P.size(); r++) {
			auto &f = P[r]; Path p; p.tag = f[1];
			std::vector<std::string> parts = split(f[1], '#');
			p.pansn = parts.size() >= 3 && !parts[1].empty() && parts[1].find_first_not_of("0123456789") == std::string::npos;
			if (p.pansn) { p.sample = parts[0]; p.hap_id = (u32)strtoul(parts[1].c_str(), 0, 10); }
			else { p.sample = f[1]; p.hap_id = 0; }
			u32 pos = 1;
			for (auto &s : split(f[2], ',')) {
				if (s.size() < 2 || (s.back() != '+' && s.back() != '-')) throw bad("liteseq could not parse the file");
				u32 id = (u32)strtoul(s.substr(0, s.size() - 1).c_str(), 0, 10);
				if (!seg.count(id)) throw bad("liteseq could not parse the file");
				p.v_ids.push_back(id); p.strands.push_back(s.back() == '+' ? 0 : 1); p.loci.push_back(pos);
				pos += (u32)seg[id].size();
			}
			p.hap_len = pos - 1;
			for (u32 k = 0; k < p.v_ids.size(); k++) {   /* bidirected.cpp:347-355 */
				std::vector<u32> &v = g.v2steps[g.id2idx[p.v_ids[k]]][r];
				auto it = std::lower_bound(v.begin(), v.end(), k);
				if (it == v.end() || *it != k) v.insert(it, k);
			}
			g.paths.push_back(std::move(p));
		}
	}
	/* tips, from_gfa.cpp:262-277 */
	for (u32 v = 0; v < g.ids.size(); v++) {
		if (g.e_l[v].empty() && g.e_r[v].empty()) g.tips.insert({g.ids[v], SIDE_L});
		else if (g.e_l[v].empty()) g.tips.insert({g.ids[v], SIDE_L});
		else if (g.e_r[v].empty()) g.tips.insert({g.ids[v], SIDE_R});
	}
	return g;
}

} // namespace oracle
