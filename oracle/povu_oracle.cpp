/* TEST INFRASTRUCTURE ONLY (oracle) -- CLI + C entry points over the CPU restatement in
 * oracle_*.hpp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may run this.  Mirrors the reference driver contract
 * (app/subcommand/decompose.cpp:94-160: one <component_id>.pvst per component >= 3 vertices). */
#include <cstring>
#include <iostream>
#include "oracle_call.hpp"

using namespace oracle;

struct OracleResult {
	std::vector<u32> comp, a_id, z_id, parent; std::vector<uint8_t> a_fwd, z_fwd;
	std::vector<u32> comp_ids, comp_nv, comp_nfl;
};

static OracleResult *pack(const std::vector<DecomposedComponent> &d) {
	OracleResult *r = new OracleResult();
	for (auto &c : d) {
		r->comp_ids.push_back(c.component_id); r->comp_nv.push_back(c.n_vertices); r->comp_nfl.push_back((u32)c.flubbles.size());
		for (auto &f : c.flubbles) {
			r->comp.push_back(c.component_id); r->a_id.push_back(f.a_id); r->z_id.push_back(f.z_id);
			r->a_fwd.push_back(f.a_fwd); r->z_fwd.push_back(f.z_fwd); r->parent.push_back(f.parent);
		}
	}
	return r;
}

extern "C" {
/* Graph given as arrays: ids ascending; edges as (v1_idx, side1, v2_idx, side2) in L-line order;
 * side 0 = left, 1 = right.  Tips are inferred as in from_gfa.cpp:262-277. */
OracleResult *oracle_decompose_arrays(uint32_t V, const uint32_t *ids, uint64_t E, const uint32_t *v1, const uint8_t *s1,
				      const uint32_t *v2, const uint8_t *s2) {
	try {
		Graph g;
		for (u32 i = 0; i < V; i++) g.add_vertex(ids[i], "");
		for (uint64_t e = 0; e < E; e++) g.add_edge(ids[v1[e]], s1[e], ids[v2[e]], s2[e]);
		for (u32 v = 0; v < V; v++) {
			if (g.e_l[v].empty() && g.e_r[v].empty()) g.tips.insert({g.ids[v], SIDE_L});
			else if (g.e_l[v].empty()) g.tips.insert({g.ids[v], SIDE_L});
			else if (g.e_r[v].empty()) g.tips.insert({g.ids[v], SIDE_R});
		}
		return pack(decompose(g));
	} catch (std::exception &e) { fprintf(stderr, "oracle error: %s\n", e.what()); return nullptr; }
}
OracleResult *oracle_decompose_gfa(const char *fp) {
	try { return pack(decompose(load_gfa(fp, false, false))); }
	catch (std::exception &e) { fprintf(stderr, "oracle error: %s\n", e.what()); return nullptr; }
}
uint64_t oracle_result_flubble_count(const OracleResult *r) { return r->a_id.size(); }
uint64_t oracle_result_component_count(const OracleResult *r) { return r->comp_ids.size(); }
void oracle_result_copy(const OracleResult *r, uint32_t *comp, uint32_t *a_id, uint8_t *a_fwd, uint32_t *z_id, uint8_t *z_fwd, uint32_t *parent) {
	size_t n = r->a_id.size();
	memcpy(comp, r->comp.data(), 4 * n); memcpy(a_id, r->a_id.data(), 4 * n); memcpy(z_id, r->z_id.data(), 4 * n);
	memcpy(parent, r->parent.data(), 4 * n); memcpy(a_fwd, r->a_fwd.data(), n); memcpy(z_fwd, r->z_fwd.data(), n);
}
void oracle_result_copy_components(const OracleResult *r, uint32_t *ids, uint32_t *nv, uint32_t *nfl) {
	size_t n = r->comp_ids.size();
	memcpy(ids, r->comp_ids.data(), 4 * n); memcpy(nv, r->comp_nv.data(), 4 * n); memcpy(nfl, r->comp_nfl.data(), 4 * n);
}
void oracle_result_free(OracleResult *r) { delete r; }

/* class stack of the k-th decomposed component of a GFA: entries (tree_edge_id, class, v_id, forward);
 * tree_edge_id is the reference's creation-order edge id (spanning_tree.cpp:784-803) that the
 * conformance cycle oracle (tests/lean4_conformance/src/main.rs:1594-1622) is keyed on.
 * Returns the number of entries written (<= cap), or -1 on error. */
int64_t oracle_class_stack_gfa(const char *fp, uint32_t k, uint64_t cap, uint32_t *edge_id, uint32_t *cls, uint32_t *v_id, uint8_t *fwd) {
	try {
		Graph g = load_gfa(fp, false, false);
		std::vector<Component> comps = componetize(g);
		uint32_t seen = 0;
		for (auto &c : comps) {
			if (c.g.ids.size() < 3) continue;
			if (seen++ != k) continue;
			STree t = spanning_tree(c.g);
			cycle_equiv(t);
			std::vector<StackEntry> s = class_stack(t);
			uint64_t n = std::min<uint64_t>(cap, s.size());
			for (uint64_t i = 0; i < n; i++) { edge_id[i] = t.tedges[s[i].tree_edge].id; cls[i] = s[i].cls; v_id[i] = s[i].v_id; fwd[i] = s[i].forward; }
			return (int64_t)n;
		}
		return 0;
	} catch (std::exception &e) { fprintf(stderr, "oracle error: %s\n", e.what()); return -1; }
}
/* decompose without tip inference (a VG assembled in memory, as in tests/integration_tests/pvst_tests.cc:12-39) */
OracleResult *oracle_decompose_arrays_notips(uint32_t V, const uint32_t *ids, uint64_t E, const uint32_t *v1, const uint8_t *s1,
					     const uint32_t *v2, const uint8_t *s2) {
	try {
		Graph g;
		for (u32 i = 0; i < V; i++) g.add_vertex(ids[i], "");
		for (uint64_t e = 0; e < E; e++) g.add_edge(ids[v1[e]], s1[e], ids[v2[e]], s2[e]);
		return pack(decompose(g));
	} catch (std::exception &e) { fprintf(stderr, "oracle error: %s\n", e.what()); return nullptr; }
}
}

#ifndef ORACLE_NO_MAIN
int main(int argc, char **argv) {
	if (argc < 2) { fprintf(stderr, "usage: povu_oracle decompose -i gfa -o dir | gfa2vcf -i gfa -P prefix [-c chunk]\n"); return 2; }
	std::string task = argv[1], gfa, out = ".";
	std::vector<std::string> prefixes; size_t chunk = 100;
	for (int i = 2; i < argc; i++) {
		std::string a = argv[i];
		if (a == "-i" && i + 1 < argc) gfa = argv[++i];
		else if (a == "-o" && i + 1 < argc) out = argv[++i];
		else if (a == "-P" && i + 1 < argc) prefixes.push_back(argv[++i]);
		else if (a == "-c" && i + 1 < argc) chunk = (size_t)atol(argv[++i]);
		else if (a == "-t" && i + 1 < argc) ++i;
	}
	try {
		if (task == "decompose") {
			Graph g = load_gfa(gfa, false, false);
			for (auto &c : decompose(g)) {
				std::ofstream f(out + "/" + std::to_string(c.component_id) + ".pvst");
				f << pvst_text(c.flubbles);
			}
		} else if (task == "gfa2vcf") {
			Graph g = load_gfa(gfa, true, true);
			std::vector<Pvst> pvsts;
			for (auto &c : decompose(g)) { Pvst p; p.component_id = c.component_id; p.fl = c.flubbles; p.finish(); pvsts.push_back(std::move(p)); }
			CallStats st;
			std::string vcf = call_vcf(g, pvsts, prefixes, chunk, true, st);
			fwrite(vcf.data(), 1, vcf.size(), stdout);
			fprintf(stderr, "[oracle] rovs=%llu tangled_skipped=%llu pin_pairs=%llu too_small=%llu hard_errors=%llu\n", (unsigned long long)st.rovs,
				(unsigned long long)st.tangled_skipped, (unsigned long long)st.pin_pairs, (unsigned long long)st.too_small, (unsigned long long)st.hard_errors);
		} else { fprintf(stderr, "unknown task\n"); return 2; }
	} catch (std::exception &e) { std::cerr << "[povu::subcommands::do_" << task << "] Error: " << e.what() << std::endl; return 1; }
	return 0;
}
#endif
