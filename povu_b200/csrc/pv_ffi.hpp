// povu_b200 -- implementation of include/povu_ffi.h on top of the session layer (pv_engine.hpp).
// Behaviour follows /root/reference/povu-rs/povu-ffi/povu_ffi.cpp:56-536 (handles, ownership, error
// reporting); the compute entry points run the sm_100a pipelines instead of the reference's stubs.
#pragma once
#include "../../include/povu_ffi.h"
#include "pv_engine.hpp"

struct PovuGraph {
	pv::Ctx ctx; bool ctx_ready = false;
	pve::Session *s = nullptr;
	// builder state (povu_graph_new / add_*): vertex idx = insertion order, like bd::VG::add_vertex
	std::vector<uint64_t> b_ids; std::vector<std::string> b_seqs;
	std::unordered_map<uint64_t, size_t> b_index;  // id -> builder vertex idx (add_edge resolves ids in O(1); duplicates are rejected)
	std::vector<uint32_t> b_ep1, b_ep2;            // by builder vertex idx
	struct BPath { std::string name; std::vector<PovuStep> steps; };
	std::vector<BPath> b_paths;
	bool from_gfa = false, synced = false;
	std::vector<std::string> prefixes;
	~PovuGraph() { delete s; if (ctx_ready) ctx.destroy(); }
	void need_device() {
		if (ctx_ready) return;
		int dev = 0;
		if (const char *e = getenv("POVU_B200_DEVICE")) dev = atoi(e);
#ifndef PV_EMU
		int n = 0;
		if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) throw std::runtime_error("no CUDA device available (povu_b200 has no CPU fallback)");
		if (dev < 0 || dev >= n) throw std::runtime_error("POVU_B200_DEVICE is not a valid CUDA ordinal");
#endif
		ctx.init(dev, nullptr); ctx_ready = true;
	}
	// materialise the builder arrays as a HostGraph.  The device encoding wants vertex idx order == ascending id order
	// (what mto::from_gfa::to_bd produces), so vertices added in any order are sorted by id here and the builder's
	// edge ports are remapped; the getters keep answering in insertion order.
	void sync() {
		if (from_gfa || synced) return;
		const uint32_t V = (uint32_t)b_ids.size();
		std::vector<uint32_t> ord(V), rank(V);
		for (uint32_t i = 0; i < V; i++) ord[i] = i;
		bool sorted = true;
		for (size_t i = 1; i < b_ids.size(); i++) if (b_ids[i - 1] >= b_ids[i]) { sorted = false; break; }
		if (!sorted) std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return b_ids[a] < b_ids[b]; });
		for (uint32_t i = 0; i < V; i++) rank[ord[i]] = i;
		pvh::HostGraph &g = s->g;
		g = pvh::HostGraph(); g.has_labels = true;
		g.ids.resize(V); g.segLen.resize(V); g.lblOff.assign((size_t)V + 1, 0);
		for (uint32_t i = 0; i < V; i++) {
			const uint32_t b = ord[i];
			if (b_ids[b] > 0xFFFFFFFFull) throw std::runtime_error("povu_b200: vertex id does not fit 32 bits");
			g.ids[i] = (uint32_t)b_ids[b]; g.segLen[i] = (uint32_t)b_seqs[b].size(); g.lblOff[i + 1] = g.lblOff[i] + b_seqs[b].size(); g.lbl += b_seqs[b];
		}
		g.build_id_table();
		g.ep1.resize(b_ep1.size()); g.ep2.resize(b_ep2.size());
		for (size_t e = 0; e < b_ep1.size(); e++) { g.ep1[e] = (rank[b_ep1[e] >> 1] << 1) | (b_ep1[e] & 1u); g.ep2[e] = (rank[b_ep2[e] >> 1] << 1) | (b_ep2[e] & 1u); }
		g.stepOff.assign(1, 0);
		for (auto &p : b_paths) {
			pvh::HostPath hp = pvh::parse_path_name(p.name);
			uint64_t len = 0;
			for (auto &st : p.steps) { uint32_t v = g.idx_of(st.vertex_id); if (v == pvh::NONE32) throw std::runtime_error("povu_b200: path step names an unknown vertex"); g.steps.push_back((v << 1) | (st.orientation == POVU_ORIENTATION_REVERSE ? 1u : 0u)); len += g.segLen[v]; }
			hp.hap_len = (uint32_t)len;
			g.paths.push_back(std::move(hp)); g.stepOff.push_back((uint32_t)g.steps.size());
		}
		// the graph changed: nothing derived from the previous one may survive (device arrays, text tables, forest, labels, records)
		s->host_edges = s->host_steps = s->host_lbl = true;
		s->drop_device(); s->invalidate_graph_state();
		synced = true;
	}
};
struct PovuFlubbles { PovuGraph *graph; pvh::Forest forest; };
struct PovuPvstTree { const PovuFlubbles *f; };
struct PovuVcfOutput { std::string vcf_content; };

static void ffi_set_error(PovuError *error, int code, const char *message) {
	if (!error) return;
	const char *m = message ? message : "Unknown error";
	size_t n = strlen(m);
	error->code = code; error->message = new char[n + 1];
	memcpy(error->message, m, n + 1);
}
static std::vector<std::string> ffi_read_lines(const char *fp) {
	std::ifstream in(fp);
	if (!in) throw std::runtime_error(std::string("could not open ") + fp);
	std::vector<std::string> out; std::string l;
	while (std::getline(in, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); if (!l.empty()) out.push_back(l); }
	return out;
}

extern "C" {

PovuGraph *povu_graph_new(size_t vc, size_t ec, size_t pc) {
	try {
		PovuGraph *g = new PovuGraph();
		g->s = new pve::Session(g->ctx);
		g->b_ids.reserve(vc); g->b_seqs.reserve(vc); g->b_ep1.reserve(ec); g->b_ep2.reserve(ec); g->b_paths.reserve(pc);
		return g;
	} catch (...) { return nullptr; }
}
PovuGraph *povu_graph_from_gfa(const char *gfa_path, PovuError *error) {
	PovuGraph *g = nullptr;
	try {
		if (!gfa_path || gfa_path[0] == '\0') { ffi_set_error(error, 1, "GFA path must not be empty"); return nullptr; }
		struct stat st;
		if (stat(gfa_path, &st) != 0 || !S_ISREG(st.st_mode)) { ffi_set_error(error, 1, (std::string("GFA file does not exist: ") + gfa_path).c_str()); return nullptr; }
		g = new PovuGraph();
		g->s = new pve::Session(g->ctx);
		g->s->g = pvh::load_gfa(gfa_path, true, true);
		g->from_gfa = true;
		return g;
	} catch (std::exception &e) { ffi_set_error(error, 1, e.what()); delete g; return nullptr; }
}
void povu_graph_free(PovuGraph *graph) { delete graph; }

size_t povu_graph_add_vertex(PovuGraph *graph, uint64_t id, const char *sequence) {
	if (!graph || !sequence || graph->from_gfa) return (size_t)-1;
	try {
		if (!graph->b_index.emplace(id, graph->b_ids.size()).second) return (size_t)-1;   // duplicate id
		graph->b_ids.push_back(id); graph->b_seqs.emplace_back(sequence); graph->synced = false; return graph->b_ids.size() - 1;
	} catch (...) { return (size_t)-1; }
}
size_t povu_graph_add_edge(PovuGraph *graph, uint64_t from_id, PovuOrientation fo, uint64_t to_id, PovuOrientation to) {
	if (!graph || graph->from_gfa) return (size_t)-1;
	try {
		auto find = [&](uint64_t id) -> size_t { auto it = graph->b_index.find(id); return it == graph->b_index.end() ? (size_t)-1 : it->second; };
		size_t a = find(from_id), b = find(to_id);
		if (a == (size_t)-1 || b == (size_t)-1) return (size_t)-1;
		graph->b_ep1.push_back((uint32_t)(2 * a + (fo == POVU_ORIENTATION_FORWARD ? 0u : 1u)));
		graph->b_ep2.push_back((uint32_t)(2 * b + (to == POVU_ORIENTATION_FORWARD ? 0u : 1u)));
		graph->synced = false;
		return graph->b_ep1.size() - 1;
	} catch (...) { return (size_t)-1; }
}
bool povu_graph_add_path(PovuGraph *graph, const char *name, const PovuStep *steps, size_t n) {
	if (!graph || !name || !steps || graph->from_gfa) return false;
	try { PovuGraph::BPath p; p.name = name; p.steps.assign(steps, steps + n); graph->b_paths.push_back(std::move(p)); graph->synced = false; return true; } catch (...) { return false; }
}
void povu_graph_finalize(PovuGraph *graph) { if (!graph) return; try { graph->sync(); } catch (...) {} }
size_t povu_graph_vertex_count(const PovuGraph *g) { return g ? (g->from_gfa ? g->s->g.V() : g->b_ids.size()) : 0; }
size_t povu_graph_edge_count(const PovuGraph *g) { return g ? (g->from_gfa ? (size_t)g->s->g.E() : g->b_ep1.size()) : 0; }
size_t povu_graph_path_count(const PovuGraph *g) { return g ? (g->from_gfa ? g->s->g.H() : g->b_paths.size()) : 0; }

PovuVertex *povu_graph_get_vertices(const PovuGraph *graph, size_t *count) {
	if (!graph || !count) { if (count) *count = 0; return nullptr; }
	const size_t n = povu_graph_vertex_count(graph);
	*count = n;
	if (graph->from_gfa) { try { graph->s->ensure_host(false, false, true); } catch (...) { *count = 0; return nullptr; } }
	PovuVertex *out = new PovuVertex[n];
	for (size_t i = 0; i < n; i++) {
		std::string lab = graph->from_gfa ? std::string(graph->s->g.label((uint32_t)i), graph->s->g.segLen[i]) : graph->b_seqs[i];
		char *seq = new char[lab.size() + 1]; memcpy(seq, lab.c_str(), lab.size() + 1);
		out[i].id = graph->from_gfa ? graph->s->g.ids[i] : graph->b_ids[i]; out[i].sequence = seq; out[i].sequence_len = lab.size();
	}
	return out;
}
PovuEdge *povu_graph_get_edges(const PovuGraph *graph, size_t *count) {
	if (!graph || !count) { if (count) *count = 0; return nullptr; }
	const size_t n = povu_graph_edge_count(graph);
	*count = n;
	if (graph->from_gfa) { try { graph->s->ensure_host(true, false, false); } catch (...) { *count = 0; return nullptr; } }
	PovuEdge *out = new PovuEdge[n];
	for (size_t i = 0; i < n; i++) {
		uint32_t a = graph->from_gfa ? graph->s->g.ep1[i] : graph->b_ep1[i], b = graph->from_gfa ? graph->s->g.ep2[i] : graph->b_ep2[i];
		out[i].from_id = graph->from_gfa ? graph->s->g.ids[a >> 1] : graph->b_ids[a >> 1]; out[i].from_orientation = (PovuOrientation)(a & 1u);
		out[i].to_id = graph->from_gfa ? graph->s->g.ids[b >> 1] : graph->b_ids[b >> 1]; out[i].to_orientation = (PovuOrientation)(b & 1u);
	}
	return out;
}
PovuPath *povu_graph_get_paths(const PovuGraph *graph, size_t *count) {
	if (!graph || !count) { if (count) *count = 0; return nullptr; }
	const size_t n = povu_graph_path_count(graph);
	*count = n;
	if (n == 0) return nullptr;
	if (graph->from_gfa) { try { graph->s->ensure_host(false, true, false); } catch (...) { *count = 0; return nullptr; } }
	PovuPath *out = new PovuPath[n];
	for (size_t r = 0; r < n; r++) {
		const std::string &nm = graph->from_gfa ? graph->s->g.paths[r].tag : graph->b_paths[r].name;
		char *name = new char[nm.size() + 1]; memcpy(name, nm.c_str(), nm.size() + 1);
		out[r].name = name; out[r].name_len = nm.size();
		size_t sc = graph->from_gfa ? graph->s->g.stepOff[r + 1] - graph->s->g.stepOff[r] : graph->b_paths[r].steps.size();
		out[r].steps_count = sc; out[r].steps = sc ? new PovuStep[sc] : nullptr;
		for (size_t j = 0; j < sc; j++) {
			if (graph->from_gfa) { uint32_t st = graph->s->g.steps[graph->s->g.stepOff[r] + j]; out[r].steps[j].vertex_id = graph->s->g.ids[st >> 1]; out[r].steps[j].orientation = (PovuOrientation)(st & 1u); }
			else out[r].steps[j] = graph->b_paths[r].steps[j];
		}
	}
	return out;
}
void povu_vertices_free(PovuVertex *v, size_t count) { if (!v) return; for (size_t i = 0; i < count; i++) delete[] v[i].sequence; delete[] v; }
void povu_edges_free(PovuEdge *e, size_t) { delete[] e; }
void povu_paths_free(PovuPath *p, size_t count) { if (!p) return; for (size_t i = 0; i < count; i++) { delete[] p[i].name; delete[] p[i].steps; } delete[] p; }

bool povu_graph_set_references_from_file(PovuGraph *graph, const char *ref_file, PovuError *error) {
	if (!graph || !ref_file) { ffi_set_error(error, 1, "Invalid arguments"); return false; }
	try { graph->prefixes = ffi_read_lines(ref_file); return true; } catch (std::exception &e) { ffi_set_error(error, 1, e.what()); return false; }
}
bool povu_graph_set_references_from_prefixes(PovuGraph *graph, const char **prefixes, size_t count, PovuError *error) {
	if (!graph || !prefixes) { ffi_set_error(error, 1, "Invalid arguments"); return false; }
	try { graph->prefixes.clear(); for (size_t i = 0; i < count; i++) graph->prefixes.push_back(prefixes[i]); return true; } catch (std::exception &e) { ffi_set_error(error, 1, e.what()); return false; }
}

PovuFlubbles *povu_graph_find_flubbles(PovuGraph *graph, PovuError *error) {
	if (!graph) { ffi_set_error(error, 1, "Invalid graph"); return nullptr; }
	try {
		graph->sync(); graph->need_device();
		graph->s->decompose();
		PovuFlubbles *f = new PovuFlubbles{graph, graph->s->forest};
		return f;
	} catch (std::exception &e) { ffi_set_error(error, 1, e.what()); return nullptr; }
}
void povu_flubbles_free(PovuFlubbles *f) { delete f; }
size_t povu_flubbles_count(const PovuFlubbles *f) { return f ? f->forest.n() : 0; }
PovuFlubble *povu_flubbles_get(const PovuFlubbles *f, size_t index) {
	if (!f || index >= f->forest.n()) return nullptr;
	PovuFlubble *o = new PovuFlubble();
	o->id = index; o->type_name = "flubble";
	o->start_vertex_id = f->graph->s->g.ids[f->forest.a_v[index]]; o->end_vertex_id = f->graph->s->g.ids[f->forest.z_v[index]];
	o->walks = nullptr; o->walk_lengths = nullptr; o->walks_count = 0;
	return o;
}
void povu_flubble_free(PovuFlubble *fl) {
	if (!fl) return;
	if (fl->walks) { for (size_t i = 0; i < fl->walks_count; i++) delete[] fl->walks[i]; delete[] fl->walks; }
	delete[] fl->walk_lengths;
	delete fl;
}
PovuPvstTree *povu_flubbles_get_pvst_tree(const PovuFlubbles *f) { return f ? new PovuPvstTree{f} : nullptr; }
void povu_pvst_tree_free(PovuPvstTree *t) { delete t; }
size_t povu_pvst_tree_vertex_count(const PovuPvstTree *t) { return t ? t->f->forest.n() + t->f->forest.comp_id.size() : 0; }

PovuVcfOutput *povu_flubbles_call_variants(PovuFlubbles *f, PovuError *error) {
	if (!f || !f->graph) { ffi_set_error(error, 1, "Invalid flubbles or graph"); return nullptr; }
	try {
		PovuGraph *g = f->graph;
		if (g->prefixes.empty()) throw std::runtime_error("no reference prefixes set (povu_graph_set_references_from_prefixes / _from_file)");
		g->need_device();
		pve::Session &s = *g->s;
		s.forest = f->forest; s.forest_ready = true;
		s.set_prefixes(g->prefixes);
		s.call();
		if (s.rec.n_hard_errors) throw std::runtime_error("call: the reference aborts on this input (missing context boundary in a haplotype lap)");
		PovuVcfOutput *o = new PovuVcfOutput();
		pv::TextOut body;
		s.vcf_records_device(100, nullptr, body);
		o->vcf_content = s.vcf_header(s.prefixes, true);
		if (!body.text.empty()) o->vcf_content.append(body.text.data(), body.text.size());
		return o;
	} catch (std::exception &e) { ffi_set_error(error, 1, e.what()); return nullptr; }
}
bool povu_vcf_write_to_file(const PovuVcfOutput *vcf, const char *path, PovuError *error) {
	if (!vcf || !path) { ffi_set_error(error, 1, "Invalid arguments"); return false; }
	try {
		std::ofstream out(path, std::ios::binary);
		if (!out) { ffi_set_error(error, 1, "Failed to open output file"); return false; }
		out.write(vcf->vcf_content.data(), (std::streamsize)vcf->vcf_content.size());
		return true;
	} catch (std::exception &e) { ffi_set_error(error, 1, e.what()); return false; }
}
char *povu_vcf_to_string(const PovuVcfOutput *vcf, size_t *length) {
	if (!vcf || !length) { if (length) *length = 0; return nullptr; }
	*length = vcf->vcf_content.size();
	char *r = new char[*length + 1];
	memcpy(r, vcf->vcf_content.c_str(), *length + 1);
	return r;
}
void povu_vcf_free(PovuVcfOutput *vcf) { delete vcf; }
void povu_string_free(char *str) { delete[] str; }

bool povu_gfa_to_vcf(const char *gfa_path, const char *vcf_path, const char *ref_file, PovuError *error) {
	if (!gfa_path || !vcf_path) { ffi_set_error(error, 1, "Invalid arguments"); return false; }
	PovuGraph *g = povu_graph_from_gfa(gfa_path, error);
	if (!g) return false;
	bool ok = false;
	PovuFlubbles *f = nullptr; PovuVcfOutput *o = nullptr;
	do {
		if (!ref_file) { ffi_set_error(error, 1, "a reference prefix file is required"); break; }
		if (!povu_graph_set_references_from_file(g, ref_file, error)) break;
		f = povu_graph_find_flubbles(g, error); if (!f) break;
		o = povu_flubbles_call_variants(f, error); if (!o) break;
		ok = povu_vcf_write_to_file(o, vcf_path, error);
	} while (0);
	povu_vcf_free(o); povu_flubbles_free(f); povu_graph_free(g);
	return ok;
}
void povu_error_free(PovuError *error) { if (error && error->message) { delete[] error->message; error->message = nullptr; error->code = 0; } }

}
