/* TEST INFRASTRUCTURE ONLY (oracle). ~100-line stand-in for liteseq's GFA 1.0
 * parser so that the UNMODIFIED reference sources under /root/reference can be
 * compiled into oracle/_ref/.  Rules: vertex table is
 * indexed by numeric segment id (size max_id+1); `L a o1 b o2 ovl`: a's side is
 * RIGHT if o1=='+' else LEFT, b's side is LEFT if o2=='+' else RIGHT; ovl must
 * be 0M or * (else status=1); `P name steps ovl` needs >=4 columns (else
 * status=1); loci[k] = 1-based start of step k; PanSN `sample#hap#contig`. */
#include <liteseq/gfa.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
namespace liteseq {
static std::vector<std::string> split(const std::string &s, char d) {
	std::vector<std::string> out; size_t b = 0;
	for (;;) { size_t e = s.find(d, b); out.push_back(s.substr(b, e == std::string::npos ? e : e - b)); if (e == std::string::npos) break; b = e + 1; }
	return out;
}
static char *dup(const std::string &s) { char *p = (char *)malloc(s.size() + 1); memcpy(p, s.c_str(), s.size() + 1); return p; }
gfa_props *gfa_new(gfa_config *conf) {
	gfa_props *g = (gfa_props *)calloc(1, sizeof(gfa_props));
	std::ifstream in(conf->fp);
	if (!in) { g->status = 1; return g; }
	std::string line; std::vector<std::vector<std::string>> S, L, P; id_t max_id = 0; bool any = false;
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		if (line.empty()) continue;
		char t = line[0];
		if (t != 'S' && t != 'L' && t != 'P') continue;
		std::vector<std::string> f = split(line, '\t');
		if (t == 'S') { if (f.size() < 3) { g->status = 1; return g; } id_t id = (id_t)strtoul(f[1].c_str(), 0, 10); if (id > max_id) max_id = id; any = true; S.push_back(f); }
		else if (t == 'L') { if (f.size() < 6 || (f[5] != "0M" && f[5] != "*")) { g->status = 1; return g; } L.push_back(f); }
		else { if (f.size() < 4) { g->status = 1; return g; } P.push_back(f); }
	}
	g->vtx_arr_size = any ? max_id + 1 : 0;
	g->v = (vtx **)calloc(g->vtx_arr_size ? g->vtx_arr_size : 1, sizeof(vtx *));
	for (auto &f : S) { id_t id = (id_t)strtoul(f[1].c_str(), 0, 10); if (g->v[id]) continue; vtx *x = (vtx *)calloc(1, sizeof(vtx)); x->id = id; x->seq = conf->inc_vtx_labels ? dup(f[2]) : nullptr; g->v[id] = x; }
	g->l_line_count = (idx_t)L.size();
	g->e = (edge *)calloc(L.size() ? L.size() : 1, sizeof(edge));
	for (size_t i = 0; i < L.size(); i++) {
		auto &f = L[i]; edge &e = g->e[i];
		e.v1_id = (id_t)strtoul(f[1].c_str(), 0, 10); e.v1_side = f[2] == "+" ? RIGHT : LEFT;
		e.v2_id = (id_t)strtoul(f[3].c_str(), 0, 10); e.v2_side = f[4] == "+" ? LEFT : RIGHT;
		if (e.v1_id >= g->vtx_arr_size || e.v2_id >= g->vtx_arr_size || !g->v[e.v1_id] || !g->v[e.v2_id]) { g->status = 1; return g; }
	}
	if (!conf->inc_refs) return g;
	g->ref_count = (idx_t)P.size();
	g->refs = (ref **)calloc(P.size() ? P.size() : 1, sizeof(ref *));
	for (size_t i = 0; i < P.size(); i++) {
		auto &f = P[i]; ref *r = (ref *)calloc(1, sizeof(ref)); g->refs[i] = r;
		r->tag = dup(f[1]);
		std::vector<std::string> parts = split(f[1], '#');
		bool pansn = parts.size() >= 3 && !parts[1].empty() && parts[1].find_first_not_of("0123456789") == std::string::npos;
		if (pansn) { r->id_type = REF_ID_PANSN; r->sample_name = dup(parts[0]); r->hap_id = (idx_t)strtoul(parts[1].c_str(), 0, 10); }
		else { r->id_type = REF_ID_RAW; r->sample_name = dup(f[1]); r->hap_id = 0; }
		std::vector<std::string> steps = split(f[2], ',');
		ref_walk *w = (ref_walk *)calloc(1, sizeof(ref_walk)); r->walk = w;
		w->step_count = (idx_t)steps.size();
		w->v_ids = (id_t *)calloc(steps.size(), sizeof(id_t)); w->strands = (strand *)calloc(steps.size(), sizeof(strand)); w->loci = (idx_t *)calloc(steps.size(), sizeof(idx_t));
		for (size_t k = 0; k < steps.size(); k++) {
			const std::string &s = steps[k];
			if (s.size() < 2 || (s.back() != '+' && s.back() != '-')) { g->status = 1; return g; }
			id_t id = (id_t)strtoul(s.substr(0, s.size() - 1).c_str(), 0, 10);
			if (id >= g->vtx_arr_size || !g->v[id]) { g->status = 1; return g; }
			w->v_ids[k] = id; w->strands[k] = s.back() == '+' ? STRAND_FWD : STRAND_REV;
		}
	}
	/* second pass for loci: build id -> length table */
	std::vector<idx_t> len(g->vtx_arr_size, 0);
	for (auto &f : S) { id_t id = (id_t)strtoul(f[1].c_str(), 0, 10); if (!len[id]) len[id] = (idx_t)f[2].size(); }
	for (idx_t i = 0; i < g->ref_count; i++) {
		ref_walk *w = g->refs[i]->walk; idx_t pos = 1;
		for (idx_t k = 0; k < w->step_count; k++) { w->loci[k] = pos; pos += len[w->v_ids[k]]; }
		g->refs[i]->hap_len = pos - 1;
	}
	return g;
}
void gfa_free(gfa_props *g) {
	if (!g) return;
	if (g->v) { for (idx_t i = 0; i < g->vtx_arr_size; i++) if (g->v[i]) { free(g->v[i]->seq); free(g->v[i]); } free(g->v); }
	free(g->e);
	if (g->refs) { for (idx_t i = 0; i < g->ref_count; i++) { ref *r = g->refs[i]; if (!r) continue; if (r->walk) { free(r->walk->v_ids); free(r->walk->strands); free(r->walk->loci); free(r->walk); } free(r->tag); free(r->sample_name); free(r); } free(g->refs); }
	free(g);
}
vtx *get_vtx(gfa_props *g, idx_t i) { return i < g->vtx_arr_size ? g->v[i] : nullptr; }
ref *get_ref(gfa_props *g, idx_t i) { return i < g->ref_count ? g->refs[i] : nullptr; }
const char *get_tag(const ref *r) { return r->tag; }
const char *get_sample_name(const ref *r) { return r->sample_name; }
idx_t get_hap_id(const ref *r) { return r->hap_id; }
idx_t get_hap_len(const ref *r) { return r->hap_len; }
idx_t get_step_count(const ref *r) { return r->walk->step_count; }
ref_id_type get_ref_id_type(const ref *r) { return r->id_type; }
const id_t *get_walk_v_ids(const ref *r) { return r->walk->v_ids; }
const strand *get_walk_strands(const ref *r) { return r->walk->strands; }
}
