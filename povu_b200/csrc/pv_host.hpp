// povu_b200 -- host side of the decompose -> call path: GFA -> flat (pinned-able) arrays, path / genotype
// metadata, the .pvst v0.0.3 wire format and VCF 4.2 text from the device's binary records.
// Mirrors (same behaviour, own design):
//   mto::from_gfa::to_bd          /root/reference/src/mto/from_gfa.cpp:28-113,141-280   (liteseq rules: SURVEY.md Appendix A)
//   pr::Refs                      /root/reference/include/povu/refs/refs.hpp:278-301,359-412
//   mto::from_pvst::read_pvst     /root/reference/src/mto/from_pvst.cpp:82-302, comp_heights pvst.hpp:807-836
//   mto::to_vcf                   /root/reference/src/mto/to_vcf.cpp:19-139, VcfRec text vcf.hpp:295-383
//   hap_slice::{as_str,as_dna_str} /root/reference/include/ita/genomics/allele.hpp:134-209
// No graph algorithm lives here: everything between "arrays in" and "records out" runs on the device.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>
#include "pv_common.cuh"

namespace pvh {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
static const u32 NONE32 = 0xFFFFFFFFu;

struct HostPath { std::string tag, sample; u32 hap_id = 0; bool pansn = false; u32 hap_len = 0; };

struct HostGraph {
	std::vector<u32> ids;                 // vertex idx -> segment id, ascending
	std::vector<u64> lblOff; std::string lbl;   // concatenated sequences (when labels were requested)
	std::vector<u32> segLen;
	std::vector<u32> ep1, ep2;            // edge endpoint ports in L-line order: 2*vertex + side (0 = left, 1 = right)
	std::vector<HostPath> paths;
	std::vector<u32> stepOff;             // H + 1
	std::vector<u32> steps;               // (vertex idx << 1) | strand
	std::vector<u32> id_table;            // dense id -> idx (NONE32 = absent) when ids are dense enough
	bool has_labels = false;
	u64 devE = 0;                         // link count when the endpoint arrays live only in HBM (device-parsed GFA)

	u32 V() const { return (u32)ids.size(); }
	u64 E() const { return ep1.empty() ? devE : ep1.size(); }
	u32 H() const { return (u32)paths.size(); }
	u32 idx_of(u64 id) const {
		if (!id_table.empty()) return id < id_table.size() ? id_table[id] : NONE32;
		if (id > 0xFFFFFFFFull) return NONE32;
		auto it = std::lower_bound(ids.begin(), ids.end(), (u32)id);
		return (it != ids.end() && *it == id) ? (u32)(it - ids.begin()) : NONE32;
	}
	const char *label(u32 v) const { return lbl.data() + lblOff[v]; }
	void build_id_table() {
		id_table.clear();
		if (ids.empty()) return;
		u64 mx = ids.back();
		if (mx <= 4ull * ids.size() + 1024) { id_table.assign(mx + 1, NONE32); for (u32 i = 0; i < ids.size(); i++) id_table[ids[i]] = i; }
	}
};

// ---------------------------------------------------------------------------------------------
// GFA
// ---------------------------------------------------------------------------------------------
struct MappedFile {
	const char *p = nullptr; size_t n = 0; int fd = -1;
	explicit MappedFile(const std::string &fp) {
		fd = open(fp.c_str(), O_RDONLY);
		if (fd < 0) return;
		struct stat st;
		if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); fd = -1; return; }
		n = (size_t)st.st_size;
		if (n) { void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0); if (m == MAP_FAILED) { close(fd); fd = -1; n = 0; return; } p = (const char *)m; madvise(m, n, MADV_SEQUENTIAL); }
	}
	~MappedFile() { if (p) munmap((void *)p, n); if (fd >= 0) close(fd); }
	bool ok() const { return fd >= 0; }
};

inline bool parse_u64(const char *b, const char *e, u64 &out) {
	if (b == e) return false;
	u64 v = 0;
	for (const char *q = b; q < e; q++) { if (*q < '0' || *q > '9') return false; v = v * 10 + (u64)(*q - '0'); }
	out = v; return true;
}

// PanSN `sample#hap#contig` with a numeric hap -> (sample, hap id); anything else is a raw name whose sample
// name is the whole tag (liteseq ref_id_type)
inline HostPath parse_path_name(const std::string &tag) {
	HostPath p; p.tag = tag;
	size_t h1 = tag.find('#'), h2 = h1 == std::string::npos ? h1 : tag.find('#', h1 + 1);
	if (h2 != std::string::npos) {
		std::string hap = tag.substr(h1 + 1, h2 - h1 - 1);
		if (!hap.empty() && hap.find_first_not_of("0123456789") == std::string::npos) { p.pansn = true; p.sample = tag.substr(0, h1); p.hap_id = (u32)strtoul(hap.c_str(), nullptr, 10); }
	}
	if (!p.pansn) p.sample = tag;
	return p;
}

inline HostGraph load_gfa(const std::string &fp, bool inc_labels, bool inc_refs) {
	auto bad = [&](const std::string &d) { return std::runtime_error("Invalid GFA '" + fp + "': " + d); };
	MappedFile mf(fp);
	if (!mf.ok()) throw bad("could not open file");
	const char *B = mf.p, *Eo = mf.p + mf.n;
	struct Line { const char *b, *e; };
	std::vector<Line> S, L, P;
	// pass 1 = the reference's pre-scan (from_gfa.cpp:57-98): record types and S lines, whole file first
	{
		size_t line_no = 1;
		for (const char *b = B; b < Eo;) {
			const char *nl = (const char *)memchr(b, '\n', (size_t)(Eo - b));
			const char *e = nl ? nl : Eo;
			const char *next = nl ? nl + 1 : Eo;
			if (e > b && e[-1] == '\r') e--;
			if (e == b) { line_no++; b = next; continue; }
			switch (*b) {
			case 'H': case 'W': break;
			case 'L': L.push_back({b, e}); break;
			case 'P': P.push_back({b, e}); break;
			case 'S': {
				const char *t1 = (const char *)memchr(b, '\t', (size_t)(e - b));
				if (!t1) throw bad("S record on line " + std::to_string(line_no) + " is missing a segment id and sequence");
				const char *t2 = (const char *)memchr(t1 + 1, '\t', (size_t)(e - t1 - 1));
				if (!t2) throw bad("S record on line " + std::to_string(line_no) + " is missing a sequence");
				const char *sb = t2 + 1; const char *se = (const char *)memchr(sb, '\t', (size_t)(e - sb));
				if (sb >= e || sb == se) throw bad("S record on line " + std::to_string(line_no) + " has an empty sequence");
				S.push_back({b, e});
				break;
			}
			default:
				throw bad("unsupported record type '" + std::string(1, *b) + "' on line " + std::to_string(line_no));
			}
			line_no++; b = next;
		}
	}
	auto field = [](const Line &ln, int k, const char *&fb, const char *&fe) -> bool {   // k-th tab-separated field
		const char *b = ln.b;
		for (int i = 0; i < k; i++) { const char *t = (const char *)memchr(b, '\t', (size_t)(ln.e - b)); if (!t) return false; b = t + 1; }
		const char *t = (const char *)memchr(b, '\t', (size_t)(ln.e - b));
		fb = b; fe = t ? t : ln.e; return true;
	};
	const std::runtime_error lq = bad("liteseq could not parse the file");
	HostGraph g; g.has_labels = inc_labels;
	// segments: ascending numeric id == liteseq's id-indexed vertex table; first S line of an id wins
	struct Seg { u32 id; const char *sb, *se; u32 ord; };
	std::vector<Seg> segs; segs.reserve(S.size());
	for (u32 i = 0; i < S.size(); i++) {
		const char *ib, *ie, *sb, *se; u64 id;
		field(S[i], 1, ib, ie); field(S[i], 2, sb, se);
		if (!parse_u64(ib, ie, id)) id = strtoull(std::string(ib, ie).c_str(), nullptr, 10);   // liteseq uses numeric ids; non-numeric names: parity unpinned
		segs.push_back({(u32)id, sb, se, i});
	}
	if (segs.empty()) throw bad("liteseq returned no vertices");
	bool sorted = true;
	for (size_t i = 1; i < segs.size(); i++) if (segs[i - 1].id >= segs[i].id) { sorted = false; break; }
	if (!sorted) {
		std::stable_sort(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.id < b.id; });
		segs.erase(std::unique(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.id == b.id; }), segs.end());
	}
	const u32 V = (u32)segs.size();
	g.ids.resize(V); g.segLen.resize(V); g.lblOff.assign((size_t)V + 1, 0);
	u64 tot = 0;
	for (u32 i = 0; i < V; i++) { g.ids[i] = segs[i].id; g.segLen[i] = (u32)(segs[i].se - segs[i].sb); g.lblOff[i] = tot; tot += g.segLen[i]; }
	g.lblOff[V] = tot;
	if (inc_labels) { g.lbl.resize(tot); for (u32 i = 0; i < V; i++) memcpy(&g.lbl[g.lblOff[i]], segs[i].sb, g.segLen[i]); }
	g.build_id_table();
	// links
	g.ep1.resize(L.size()); g.ep2.resize(L.size());
	for (size_t i = 0; i < L.size(); i++) {
		const char *f[6][2];
		for (int k = 0; k < 6; k++) if (!field(L[i], k, f[k][0], f[k][1])) throw lq;
		size_t ol = (size_t)(f[5][1] - f[5][0]);
		if (!((ol == 2 && f[5][0][0] == '0' && f[5][0][1] == 'M') || (ol == 1 && f[5][0][0] == '*'))) throw lq;
		u64 a, b;
		if (!parse_u64(f[1][0], f[1][1], a) || !parse_u64(f[3][0], f[3][1], b)) throw lq;
		u32 ia = g.idx_of(a), ib = g.idx_of(b);
		if (ia == NONE32 || ib == NONE32) throw lq;
		bool o1 = (f[2][1] - f[2][0]) == 1 && f[2][0][0] == '+', o2 = (f[4][1] - f[4][0]) == 1 && f[4][0][0] == '+';
		g.ep1[i] = 2 * ia + (o1 ? 1u : 0u);     // `L a + b +` joins a.right to b.left; '-' flips the side
		g.ep2[i] = 2 * ib + (o2 ? 0u : 1u);
	}
	// paths
	g.stepOff.assign(1, 0);
	if (inc_refs) {
		for (size_t r = 0; r < P.size(); r++) {
			const char *nb, *ne, *sb, *se, *ob, *oe;
			if (!field(P[r], 1, nb, ne) || !field(P[r], 2, sb, se) || !field(P[r], 3, ob, oe)) throw lq;
			HostPath p = parse_path_name(std::string(nb, ne));
			u64 pos = 1;
			for (const char *q = sb; q <= se;) {
				const char *cm = (const char *)memchr(q, ',', (size_t)(se - q));
				const char *te = cm ? cm : se;
				if (te - q < 2 || (te[-1] != '+' && te[-1] != '-')) throw lq;
				u64 id;
				if (!parse_u64(q, te - 1, id)) throw lq;
				u32 v = g.idx_of(id);
				if (v == NONE32) throw lq;
				g.steps.push_back((v << 1) | (te[-1] == '-' ? 1u : 0u));
				pos += g.segLen[v];
				if (!cm) break;
				q = cm + 1;
			}
			p.hap_len = (u32)(pos - 1);
			g.paths.push_back(std::move(p));
			if (g.steps.size() >= 0xFFFFFFF0ull) throw std::runtime_error("povu_b200: more than 2^32 path steps in one graph are not supported yet");
			g.stepOff.push_back((u32)g.steps.size());
		}
	}
	return g;
}

// ---------------------------------------------------------------------------------------------
// path selection and genotype columns (refs.hpp:278-301, 359-412)
// ---------------------------------------------------------------------------------------------
inline bool is_prefix(const std::string &p, const std::string &s) { return p.size() <= s.size() && s.compare(0, p.size(), p) == 0; }
inline std::set<u32> refs_in_sample(const HostGraph &g, const std::string &sample) {
	std::set<u32> out;
	for (u32 r = 0; r < g.H(); r++) {
		const HostPath &p = g.paths[r];
		if (p.pansn) { if (p.sample == sample) out.insert(r); }
		else if (is_prefix(sample, p.sample)) out.insert(r);
		if (is_prefix(sample, p.tag)) out.insert(r);
	}
	return out;
}
struct GtMeta {
	std::vector<std::string> col_names; std::vector<u32> ploidy, colOff;   // colOff: first slot of a column in the flat slot array
	std::vector<u32> col, phase;                                           // per path
	u32 nSlots = 0;
};
inline GtMeta genotype_metadata(const HostGraph &g) {
	GtMeta m; const u32 H = g.H();
	m.col.assign(H, NONE32); m.phase.assign(H, 0);
	// refs_in_sample(sample) per path is O(H^2) in the reference; group once by the same predicate instead
	std::vector<char> handled(H, 0);
	for (u32 r = 0; r < H; r++) {
		if (handled[r]) continue;
		handled[r] = 1;
		std::set<u32> same = refs_in_sample(g, g.paths[r].sample);
		u32 hc = (u32)m.col_names.size();
		if (same.size() <= 1) { m.col[r] = hc; m.phase[r] = 0; m.ploidy.push_back(1); }
		else {
			std::set<u32> haps;
			for (u32 x : same) { m.col[x] = hc; m.phase[x] = g.paths[x].hap_id - 1; haps.insert(g.paths[x].hap_id); handled[x] = 1; }
			m.ploidy.push_back((u32)haps.size());
		}
		m.col_names.push_back(g.paths[r].sample);
	}
	m.colOff.assign(m.ploidy.size() + 1, 0);
	for (size_t c = 0; c < m.ploidy.size(); c++) m.colOff[c + 1] = m.colOff[c] + m.ploidy[c];
	m.nSlots = m.colOff.back();
	return m;
}

// ---------------------------------------------------------------------------------------------
// PVST forest: flubbles of all components concatenated
// ---------------------------------------------------------------------------------------------
struct Forest {
	std::vector<u32> comp_id;                  // per tree (file): component id
	std::vector<u32> treeOff;                  // per tree: first flubble (size = trees + 1)
	pv::hvec<u32> a_v, z_v, parent, height, flBase; pv::hvec<u8> a_fwd, z_fwd;   // per flubble; parent/height in PVST indices (page-locked: they are copied to the device)
	pv::hvec<char> family;                  // 'F', 'T', 'O', 'C', 'S', 'M'
	std::vector<u8> orphan;                 // 1 = not reachable from the tree's root through the children lists of the file (empty: none)
	u32 n() const { return (u32)a_v.size(); }
};

inline std::string flubble_label(const HostGraph &g, u32 a_v, bool a_fwd, u32 z_v, bool z_fwd) {
	return std::string(a_fwd ? ">" : "<") + std::to_string(g.ids[a_v]) + (z_fwd ? ">" : "<") + std::to_string(g.ids[z_v]);
}

// read one .pvst (from_pvst.cpp:162-302) and append it to the forest.  Throws where the reference exits.
inline void read_pvst(const std::string &fp, const HostGraph &g, u32 comp_id, Forest &F) {
	MappedFile mf(fp);
	if (!mf.ok()) throw std::runtime_error("could not open " + fp);
	struct Row { char typ; u32 id; std::string label, kids; };
	std::vector<Row> rows;
	size_t line_idx = 0;
	for (const char *b = mf.p, *E = mf.p + mf.n; b < E; line_idx++) {
		const char *nl = (const char *)memchr(b, '\n', (size_t)(E - b));
		const char *e = nl ? nl : E;
		std::vector<std::string> tok;
		for (const char *q = b;;) { const char *t = (const char *)memchr(q, '\t', (size_t)(e - q)); tok.emplace_back(q, t ? t : e); if (!t) break; q = t + 1; }
		b = nl ? nl + 1 : E;
		if (tok.size() != 5) throw std::runtime_error("invalid number of columns. File " + fp + ", line " + std::to_string(line_idx) + ". Expected 5, got " + std::to_string(tok.size()));
		char typ = tok[0].empty() ? '?' : tok[0][0];
		if (typ == 'H') { if (tok[1] != "0.0.3") throw std::runtime_error("Unsupported PVST version in file " + fp + ", got " + tok[1] + ". Supported versions are: 0.0.3."); rows.push_back({'H', 0, "", tok[3]}); continue; }
		rows.push_back({typ, (u32)strtoul(tok[1].c_str(), nullptr, 10), tok[2], tok[3]});
	}
	// vertices in line order; the dummy root must end up as PVST vertex 0 (it is written first)
	std::map<u32, u32> file2pvst; std::vector<u32> row2pvst(rows.size(), NONE32);
	u32 nv = 0;
	const u32 base = F.n();
	for (size_t i = 0; i < rows.size(); i++) {
		Row &r = rows[i];
		if (r.typ == 'H') continue;
		if (r.typ == 'D') { if (nv != 0) throw std::runtime_error("pvst: dummy root is not the first vertex in " + fp); file2pvst[r.id] = nv; row2pvst[i] = nv++; continue; }
		if (strchr("FTOCSM", r.typ) == nullptr) continue;   // unknown vertex type: logged and skipped by the reference
		if (nv == 0) throw std::runtime_error("pvst: no dummy root before the first flubble in " + fp);
		size_t first = r.label.find_first_of("><"), last = r.label.find_last_of("><");
		if (first == std::string::npos || first == last) throw std::runtime_error("pvst: bad label '" + r.label + "' in " + fp);
		u64 a_id = strtoull(r.label.substr(first + 1, last - first - 1).c_str(), nullptr, 10), z_id = strtoull(r.label.substr(last + 1).c_str(), nullptr, 10);
		u32 av = g.idx_of(a_id), zv = g.idx_of(z_id);
		if (av == NONE32 || zv == NONE32) throw std::runtime_error("pvst: label '" + r.label + "' names a segment that is not in the graph (" + fp + ")");
		F.a_v.push_back(av); F.z_v.push_back(zv); F.a_fwd.push_back(r.label[first] == '>'); F.z_fwd.push_back(r.label[last] == '>');
		F.parent.push_back(0); F.height.push_back(1); F.flBase.push_back(base); F.family.push_back(r.typ);
		file2pvst[r.id] = nv; row2pvst[i] = nv++;
	}
	std::vector<std::vector<u32>> kidsOf(nv);      // children lists as listed (a vertex may be listed under several parents)
	for (size_t i = 0; i < rows.size(); i++) {
		if (rows[i].kids == "." || row2pvst[i] == NONE32) continue;
		const std::string &s = rows[i].kids;
		for (size_t pos = 0; pos < s.size();) {
			size_t comma = s.find(',', pos);
			std::string t = s.substr(pos, comma == std::string::npos ? comma : comma - pos);
			size_t a = 0, z = t.size();
			while (a < z && isspace((unsigned char)t[a])) a++;
			while (z > a && isspace((unsigned char)t[z - 1])) z--;
			if (z > a) { char *end = nullptr; std::string tt = t.substr(a, z - a); long v = strtol(tt.c_str(), &end, 10); if (end != tt.c_str()) { auto it = file2pvst.find((u32)v); if (it != file2pvst.end() && it->second > 0) { F.parent[base + it->second - 1] = row2pvst[i]; kidsOf[row2pvst[i]].push_back(it->second); } } }
			if (comma == std::string::npos) break;
			pos = comma + 1;
		}
	}
	// The reference only ever looks at vertices it reaches from the root through the children lists (color.cpp:104-163).  Its
	// sub-flubble passes leave vertices that no list names (`decompose -s`: flubbles whose edge to the parent was deleted and
	// re-added the wrong way round): those are never called.
	{
		std::vector<u8> seen(nv, 0); std::vector<u32> st(1, 0); seen[0] = 1;
		while (!st.empty()) { u32 x = st.back(); st.pop_back(); for (u32 k : kidsOf[x]) if (!seen[k]) { seen[k] = 1; st.push_back(k); } }
		bool any = false;
		for (u32 j = 1; j < nv; j++) if (!seen[j]) any = true;
		if (any || !F.orphan.empty()) { F.orphan.resize(F.n(), 0); for (u32 j = 1; j < nv; j++) if (!seen[j]) F.orphan[base + j - 1] = 1; }
	}
	// heights: root 0, child = parent + 1 (pvst.hpp:807-836); parents precede children in creation order only for
	// files we wrote ourselves, so resolve iteratively
	const u32 n = F.n() - base;
	std::vector<u32> h(n + 1, NONE32); h[0] = 0;
	for (u32 j = 0; j < n; j++) {
		u32 v = j + 1, d = 0; u32 x = v;
		while (h[x] == NONE32) { x = F.parent[base + x - 1]; d++; if (d > n + 1) throw std::runtime_error("pvst: cycle in " + fp); }
		u32 hv = h[x] + d; x = v;
		while (h[x] == NONE32) { h[x] = hv--; x = F.parent[base + x - 1]; }
	}
	for (u32 j = 0; j < n; j++) F.height[base + j] = h[j + 1];
	F.comp_id.push_back(comp_id);
	if (F.treeOff.empty()) F.treeOff.push_back(0);
	F.treeOff.push_back(F.n());
}

// ---------------------------------------------------------------------------------------------
// VCF text
// ---------------------------------------------------------------------------------------------
inline std::string today() {
	time_t t = time(nullptr); struct tm *lt = localtime(&t);
	char b[32]; snprintf(b, sizeof b, "%04d%02d%02d", lt->tm_year + 1900, lt->tm_mon + 1, lt->tm_mday);
	return b;
}
inline std::string vcf_header_common(bool with_date) {
	std::string o = "##fileformat=VCFv4.2\n";
	if (with_date) o += "##fileDate=" + today() + "\n";
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
	return o;
}
inline std::string vcf_contig_line(const HostPath &p) { return "##contig=<ID=" + p.tag + ",length=" + std::to_string(p.hap_len) + ">\n"; }
inline std::string vcf_col_header(const GtMeta &gm) {
	std::string o = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
	for (auto &c : gm.col_names) { o += '\t'; o += c; }
	return o + "\n";
}

// Binary records as produced by the device (pv_call.cuh); their VCF text is written on the device (pv_text.cuh)
typedef pv::Records Records;

// Record order of the reference (genomics.cpp:114-156, to_vcf.cpp:121-139): chunks of `chunk_size` regions,
// inside a chunk grouped by ref path id ascending, then region order, then context / INS,SUB,DEL order.
inline std::vector<size_t> record_order(const Records &R, size_t chunk_size) {
	std::vector<size_t> ord(R.n());
	for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
	if (chunk_size == 0) chunk_size = 1;
	std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
		// SUBR records hang off per-ref pseudo regions after the real ones: they are emitted with the final chunk,
		// after that chunk's ordinary records of the same ref (genomics.cpp:137-150, vcf.cpp:275-312)
		const size_t last = R.n_regions ? (size_t)R.n_regions - 1 : 0;
		size_t ca = std::min<size_t>(R.rec_rov[a], last) / chunk_size, cb = std::min<size_t>(R.rec_rov[b], last) / chunk_size;
		if (ca != cb) return ca < cb;
		if (R.rec_ref[a] != R.rec_ref[b]) return R.rec_ref[a] < R.rec_ref[b];
		return R.rec_rov[a] < R.rec_rov[b];
	});
	return ord;
}

} // namespace pvh
