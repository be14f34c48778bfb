// povu_b200 -- session layer: one loaded graph on one device, tying the host formats (pv_host.hpp) to
// the device pipelines (pv_decompose.cuh, pv_call.cuh).  This is what the `povu` CLI clone, the povu_ffi
// C API and the pvgpu_session_* C ABI all sit on; it mirrors the reference's drivers
//   do_decompose  /root/reference/app/subcommand/decompose.cpp:94-160
//   do_call       /root/reference/app/subcommand/call.cpp:71-173
//   do_gfa2vcf    /root/reference/app/subcommand/gfa2vcf.cpp:18-87
#pragma once
#include <cerrno>
#include <chrono>
#include <dirent.h>
#include <array>
#include <fstream>
#include <type_traits>
#include "pv_text.cuh"
#include "pv_gfa.cuh"
#include "pv_shard.cuh"
#include "pv_host.hpp"

namespace pve {

using pv::u32; using pv::u64; using pv::u8;

struct Session {
	pv::Ctx *c = nullptr;
	pvh::HostGraph g;
	pvh::GtMeta gm; bool gm_ready = false;
	pv::Buf<u32> d_ep1, d_ep2;
	pv::DevGraph dg; bool edges_up = false, paths_up = false;
	pvh::Forest forest; bool forest_ready = false; int forest_subfl = -1; std::vector<u32> tree_label;   // tree_label: component label per tree (call slices the forest by it)   // forest_subfl: does the forest hold sub-flubble families (-1 = not looked yet)
	pv::DecompOut dec;
	bool debug_numbering = false; pv::RefNumOut refnum;   // --hairpins / --structure-export: decompose also derives the reference's internal numbering (pv_refnum.cuh)
	mutable pvh::Records rec;    // binary records + counters of the last call: born in HBM, copied to the host on demand (rec.to_host)
	std::vector<std::string> prefixes;
	std::set<u32> ref_ids;

	explicit Session(pv::Ctx &ctx) : c(&ctx) {}
	~Session() { unpin(); pinned.swap(full_pinned); unpin(); }
	Session(const Session &) = delete;
	Session &operator=(const Session &) = delete;
	// page-lock the host arrays that are copied to the device (so H2D runs at PCIe speed, asynchronously)
	std::vector<void *> pinned;
	void pin() {
#ifndef PV_EMU
		auto reg = [&](const void *p, size_t bytes) { if (p && bytes && cudaHostRegister((void *)p, bytes, cudaHostRegisterDefault) == cudaSuccess) pinned.push_back((void *)p); else cudaGetLastError(); };
		reg(g.ep1.data(), g.ep1.size() * 4); reg(g.ep2.data(), g.ep2.size() * 4); reg(g.steps.data(), g.steps.size() * 4); reg(g.segLen.data(), g.segLen.size() * 4);
#endif
	}
	void unpin() {
#ifndef PV_EMU
		for (void *p : pinned) cudaHostUnregister(p);
#endif
		pinned.clear();
	}

	// GFA -> graph.  Default: the text is parsed on the device (pv_gfa.cuh); the big arrays (links, walks, sequences)
	// then exist only in HBM and come to the host on demand (ensure_host).  The sequential host loader of pv_host.hpp is
	// reachable only through the test hook pvgpu_session_open_hostparse (same results; the checker of the parser's tests
	// and the "host_loader_ms" figure of the bench).
	bool host_edges = true, host_steps = true, host_lbl = true, tt_graph_dev = false;
	double t_gfa_read_ms = 0, t_gfa_parse_ms = 0; u64 gfa_text_bytes = 0;
	void load_gfa(const std::string &fp, bool labels, bool refs, bool force_host_parser = false) {
		unpin(); c->bg_sync();
		d_ep1.release(); d_ep2.release(); d_steps.release(); d_segLen.release(); dg = pv::DevGraph();
		edges_up = raw_paths_up = paths_up = false; sharded = false;
		invalidate_graph_state();
		// force_host_parser: test hook only (pvgpu_session_open_hostparse), the checker of the device parser
		if (force_host_parser) { double t0 = now_ms(); g = pvh::load_gfa(fp, labels, refs); t_gfa_read_ms = 0; t_gfa_parse_ms = now_ms() - t0; host_edges = host_steps = host_lbl = true; pin(); return; }
		load_gfa_device(fp, labels, refs);
	}
	void load_gfa_device(const std::string &fp, bool labels, bool refs) {
		auto bad = [&](const std::string &d) { return std::runtime_error("Invalid GFA '" + fp + "': " + d); };
		struct Fd { int fd; ~Fd() { if (fd >= 0) close(fd); } } f{open(fp.c_str(), O_RDONLY)};
		struct stat st;
		if (f.fd < 0 || fstat(f.fd, &st) != 0 || !S_ISREG(st.st_mode)) throw bad("could not open file");
		const u64 n = (u64)st.st_size;
		char last = 0;
		if (n && pread(f.fd, &last, 1, (off_t)(n - 1)) != 1) throw bad("could not open file");
		double t0 = now_ms();
		pv::GfaDev o;
		gfa_text_bytes = n;
		{
			pv::Buf<u8> T;
			pv::gfa_text_to_device(*c, f.fd, n, T);
			t_gfa_read_ms = now_ms() - t0;
			pv::gfa_parse_device(*c, T.p, n, last == '\n', labels, refs, o);
		}
		if (o.err != ~0ull) {
			const std::string line = std::to_string(o.err >> 16); const u32 code = (u32)((o.err >> 8) & 0xFF);
			if (code == 1) throw bad("S record on line " + line + " is missing a segment id and sequence");
			if (code == 2) throw bad("S record on line " + line + " is missing a sequence");
			if (code == 3) throw bad("S record on line " + line + " has an empty sequence");
			throw bad("unsupported record type '" + std::string(1, (char)(o.err & 0xFF)) + "' on line " + line);
		}
		if (o.V == 0) throw bad("liteseq returned no vertices");
		if (o.lq) throw bad("liteseq could not parse the file");
		if (o.S >= 0xFFFFFFF0ull) throw std::runtime_error("povu_b200: more than 2^32 path steps in one graph are not supported yet");
		g = pvh::HostGraph(); g.has_labels = labels;
		g.ids.swap(o.h_ids); g.segLen.swap(o.h_segLen); g.stepOff = o.h_stepOff; g.devE = o.E;
		for (u32 r = 0; r < o.H; r++) {
			std::string tag((size_t)(o.nameE[r] - o.nameB[r]), '\0');
			if (!tag.empty() && pread(f.fd, &tag[0], tag.size(), (off_t)o.nameB[r]) != (ssize_t)tag.size()) throw bad("could not open file");
			pvh::HostPath p = pvh::parse_path_name(tag);
			p.hap_len = (u32)o.hapLen[r];
			g.paths.push_back(std::move(p));
		}
		d_ep1 = std::move(o.ep1); d_ep2 = std::move(o.ep2); edges_up = true;
		d_steps = o.steps.p ? std::move(o.steps) : pv::Buf<u32>(*c, 1); d_segLen = std::move(o.segLen); raw_paths_up = true;
		tt.ids = std::move(o.ids); tt.lblOff = std::move(o.lblOff);
		if (labels) tt.lbl = std::move(o.lbl);
		tt_graph_dev = true;
		host_edges = o.E == 0; host_steps = g.stepOff.back() == 0; host_lbl = false;
		t_gfa_parse_ms = now_ms() - t0 - t_gfa_read_ms;
	}
	// bring the arrays that a device-parsed graph keeps only in HBM to the host (sharding, the povu_ffi getters, re-upload)
	void ensure_host(bool edges, bool steps, bool lbl) {
		if (edges && !host_edges) { g.ep1.resize(g.devE); g.ep2.resize(g.devE); c->d2h(g.ep1.data(), d_ep1.p, g.devE * 4); c->d2h(g.ep2.data(), d_ep2.p, g.devE * 4); host_edges = true; }
		if (steps && !host_steps) { g.steps.resize(g.stepOff.back()); c->d2h(g.steps.data(), d_steps.p, g.steps.size() * 4); host_steps = true; }
		if (lbl && !host_lbl) {
			g.lblOff.resize((size_t)g.V() + 1); c->d2h(g.lblOff.data(), tt.lblOff.p, g.lblOff.size() * 8);
			if (g.has_labels) { g.lbl.resize(g.lblOff.back()); c->d2h(&g.lbl[0], tt.lbl.p, g.lbl.size()); }
			host_lbl = true;
		}
	}

	// raw arrays resident in HBM: edge ports, packed steps, segment lengths
	pv::Buf<u32> d_steps, d_segLen; bool raw_paths_up = false;
	pv::Buf<u32> d_labels; bool labels_ready = false, want_labels = false;
	pv::DevForest dforest;   // decompose's flubble table in HBM, alive from decompose to the end of call on the hot path   // whole-graph component labels handed from decompose to the index build (hot path only)
	double t_upload_ms = 0, t_dec_ms = 0, t_build_ms = 0, t_call_ms = 0;
	static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

	void upload_edges() {
		if (edges_up) return;
		const u64 E = g.E();
		d_ep1 = pv::Buf<u32>(*c, E ? E : 1); d_ep2 = pv::Buf<u32>(*c, E ? E : 1);
		c->h2d(d_ep1.p, g.ep1.data(), E * 4); c->h2d(d_ep2.p, g.ep2.data(), E * 4);
		c->sync();
		edges_up = true;
	}
	void upload_raw_paths() {
		if (raw_paths_up) return;
		if (g.stepOff.empty()) g.stepOff.assign(1, 0);
		const u32 S = g.stepOff.back();
		d_steps = pv::Buf<u32>(*c, S ? S : 1); d_segLen = pv::Buf<u32>(*c, g.V() ? g.V() : 1);
		c->h2d(d_steps.p, g.steps.data(), (size_t)S * 4); c->h2d(d_segLen.p, g.segLen.data(), (size_t)g.V() * 4);
		c->sync();
		raw_paths_up = true;
	}
	// same, but the copies run on the upload stream while the main stream goes on (the walks are first needed by the index
	// build, after decompose); whoever reads d_steps / d_segLen next calls c->up_join() first
	void upload_raw_paths_async() {
		if (raw_paths_up) return;
		if (g.stepOff.empty()) g.stepOff.assign(1, 0);
		const u32 S = g.stepOff.back();
		d_steps = pv::Buf<u32>(*c, S ? S : 1); d_segLen = pv::Buf<u32>(*c, g.V() ? g.V() : 1);
		c->up_begin();
		c->h2d_up(d_steps.p, g.steps.data(), (size_t)S * 4); c->h2d_up(d_segLen.p, g.segLen.data(), (size_t)g.V() * 4);
		raw_paths_up = true;
	}
	// derived device index (adjacency in L-line order, loci, vertex -> steps lists, path components)
	void upload_paths() {
		if (paths_up) return;
		upload_edges(); upload_raw_paths();
		c->up_join();
		pv::graph_build(*c, g.V(), g.E(), d_ep1.p, d_ep2.p, g.H(), g.stepOff.data(), d_steps.p, d_segLen.p, dg, labels_ready ? &d_labels : nullptr);
		labels_ready = false; d_labels.release();
		paths_up = true;
	}
	// the loaded graph changed (new file, shard, builder mutation): nothing derived from the previous graph may survive
	void invalidate_graph_state() {
		forest_ready = false; gm_ready = false; forest_subfl = -1; tree_label.clear();
		tt = pv::TextTables(); tt_graph_dev = false; pvst_cache_ok = false;
		dforest = pv::DevForest(); labels_ready = false; d_labels.release();
		rec.reset();
	}
	// forget everything on the device (the next decompose / call re-uploads from the host arrays)
	void drop_device() {
		ensure_host(true, true, false);
		c->up_join();
		d_ep1.release(); d_ep2.release(); d_steps.release(); d_segLen.release(); dg = pv::DevGraph();
		edges_up = raw_paths_up = paths_up = false;
	}
	// the whole hot path once: decompose -> index -> call.  keep_raw: the raw arrays stay resident in HBM
	// (bench `value`); otherwise they are uploaded from the host inside this call (bench `e2e`).
	void hot_path(bool keep_raw) {
		double t0 = now_ms();
		if (!keep_raw) drop_device();
		upload_edges(); upload_raw_paths_async();
		double t1 = now_ms();
		want_labels = true;
		decompose();
		want_labels = false;
		double t2 = now_ms();
		paths_up = false; dg = pv::DevGraph();
		upload_paths();
		double t3 = now_ms();
		call();
		if (!keep_raw) rec.to_host(*c);   // e2e: the binary records come to the host; otherwise they stay in HBM for the device text writer
		c->bg_sync();              // the flubble table has reached the host (copied behind the index build and the call)
		double t4 = now_ms();      // (the device copy of the forest stays for later calls of this session)
		t_upload_ms = t1 - t0; t_dec_ms = t2 - t1; t_build_ms = t3 - t2; t_call_ms = t4 - t3;
	}

	// ---- multi-GPU: keep only this rank's share of the weakly connected components.
	// Components are independent units for decompose, and every path lies inside one component, so ranks never
	// exchange anything while computing.  Balancing: greedy longest-processing-time on (edges + path steps).
	std::vector<u32> comp_global;          // local component k (ascending lowest vertex) -> global component id (1-based)
	u32 n_comp_total = 0; bool sharded = false;
	std::vector<u32> shard_owner;          // per global component (0-based): owning rank
	// The whole graph, kept aside by the first shard() so that the partitioning can be repeated (the multi-GPU end-to-end
	// step re-shards from the page-locked host arrays every time): host side in `full`, device side in f_*.
	pvh::HostGraph full; bool have_full = false, full_dev = false, full_steps_dev = false;
	pv::Buf<u32> f_ep1, f_ep2, f_steps, f_segLen, f_ids; pv::Buf<u64> f_lblOff; pv::Buf<u8> f_lbl;
	std::vector<void *> full_pinned;
	double t_shard_ms = 0;
	// from_host: forget the device copy of the whole graph first, so that this call uploads it from the host arrays again
	void shard(u32 rank, u32 world, bool from_host = false) {
		const double t0 = now_ms();
		pv::Trace trs(*c, "session.shard");
		if (!have_full) {
			upload_edges(); upload_raw_paths(); c->up_join();
			if (from_host) ensure_host(true, true, true);
			const bool dev_tables = tt_graph_dev && tt.ids.p;
			f_ep1 = std::move(d_ep1); f_ep2 = std::move(d_ep2); f_steps = std::move(d_steps); f_segLen = std::move(d_segLen);
			if (dev_tables) { f_ids = std::move(tt.ids); if (g.has_labels) { f_lblOff = std::move(tt.lblOff); f_lbl = std::move(tt.lbl); } }
			else {
				f_ids = pv::Buf<u32>(*c, g.V() ? g.V() : 1); c->h2d(f_ids.p, g.ids.data(), (size_t)g.V() * 4);
				if (g.has_labels) { f_lblOff = pv::Buf<u64>(*c, g.lblOff.size()); c->h2d(f_lblOff.p, g.lblOff.data(), g.lblOff.size() * 8); f_lbl = pv::Buf<u8>(*c, g.lbl.size() ? g.lbl.size() : 1); c->h2d(f_lbl.p, g.lbl.data(), g.lbl.size()); }
				c->sync();
			}
			full_pinned.swap(pinned);      // (the page-locked registrations belong to the whole graph's arrays)
			full = std::move(g);
			full.devE = full.ep1.empty() ? full.devE : full.ep1.size();
			have_full = true; full_dev = true; full_steps_dev = true;
		}
		if (from_host && full_dev) { f_ep1.release(); f_ep2.release(); f_steps.release(); f_segLen.release(); f_ids.release(); f_lblOff.release(); f_lbl.release(); full_dev = false; }
		const u32 V = full.V(), H = full.H(); const u64 E = full.E(); const u64 S = full.stepOff.empty() ? 0 : full.stepOff.back();
		if (!full_dev) {   // the whole graph travels host -> HBM again (page-locked arrays: PCIe speed)
			if ((E && full.ep1.empty()) || (S && full.steps.empty())) throw std::runtime_error("shard: the whole graph is no longer on the host");
			// links and segments of the whole graph (the component labels need every link); of the walks only this rank's own
			// paths will travel (shard_device, host-walk mode): f_steps stays empty
			f_ep1 = pv::Buf<u32>(*c, E ? E : 1); f_ep2 = pv::Buf<u32>(*c, E ? E : 1); f_segLen = pv::Buf<u32>(*c, V ? V : 1); f_ids = pv::Buf<u32>(*c, V ? V : 1);
			c->h2d(f_ep1.p, full.ep1.data(), E * 4); c->h2d(f_ep2.p, full.ep2.data(), E * 4);
			c->h2d(f_segLen.p, full.segLen.data(), (size_t)V * 4); c->h2d(f_ids.p, full.ids.data(), (size_t)V * 4);
			if (full.has_labels) { f_lblOff = pv::Buf<u64>(*c, full.lblOff.size()); c->h2d(f_lblOff.p, full.lblOff.data(), full.lblOff.size() * 8); f_lbl = pv::Buf<u8>(*c, full.lbl.size() ? full.lbl.size() : 1); c->h2d(f_lbl.p, full.lbl.data(), full.lbl.size()); }
			full_dev = true; full_steps_dev = false;
		}
		pv::ShardIn in; in.V = V; in.H = H; in.E = E; in.S = S; in.ep1 = f_ep1.p; in.ep2 = f_ep2.p; in.steps = full_steps_dev ? f_steps.p : nullptr; in.h_steps = full.steps.data(); in.segLen = f_segLen.p; in.ids = f_ids.p;
		if (full.has_labels) { in.lblOff = f_lblOff.p; in.lbl = f_lbl.p; }
		if (full.stepOff.empty()) full.stepOff.assign(1, 0);
		in.h_stepOff = full.stepOff.data();
		pv::ShardOut o;
		trs.mark("whole graph host -> HBM");
		pv::shard_device(*c, in, rank, world, o);
		trs.mark("shard_device");
		// install the share as the session's graph: the big arrays exist only in HBM (ensure_host brings them on demand)
		c->bg_sync(); c->up_join();
		invalidate_graph_state();
		dg = pv::DevGraph(); paths_up = false;
		g = pvh::HostGraph(); g.has_labels = full.has_labels; g.paths = full.paths;
		g.ids.swap(o.h_ids); g.segLen.swap(o.h_segLen); g.stepOff = o.h_stepOff; g.devE = o.E;
		g.build_id_table();
		d_ep1 = std::move(o.ep1); d_ep2 = std::move(o.ep2); d_steps = std::move(o.steps); d_segLen = std::move(o.segLen);
		edges_up = true; raw_paths_up = true;
		tt.ids = std::move(o.ids);
		if (full.has_labels) { tt.lblOff = std::move(o.lblOff); tt.lbl = std::move(o.lbl); }
		tt_graph_dev = true;
		host_edges = o.E == 0; host_steps = o.S == 0; host_lbl = false;
		comp_global = o.comp_global; shard_owner = o.owner; n_comp_total = o.n_comp_total;
		sharded = true;
		trs.mark("install the share");
		t_shard_ms = now_ms() - t0;
	}

	// decompose: flubbles + PVST of every component with >= 3 vertices (decompose.cpp:129-142)
	void decompose() {
		c->bg_sync();
		upload_edges();
		pv::DecompOpts opt;
		dforest = pv::DevForest();
		// the flubble table and the component labels stay in HBM for the calls that follow (no re-upload, no second
		// connectivity pass); on the hot path the host copies of the results additionally travel behind the next phases
		opt.keep_labels = &d_labels; opt.keep_forest = &dforest;
		if (want_labels) opt.async_results = true;
		refnum.reset();
		if (debug_numbering) { opt.want_stack = true; opt.refnum = &refnum; }
		pv::Trace tr(*c, "session");
		pv::decompose_device(*c, g.V(), g.E(), d_ep1.p, d_ep2.p, opt, dec);
		labels_ready = d_labels.p != nullptr && (want_labels || !paths_up);   // (the hot path rebuilds the index right after)
		if (!labels_ready) d_labels.release();
		tr.mark("decompose_device");
		if (dec.collisions) throw std::runtime_error("cycle-equivalence label collision detected by the exact bracket-set check");
		// the flubble table is already flat in component order: swap the (page-locked) arrays instead of copying.
		// Components with < 3 vertices own no flubbles and get no tree (decompose.cpp:135-142).
		forest.a_v.swap(dec.a_v); forest.z_v.swap(dec.z_v); forest.a_fwd.swap(dec.a_fwd); forest.z_fwd.swap(dec.z_fwd);
		forest.parent.swap(dec.parent); forest.height.swap(dec.depth); forest.flBase.swap(dec.fl_base);
		forest.family.assign(forest.a_v.size(), 'F'); forest.orphan.clear();
		forest.comp_id.clear(); forest.treeOff.assign(1, 0);
		u32 base = 0;
		for (u32 k = 0; k < dec.comp_nv.size(); k++) {
			base += dec.comp_nfl[k];
			if (dec.comp_nv[k] >= 3) { forest.comp_id.push_back(sharded ? comp_global.at(k) : k + 1); forest.treeOff.push_back(base); }
			else if (dec.comp_nfl[k]) throw std::runtime_error("decompose: flubbles in a component with fewer than 3 vertices");
		}
		forest_ready = true; pvst_cache_ok = false; forest_subfl = 0; tree_label.clear();
		tr.mark("forest conversion");
	}

	// .pvst text of every tree, written on the device (pv_text.cuh: pvst_text_device) once per forest
	mutable pv::PvstTextOut pvst_cache; mutable bool pvst_cache_ok = false;
	void build_pvst_text() const {
		if (pvst_cache_ok) return;
		c->bg_sync();
		if (!tt.ids.p) { tt.ids = pv::Buf<u32>(*c, g.ids.size() ? g.ids.size() : 1); c->h2d(tt.ids.p, g.ids.data(), g.ids.size() * 4); }
		pv::pvst_text_device(*c, tt.ids.p, forest.n(), (u32)forest.comp_id.size(), forest.treeOff.data(), forest.a_v.data(), forest.a_fwd.data(),
				     forest.z_v.data(), forest.z_fwd.data(), forest.parent.data(), forest.family.data(), pvst_cache);
		pvst_cache_ok = true;
	}
	std::string pvst_text(u32 tree) const {
		build_pvst_text();
		return std::string(pvst_cache.text.data() + pvst_cache.treeOff[tree], (size_t)(pvst_cache.treeOff[tree + 1] - pvst_cache.treeOff[tree]));
	}
	void write_pvsts(const std::string &dir) const {
		build_pvst_text();
		for (u32 t = 0; t < forest.comp_id.size(); t++) {
			std::string fp = dir + "/" + std::to_string(forest.comp_id[t]) + ".pvst";
			int fd = open(fp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
			if (fd < 0) throw std::runtime_error("could not open " + fp + " for writing");
			const char *p = pvst_cache.text.data() + pvst_cache.treeOff[t]; size_t left = (size_t)(pvst_cache.treeOff[t + 1] - pvst_cache.treeOff[t]);
			while (left) { ssize_t w = write(fd, p, left); if (w <= 0) { close(fd); throw std::runtime_error("could not write " + fp); } p += w; left -= (size_t)w; }
			close(fd);
		}
	}
	// every *.pvst of a directory (call.cpp:36-57).  The reference visits them in directory order
	// (unspecified); here: ascending component id, so the output is reproducible.
	void read_pvsts(const std::string &dir) {
		DIR *d = opendir(dir.c_str());
		if (!d) throw std::runtime_error("Directory does not exist: " + dir);
		std::vector<std::pair<u64, std::string>> files;
		while (struct dirent *e = readdir(d)) {
			std::string n = e->d_name;
			if (n.size() > 5 && n.compare(n.size() - 5, 5, ".pvst") == 0) files.push_back({strtoull(n.c_str(), nullptr, 10), n});
		}
		closedir(d);
		if (files.empty()) throw std::runtime_error("Could not find pvst files in " + dir);
		std::sort(files.begin(), files.end());
		forest = pvh::Forest();
		for (auto &f : files) pvh::read_pvst(dir + "/" + f.second, g, (u32)f.first, forest);
		forest_ready = true; pvst_cache_ok = false; dforest = pv::DevForest(); forest_subfl = -1; tree_label.clear();
	}

	void set_prefixes(const std::vector<std::string> &p) {
		prefixes = p; ref_ids.clear();
		for (auto &x : p) for (u32 r : pvh::refs_in_sample(g, x)) ref_ids.insert(r);
		if (!gm_ready) { gm = pvh::genotype_metadata(g); gm_ready = true; }
	}

	// -g/--restrict ref:start-end -- parse_genomic_region (/root/reference/src/ita/variation/rov.cpp:51-118), same messages
	bool restrict_on = false; std::string restrict_ref; u64 restrict_start = 0, restrict_end = 0;
	void set_restrict(const char *region) {
		restrict_on = false;
		if (!region) return;
		const std::string r = region;
		auto bad = [&](const std::string &why) { throw std::runtime_error("Invalid region format '" + r + "': " + why); };
		size_t colon = r.find(':');
		if (colon == std::string::npos) bad("missing colon separator");
		std::string name = r.substr(0, colon);
		if (name.empty()) bad("empty reference name");
		std::string range = r.substr(colon + 1);
		size_t dash = range.find('-');
		if (dash == std::string::npos) bad("missing dash separator in range");
		std::string a = range.substr(0, dash), b = range.substr(dash + 1);
		for (char ch : a) if (!isdigit((unsigned char)ch)) bad("non-numeric start position");
		for (char ch : b) if (!isdigit((unsigned char)ch)) bad("non-numeric end position");
		errno = 0;
		u64 st = (u32)strtoull(a.c_str(), nullptr, 10), en = (u32)strtoull(b.c_str(), nullptr, 10);
		if (a.empty() || b.empty() || errno == ERANGE) throw std::runtime_error("Failed to parse region '" + r + "': stoull");   // std::stoull throws   // pt::idx_t is 32 bits wide: std::stoull's result is truncated
		if (st >= en) throw std::runtime_error("Invalid region '" + r + "': start position (" + std::to_string(st) + ") must be less than end position (" + std::to_string(en) + ")");
		restrict_on = true; restrict_ref = name; restrict_start = st; restrict_end = en;
	}

	void call() {
		if (!forest_ready) throw std::runtime_error("call: no PVST forest (run decompose or read .pvst files first)");
		upload_paths();
		std::vector<u32> refs(ref_ids.begin(), ref_ids.end());
		// the sub-flubble CLAN -- concealed, smothered, midi (pvst.hpp:110-125) -- and everything below such a vertex is never
		// selected (color.cpp:117-125); tiny (T) and parallel (O) vertices are flubble-like and are called like flubbles
		std::vector<u8> excl;
		auto subfl = [](char f) { return f == 'C' || f == 'S' || f == 'M'; };
		if (forest_subfl < 0) { forest_subfl = forest.orphan.empty() ? 0 : 1; for (char f : forest.family) if (subfl(f)) { forest_subfl = 1; break; } }   // once per forest, not per call
		const bool any_excl = forest_subfl == 1;
		if (any_excl) {
			excl.assign(forest.n(), 0);
			std::vector<u32> ord(forest.n());
			for (u32 i = 0; i < ord.size(); i++) ord[i] = i;
			std::stable_sort(ord.begin(), ord.end(), [&](u32 a, u32 b) { return forest.height[a] < forest.height[b]; });
			for (u32 f : ord) {
				u32 p = forest.parent[f];
				if (subfl(forest.family[f]) || (f < forest.orphan.size() && forest.orphan[f]) || (p != 0 && excl[forest.flBase[f] + p - 1])) excl[f] = 1;   // (orphans: vertices the file's children lists never reach from the root)
			}
		}
		pv::CallIn in;
		in.nFl = forest.n(); in.nRefs = (u32)refs.size(); in.refs = refs.data();
		// A flubble is selectable only if every reference path visits both of its boundaries (color.cpp:15-30), so when all
		// reference paths lie in ONE component only that component's flubbles can be called: hand call_device that slice of
		// the forest (the usual whole-genome case: one run per chromosome reference over a graph of 25 components).
		u32 slice_b = 0, slice_n = forest.n();
		if (!refs.empty() && forest.comp_id.size() > 1 && dg.pathComp.size() == g.H()) {
			u32 pc = dg.pathComp[refs[0]]; bool one = pc != pvh::NONE32;
			for (u32 r : refs) if (dg.pathComp[r] != pc) one = false;
			if (one) {
				// the component's flubbles = the tree whose first flubble lies in the component of the reference path's first vertex
				if (tree_label.size() != forest.comp_id.size()) {   // once per forest: component label (lowest vertex index) of every tree,
					const u32 T = (u32)forest.comp_id.size();       // gathered by one kernel and one copy (a whole-genome graph has thousands of trees)
					const bool dev = dforest.valid && dforest.n == forest.n();
					std::vector<u32> first(T, pvh::NONE32);
					for (u32 t = 0; t < T; t++) { const u32 b = forest.treeOff[t], e = forest.treeOff[t + 1]; if (e > b) first[t] = dev ? b : forest.a_v[b]; }
					pv::Buf<u32> dFirst(*c, T), dLab(*c, T);
					c->h2d(dFirst.p, first.data(), (size_t)T * 4);
					const u32 *first_ = dFirst.p, *av_ = dev ? dforest.a_v.p : nullptr, *vcomp_ = dg.vcomp.p; u32 *lab_ = dLab.p;
					pv::parallel_for<pv::C_keep>(*c, T, PV_LAMBDA(size_t t) { u32 x = first_[t]; lab_[t] = x == pv::NONE ? pv::NONE : vcomp_[av_ ? av_[x] : x]; });
					tree_label = dLab.to_host();
				}
				const u32 label = dg.pathLabel[refs[0]];   // lowest vertex index of the references' component (graph_build)
				slice_n = 0;                                                              // no tree: the component owns no flubbles
				for (u32 t = 0; t < forest.comp_id.size(); t++)
					if (tree_label[t] == label) { slice_b = forest.treeOff[t]; slice_n = forest.treeOff[t + 1] - slice_b; break; }
			}
		}
		in.nFl = slice_n; in.fl_offset = slice_b;
		if (any_excl) in.excluded = excl.data() + slice_b;
		if (dforest.valid && dforest.n == forest.n()) {
			in.on_device = true; in.a_v = dforest.a_v.p + slice_b; in.z_v = dforest.z_v.p + slice_b; in.a_fwd = dforest.a_fwd.p + slice_b; in.z_fwd = dforest.z_fwd.p + slice_b;
			in.parent = dforest.parent.p + slice_b; in.flBase = dforest.fl_base.p + slice_b;
		} else {
			in.a_v = forest.a_v.data() + slice_b; in.z_v = forest.z_v.data() + slice_b; in.a_fwd = forest.a_fwd.data() + slice_b; in.z_fwd = forest.z_fwd.data() + slice_b;
			in.parent = forest.parent.data() + slice_b; in.flBase = forest.flBase.data() + slice_b;
		}
		if (restrict_on) {
			in.restrict_on = true; in.restrict_start = restrict_start; in.restrict_end = restrict_end; in.restrict_ref = pvh::NONE32;
			for (u32 h = 0; h < g.H(); h++) if (g.paths[h].tag == restrict_ref) { in.restrict_ref = h; break; }   // Refs::get_ref_id (refs.hpp:265-274)
			if (in.restrict_ref == pvh::NONE32) fprintf(stderr, "ERR [povu::rov::gen_rov] Reference path '%s' not found in graph\n", restrict_ref.c_str());
		}
		pv::CallOpts opt;
		if (const char *e = getenv("POVU_B200_CELL_BUDGET")) opt.cell_budget = strtoull(e, nullptr, 10);
		pv::Trace tr(*c, "session");
		pv::call_device(*c, dg, in, opt, rec);
		tr.mark("call_device");
	}

	// ---- the records of the last call as ONE self-describing binary stream in HBM: what a rank contributes to the final
	// gather of a multi-GPU run (NCCL moves the bytes; nothing is converted on the host).  Layout (little endian, every array
	// 8-byte aligned): u64 header[8] = {magic "PVREC001", nRec, nGrp, nGt, H, 0, 0, 0}; then u32[nRec] region, ref path, left and
	// right context step ((SEGMENT ID << 1) | strand -- ids, not shard-local indices), VARTYPE, POS, ref slice start, ref slice
	// length; u32[nRec + 1] first alt group; u64[nRec + 1] first genotype cell; u32[nGrp] alt path, start, length, count;
	// u16[nGt] genotype cells.
	static u64 pad8(u64 x) { return (x + 7) & ~(u64)7; }
	u64 records_packed_bytes() const {
		const u64 R = rec.nRecTot, G = rec.nGrpTot, T = rec.nGtTot;
		return 64 + 8 * pad8(R * 4) + pad8((R + 1) * 4) + (R + 1) * 8 + 4 * pad8(G * 4) + pad8(T * 2);
	}
	void records_pack(void *dst, u64 cap, u64 *bytes) {
		const u64 need = records_packed_bytes();
		if (bytes) *bytes = need;
		if (cap < need) throw std::runtime_error("records_pack: destination too small");
		pv::Ctx &cx = *c;
		const pv::RecPiece *Rp = rec.nRecTot ? &rec.flat(cx) : nullptr;
		const u64 R = rec.nRecTot, G = rec.nGrpTot, T = rec.nGtTot;
		u64 hdr[8] = {0x3130304345525650ull, R, G, T, (u64)g.H(), 0, 0, 0};
		char *d = (char *)dst; u64 o = 0;
		cx.h2d(d, hdr, 64); o = 64;
		if (!tt.ids.p) { tt.ids = pv::Buf<u32>(cx, g.ids.size() ? g.ids.size() : 1); cx.h2d(tt.ids.p, g.ids.data(), g.ids.size() * 4); }
		auto put = [&](const void *src, u64 n) { if (n) cx.d2d(d + o, src, n); o += pad8(n); };
		if (Rp) {
			const pv::RecPiece &P = *Rp;
			put(P.rov.p, R * 4); put(P.ref.p, R * 4);
			for (int side = 0; side < 2; side++) {
				const u32 *src = side ? P.cxtR.p : P.cxtL.p; u32 *out = (u32 *)(d + o); const u32 *ids = tt.ids.p;
				pv::parallel_for<pv::C_keep>(cx, R, PV_LAMBDA(size_t i) { u32 x = src[i]; out[i] = (ids[x >> 1] << 1) | (x & 1u); });
				o += pad8(R * 4);
			}
			put(P.vt.p, R * 4); put(P.pos.p, R * 4); put(P.refStart.p, R * 4); put(P.refLen.p, R * 4);
			put(P.grpOff.p, (R + 1) * 4); put(P.gtOff.p, (R + 1) * 8);
			put(P.gHap.p, G * 4); put(P.gStart.p, G * 4); put(P.gLen.p, G * 4); put(P.gCnt.p, G * 4);
			put(P.gt.p, T * 2);
		} else { cx.memset0(d + o, need - o); }
		cx.sync();   // (the header lives on this stack frame)
	}

	// VCF text: header + contig lines of the selected prefixes + records in the reference's chunk order.
	// only_refs = nullptr: the combined (stdout) stream; otherwise the records of those ref paths only
	std::string vcf_header(const std::vector<std::string> &for_prefixes, bool with_date) const {
		std::string o = pvh::vcf_header_common(with_date);
		for (auto &p : for_prefixes) for (u32 r : pvh::refs_in_sample(g, p)) o += pvh::vcf_contig_line(g.paths[r]);
		return o + pvh::vcf_col_header(gm);
	}
	// for the multi-GPU merge: every record as a text line plus (component id, ordinal of its region inside that
	// component, ref path id); lines are in (ref, region, context) order.  regions_per_tree: regions that reached
	// the overlay, per PVST tree of this session.
	u32 tree_of_flubble(u32 fl) const { return (u32)(std::upper_bound(forest.treeOff.begin(), forest.treeOff.end(), fl) - forest.treeOff.begin()) - 1; }
	std::vector<u64> regions_per_tree() const {
		std::vector<u64> n(forest.comp_id.size(), 0);
		for (size_t r = 0; r < rec.n_regions; r++) n[tree_of_flubble(rec.rov_fl[r])]++;
		return n;
	}
	// ---- VCF text on the device (pv_text.cuh).  The graph-dependent tables (segment ids, sequences, path names, the
	// genotype column layout per component) are uploaded once per loaded graph; forest + records travel per request.
	mutable pv::TextTables tt;
	mutable pv::TextOut text_buf;   // where the device text lands (page-locked; the allocation survives across requests)
	void build_text_tables() const {
		if (tt.ready) return;
		if (!g.has_labels) throw std::runtime_error("VCF text needs the segment sequences (session opened without labels)");
		if (!paths_up) throw std::runtime_error("VCF text needs the path index (run call first)");
		pv::Ctx &cx = *c;
		auto up32 = [&](pv::Buf<u32> &b, const std::vector<u32> &h) { b = pv::Buf<u32>(cx, h.size() ? h.size() : 1); cx.h2d(b.p, h.data(), h.size() * 4); };
		if (!tt_graph_dev) {   // a device-parsed graph already left its ids and sequences in HBM
			up32(tt.ids, g.ids);
			tt.lblOff = pv::Buf<u64>(cx, g.lblOff.size()); cx.h2d(tt.lblOff.p, g.lblOff.data(), g.lblOff.size() * 8);
			tt.lbl = pv::Buf<u8>(cx, g.lbl.size() ? g.lbl.size() : 1); cx.h2d(tt.lbl.p, g.lbl.data(), g.lbl.size());
		}
		std::vector<u32> tagOff(1, 0); std::string tags;
		for (auto &p : g.paths) { tags += p.tag; tagOff.push_back((u32)tags.size()); }
		up32(tt.tagOff, tagOff);
		tt.tags = pv::Buf<u8>(cx, tags.size() ? tags.size() : 1); cx.h2d(tt.tags.p, tags.data(), tags.size());
		up32(tt.ploidy, gm.ploidy);
		tt.nCols = (u32)gm.ploidy.size();
		// per component: (column, phase, local path) of its paths, sorted by (column, phase) -- the merge order of the GT writer
		std::vector<u32> ceOff(1, 0), ceCol, cePhase, ceLocal;
		for (size_t k = 0; k + 1 < dg.cpOff.size(); k++) {
			std::vector<std::array<u32, 3>> e;
			for (u32 j = dg.cpOff[k]; j < dg.cpOff[k + 1]; j++) { u32 h = dg.cpList[j]; if (gm.col[h] != pvh::NONE32) e.push_back({gm.col[h], gm.phase[h], j - dg.cpOff[k]}); }
			std::sort(e.begin(), e.end());
			for (auto &x : e) { ceCol.push_back(x[0]); cePhase.push_back(x[1]); ceLocal.push_back(x[2]); }
			ceOff.push_back((u32)ceCol.size());
		}
		up32(tt.ceOff, ceOff); up32(tt.ceCol, ceCol); up32(tt.cePhase, cePhase); up32(tt.ceLocal, ceLocal);
		cx.sync();   // the host staging vectors die here
		tt.ready = true;
	}
	// records of the last call as text, in the reference's order; only_refs = nullptr: every record
	void vcf_records_device(size_t chunk_size, const std::set<u32> *only_refs, pv::TextOut &out, int fd = -1, u64 *streamed = nullptr) const {
		out.n = rec.n(); out.text.clear(); out.lineOff.assign(1, 0);
		if (streamed) *streamed = 0;
		if (!rec.n()) return;
		if (!gm_ready) throw std::runtime_error("VCF text: no reference prefixes set");
		build_text_tables();
		pv::Ctx &cx = *c;
		pv::Trace tr(cx, "text");
		const size_t n = rec.n(), nFl = forest.n();
		auto upv = [&](auto &buf, const auto &h, size_t cnt) { typedef typename std::remove_reference<decltype(*buf.p)>::type T; buf = pv::Buf<T>(cx, cnt ? cnt : 1); cx.h2d(buf.p, h.data(), cnt * sizeof(T)); };
		pv::Buf<u32> fa, fz, fh, rovFl, rovComp, rovTang, cpOff;
		pv::Buf<u8> faf, fzf;
		upv(fa, forest.a_v, nFl); upv(fz, forest.z_v, nFl); upv(fh, forest.height, nFl); upv(faf, forest.a_fwd, nFl); upv(fzf, forest.z_fwd, nFl);
		upv(rovFl, rec.rov_fl, rec.rov_fl.size()); upv(rovComp, rec.rov_comp, rec.rov_comp.size()); upv(rovTang, rec.rov_tangled, rec.rov_tangled.size());
		upv(cpOff, rec.cpOff, rec.cpOff.size());
		const pv::RecPiece &R = rec.flat(cx);   // the records never left HBM
		pv::TextView t{tt.ids.p, dg.segLen.p, dg.steps.p, dg.stepOff.p, tt.lblOff.p, tt.lbl.p, tt.tagOff.p, tt.tags.p, tt.ploidy.p, tt.nCols,
			       tt.ceOff.p, tt.ceCol.p, tt.cePhase.p, tt.ceLocal.p, fa.p, fz.p, fh.p, faf.p, fzf.p, rovFl.p, rovComp.p, rovTang.p, cpOff.p,
			       R.rov.p, R.ref.p, R.cxtL.p, R.cxtR.p, R.vt.p, R.pos.p, R.refStart.p, R.refLen.p, R.grpOff.p, R.gtOff.p, R.gHap.p, R.gStart.p, R.gLen.p, R.gCnt.p, R.gt.p};
		std::vector<u8> keep;
		if (only_refs) { keep.assign(g.H(), 0); for (u32 r : *only_refs) if (r < g.H()) keep[r] = 1; }
		tr.mark("upload records");
		pv::records_text_device(cx, t, n, rec.n_regions, chunk_size, g.H(), only_refs ? keep.data() : nullptr, out, fd, streamed);
		tr.mark("text kernels + d2h");
	}
	// for the multi-GPU merge: every record as a text line plus (component id, ordinal of its region inside that
	// component, ref path id); lines are in (ref, region, context) order.
	void record_lines(pv::TextOut &out, std::vector<u32> &keys) const {
		std::vector<u64> per = regions_per_tree(), first(per.size() + 1, 0);
		for (size_t t = 0; t < per.size(); t++) first[t + 1] = first[t] + per[t];
		vcf_records_device(~(size_t)0, nullptr, out);
		rec.to_host(*c);
		keys.clear();
		for (size_t i : pvh::record_order(rec, ~(size_t)0)) {   // same keys, same stable order as the device permutation
			u32 r = rec.rec_rov[i];
			if (r >= rec.n_regions) { keys.push_back(pvh::NONE32); keys.push_back(0); keys.push_back(rec.rec_ref[i]); continue; }   // SUBR: final chunk
			u32 t = tree_of_flubble(rec.rov_fl[r]);
			keys.push_back(forest.comp_id[t]); keys.push_back((u32)(r - first[t])); keys.push_back(rec.rec_ref[i]);
		}
	}
};

} // namespace pve
