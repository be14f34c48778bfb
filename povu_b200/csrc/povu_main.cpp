// povu -- command-line driver of povu_b200 with the reference CLI's contract for the hot path:
//   povu gfa2vcf  -i g.gfa (-P prefix... | -r file | refs...) [-o dir | --stdout] [-c N] [-q N] [-h] [-s]
//   povu decompose -i g.gfa [-o dir] [-h] [-s]
//   povu call      -i g.gfa [-f forest_dir] (-P prefix... | -r file | refs...) [-o dir | --stdout] [-c N] [-q N] [-g ref:a-b]
//   global: --version, -v/--verbosity N, -t/--threads N, --help
// Mirrors /root/reference/app/cli/cli.cpp:14-382 (flags, required -i, exactly one reference source),
// app/main.cpp:12-44 (a parse error prints the message + usage to stderr and the process then exits 0 without
// doing anything) and app/subcommand/{gfa2vcf,decompose,call}.cpp (outputs, error text, exit codes).
// Host code only: all graph work goes through the C ABI of libpovu_b200.so (include/povu_b200.h).
// -h/--hairpins prints the reference's "Boundary: b1 b2" report to stderr; --structure-export writes the conformance JSON and
// its <json>.flubble-debug.jsonl side file (to_structure_export.cpp, flubbles.cpp:108-231).
// Not provided (outside the hot path): info, prune, vcf, -s/--subflubbles -- they fail loudly instead
// of silently doing something else.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>
#include "../../include/povu_b200.h"

static const char *VERSION = "0.0.1-alpha";   // app/cli/cli.hpp:10

struct Config {
	std::string task, gfa, out_dir = ".", forest_dir = ".", prefix_file, restrict_region, structure_export;
	std::vector<std::string> prefixes_flag, prefixes_pos;
	bool have_out = false, stdout_vcf = false, hairpins = false, subflubbles = false;
	size_t chunk = 100, queue = 4;
	int verbosity = 0, threads = 1;
};

static void usage(FILE *f) {
	fprintf(f, "  povu {OPTIONS} [COMMAND]\n\n    Explore variation in a variation graph\n\n  OPTIONS:\n\n      commands\n"
		   "        gfa2vcf                           Convert GFA to VCF (decompose + call)\n"
		   "        decompose                         Find regions of variation\n"
		   "        call                              Generate a VCF from regions of variation\n"
		   "      arguments\n        --version                         The current version of povu\n"
		   "        -v[verbosity], --verbosity=[verbosity]   Level of output [default: 0]\n"
		   "        -t[threads], --threads=[threads]  Number of threads to use [default: 1]\n"
		   "        -h, --help                        help\n\n");
}

struct ParseError { std::string msg; };

static int cli(int argc, char **argv, Config &c) {
	bool version = false, help = false;
	try {
		std::vector<std::string> a(argv + 1, argv + argc);
		// split --flag=value
		for (size_t i = 0; i < a.size(); i++) if (a[i].rfind("--", 0) == 0) { size_t eq = a[i].find('='); if (eq != std::string::npos) { std::string v = a[i].substr(eq + 1); a[i] = a[i].substr(0, eq); a.insert(a.begin() + (long)i + 1, v); } }
		size_t i = 0;
		auto val = [&](const std::string &flag) -> std::string { if (i + 1 >= a.size()) throw ParseError{"Flag '" + flag + "' requires an argument but received none"}; return a[++i]; };
		auto num = [&](const std::string &flag) -> long { std::string v = val(flag); char *e = nullptr; long x = strtol(v.c_str(), &e, 10); if (e == v.c_str() || *e) throw ParseError{"Argument '" + flag + "' received invalid value type '" + v + "'"}; return x; };
		for (; i < a.size(); i++) {
			const std::string &x = a[i];
			if (x == "--version") version = true;
			else if (x == "--help") help = true;
			else if (x == "-v" || x == "--verbosity") c.verbosity = (int)num(x);
			else if (x == "-t" || x == "--threads") c.threads = (int)(unsigned char)num(x);      // set_thread_count(uint8_t), app.hpp:409
			else if (c.task.empty()) {
				if (x == "-h") help = true;
				else if (x == "gfa2vcf" || x == "decompose" || x == "call" || x == "info" || x == "prune" || x == "vcf") c.task = x;
				else throw ParseError{x.rfind("-", 0) == 0 ? "Flag could not be matched: " + x : "Unknown command: " + x};
			}
			else if (x == "-i" || x == "--input-gfa") c.gfa = val(x);
			else if (x == "-o" || x == "--output-dir") { c.out_dir = val(x); c.have_out = true; }
			else if ((x == "-h" || x == "--hairpins") && c.task != "call") c.hairpins = true;
			else if ((x == "-s" || x == "--subflubbles") && c.task != "call") c.subflubbles = true;
			else if (c.task == "decompose") throw ParseError{x.rfind("-", 0) == 0 ? "Flag could not be matched: " + x : "Passed in argument, but no positional arguments were ready to receive it: " + x};
			else if ((x == "-f" || x == "--forest-dir") && c.task == "call") c.forest_dir = val(x);
			else if ((x == "-g" || x == "--restrict") && c.task == "call") c.restrict_region = val(x);
			else if (x == "-c" || x == "--chunk-size") c.chunk = (size_t)num(x);
			else if (x == "-q" || x == "--queue-length") c.queue = (size_t)num(x);
			else if (x == "--stdout") c.stdout_vcf = true;
			else if (x == "--structure-export") c.structure_export = val(x);
			else if (x == "-r" || x == "--prefix-list") c.prefix_file = val(x);
			else if (x == "-P" || x == "--path-prefix") c.prefixes_flag.push_back(val(x));
			else if (x.rfind("-", 0) == 0 && x.size() > 1) throw ParseError{"Flag could not be matched: " + x};
			else c.prefixes_pos.push_back(x);
		}
		if (help) { usage(stdout); c.task.clear(); return 0; }
		if (!c.task.empty()) {
			if (c.gfa.empty()) throw ParseError{"Flag 'gfa' is required"};
			if (c.task == "gfa2vcf" || c.task == "call") {
				int n = (!c.prefix_file.empty()) + (!c.prefixes_flag.empty()) + (!c.prefixes_pos.empty());
				if (n != 1) throw ParseError{"Group validation failed somewhere!"};   // reference source: exactly one (cli.cpp:59-75)
			}
		}
	} catch (ParseError &e) {
		if (!version) { fprintf(stderr, "%s\n", e.msg.c_str()); usage(stderr); c.task.clear(); return 1; }
	}
	if (version) { printf("%s\n", VERSION); exit(EXIT_SUCCESS); }
	return 0;
}

struct Fail { std::string what; };
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::vector<std::string> prefixes_of(const Config &c) {
	if (!c.prefix_file.empty()) {
		std::ifstream in(c.prefix_file);
		if (!in) throw Fail{"could not open " + c.prefix_file};
		std::vector<std::string> out; std::string l;
		while (std::getline(in, l)) { if (!l.empty() && l.back() == '\r') l.pop_back(); if (!l.empty()) out.push_back(l); }
		return out;
	}
	return c.prefixes_flag.empty() ? c.prefixes_pos : c.prefixes_flag;
}

struct Dev {
	pvgpu_ctx *ctx = nullptr; pvgpu_session *s = nullptr; double t_create_ms = 0;
	~Dev() { if (s) pvgpu_session_close(s); if (ctx) pvgpu_destroy(ctx); }
	void open(const Config &c, bool labels, bool refs) {
		int dev = 0;
		if (const char *e = getenv("POVU_B200_DEVICE")) dev = atoi(e);
		double t0 = now_ms();
		if (pvgpu_create(dev, nullptr, &ctx) != 0) throw Fail{"no usable CUDA device (povu_b200 has no CPU fallback)"};
		t_create_ms = now_ms() - t0;
		if (pvgpu_session_open(ctx, c.gfa.c_str(), labels, refs, &s) != 0) throw Fail{pvgpu_last_error(ctx)};
	}
	void check(int rc) { if (rc != 0) throw Fail{pvgpu_last_error(ctx)}; }
};

static void unsupported(const Config &c) {
	if (c.subflubbles) throw Fail{"-s/--subflubbles is not provided by povu_b200 (sub-flubble passes are outside the accelerated hot path)"};
}

// text returned by the library (malloc'ed) -> std::string
static std::string take(Dev &d, char *p, size_t n) { if (!p) throw Fail{pvgpu_last_error(d.ctx)}; std::string s(p, n); pvgpu_free(p); return s; }
static std::string sidecar_path(const Config &c) { return c.structure_export + ".flubble-debug.jsonl"; }   // flubbles.cpp:217-220
static std::string read_file(const std::string &fp) { std::ifstream in(fp, std::ios::binary); return std::string((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); }
static std::string base_name(const std::string &fp) { size_t k = fp.find_last_of('/'); return k == std::string::npos ? fp : fp.substr(k + 1); }

// after decompose: the hairpin report (stderr) and the debug frames of the structure export (appended to the side file)
static void debug_outputs(const Config &c, Dev &d) {
	if (c.hairpins) { size_t n = 0; char *p = pvgpu_session_hairpin_text(d.s, &n); std::string t = take(d, p, n); fwrite(t.data(), 1, t.size(), stderr); }
	if (!c.structure_export.empty()) {
		size_t n = 0; char *p = pvgpu_session_flubble_debug_frames(d.s, &n); std::string t = take(d, p, n);
		if (t.empty()) return;      // (no component reached find_flubbles: the reference never opens the file)
		std::ofstream out(sidecar_path(c), std::ios::app | std::ios::binary);
		if (!out.is_open()) throw Fail{"could not open flubble debug sidecar: " + sidecar_path(c)};
		out << t;
	}
}
static void structure_export(const Config &c, Dev &d) {
	if (c.structure_export.empty()) return;
	std::string frames = read_file(sidecar_path(c));
	d.check(pvgpu_session_structure_export(d.s, c.structure_export.c_str(), base_name(c.gfa).c_str(), c.chunk, frames.c_str()));
}

static void do_decompose(const Config &c, Dev &d, const std::string &out_dir) {
	pvgpu_session_set_debug_numbering(d.s, c.hairpins);
	d.check(pvgpu_session_decompose(d.s));
	debug_outputs(c, d);
	if (c.verbosity > 1) fprintf(stderr, "INFO [povu::decompose] Found %u components with at least 3 vertices\n", pvgpu_session_tree_count(d.s));
	d.check(pvgpu_session_write_pvst(d.s, out_dir.c_str()));
}

static void write_vcfs(const Config &c, Dev &d, const std::vector<std::string> &prefixes) {
	std::vector<const char *> pp;
	for (auto &p : prefixes) pp.push_back(p.c_str());
	bool region_bad = false;   // the reference has printed the VCF header(s) by the time its producer thread rejects the region (genomics.cpp:76-85)
	if (!c.restrict_region.empty() && pvgpu_session_set_restrict(d.s, c.restrict_region.c_str())) {
		fprintf(stderr, "ERR [povu::rov::parse_genomic_region] %s\nERR [povu::genomics::gen_vcf_rec_map] Failed to parse genomic region\n", pvgpu_last_error(d.ctx));
		region_bad = true;
	}
	if (region_bad) d.check(pvgpu_session_set_prefixes(d.s, pp.data(), pp.size()));
	else {
		d.check(pvgpu_session_call(d.s, pp.data(), pp.size()));
		uint64_t st[12];
		pvgpu_session_call_stats(d.s, st);
		if (st[5]) throw Fail{"call: the reference aborts on this input (missing context boundary in a haplotype lap)"};
		if (st[4]) fprintf(stderr, "WARN [povu_b200::call] %llu regions have no vertex order (BFS sort and haplotype-lap order both empty): skipped\n", (unsigned long long)st[4]);
		if (st[10]) fprintf(stderr, "WARN [povu_b200::call] %llu SUBR records had a merged length past the end of a walk (undefined in the reference): clamped\n", (unsigned long long)st[10]);
	}
	const bool to_stdout = !c.have_out || c.stdout_vcf;
	if (to_stdout) {
		fflush(stdout);
		d.check(pvgpu_session_vcf_write(d.s, c.chunk, 1, nullptr, STDOUT_FILENO, nullptr));
		if (region_bad) exit(EXIT_FAILURE);
		structure_export(c, d);
		return;
	}
	mkdir(c.out_dir.c_str(), 0777);   // create_dir_if_not_exists (to_vcf.hpp:83-101)
	for (auto &p : prefixes) {
		std::string fp = c.out_dir + "/" + p + ".vcf";
		int fd = open(fp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
		if (fd < 0) throw Fail{"Append open failed: " + fp};
		int rc = pvgpu_session_vcf_write(d.s, c.chunk, 1, p.c_str(), fd, nullptr);
		close(fd);
		d.check(rc);
	}
	if (region_bad) exit(EXIT_FAILURE);
	structure_export(c, d);
}

int main(int argc, char **argv) {
	const double t_main = now_ms();
	struct AtExit { double t0; ~AtExit() { if (getenv("POVU_B200_TIMING")) fprintf(stderr, "[povu_b200] main() %.1f ms\n", now_ms() - t0); } } at_exit{t_main};
	Config c;
	cli(argc, argv, c);
	if (c.task.empty()) return 0;               // app/main.cpp:15 ignores cli()'s return value
	const std::string fn = "[povu::subcommands::do_" + c.task + "]";
	try {
		if (c.task == "info" || c.task == "prune" || c.task == "vcf") throw Fail{"subcommand '" + c.task + "' is not provided by povu_b200 (only the decompose -> call hot path is)"};
		unsupported(c);
		if (c.task == "decompose") {
			Dev d; d.open(c, false, false);
			do_decompose(c, d, c.out_dir);
		} else if (c.task == "call") {
			Dev d; d.open(c, true, true);
			if (pvgpu_session_read_pvst(d.s, c.forest_dir.c_str()) != 0) { fprintf(stderr, "ERROR [povu::subcommands::call] %s\n", pvgpu_last_error(d.ctx)); return EXIT_FAILURE; }
			write_vcfs(c, d, prefixes_of(c));
		} else {
			// gfa2vcf: the reference decomposes into a temporary forest directory and calls from the .pvst files it reads
			// back (gfa2vcf.cpp:25-70).  Here the flubble forest goes from decompose to call inside HBM: no text round trip.
			try {
				double t0 = now_ms();
				Dev d; d.open(c, true, true);
				double t1 = now_ms();
				if (!c.structure_export.empty()) remove(sidecar_path(c).c_str());   // reset_debug_sidecar (gfa2vcf.cpp:46)
				pvgpu_session_set_debug_numbering(d.s, c.hairpins || !c.structure_export.empty());
				d.check(pvgpu_session_decompose(d.s));
				debug_outputs(c, d);
				if (c.verbosity > 1) fprintf(stderr, "INFO [povu::decompose] Found %u components with at least 3 vertices\n", pvgpu_session_tree_count(d.s));
				double t2 = now_ms();
				write_vcfs(c, d, prefixes_of(c));
				double t3 = now_ms();
				if (getenv("POVU_B200_TIMING")) {
					uint64_t nb = 0; double ms[2] = {0, 0};
					pvgpu_session_gfa_stats(d.s, &nb, ms);
					fprintf(stderr, "[povu_b200] gfa2vcf: CUDA context %.1f ms, GFA open %.1f ms (file -> HBM %.1f, device parse %.1f; %llu bytes), decompose %.1f ms, call + VCF text %.1f ms\n",
						d.t_create_ms, t1 - t0 - d.t_create_ms, ms[0], ms[1], (unsigned long long)nb, t2 - t1, t3 - t2);
				}
			} catch (Fail &e) { fprintf(stderr, "%s Error: %s\n", fn.c_str(), e.what.c_str()); return EXIT_FAILURE; }
		}
	} catch (Fail &e) {
		fprintf(stderr, "%s Error: %s\n", fn.c_str(), e.what.c_str());
		return EXIT_FAILURE;
	}
	return 0;
}
