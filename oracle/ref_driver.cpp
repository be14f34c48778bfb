/* TEST INFRASTRUCTURE ONLY (oracle). Tiny argv driver that stands in for the
 * reference's Taywee/args CLI (app/cli/cli.cpp:97-262, absent third-party dep)
 * and calls the UNMODIFIED reference entry points do_gfa2vcf / do_decompose /
 * do_call (app/subcommand/{gfa2vcf,decompose,call}.cpp).  Built only into
 * oracle/_ref/povu_ref by oracle/Makefile. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <filesystem>
#include "povu/common/app.hpp"
#include "subcommand/call.hpp"
#include "subcommand/decompose.hpp"
#include "subcommand/gfa2vcf.hpp"
#include "mto/from_gfa.hpp"
#include "povu/graph/bidirected.hpp"

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: povu_ref {gfa2vcf|decompose|call} -i gfa [-o dir] [-f dir] [-P prefix]... [-r file] [-t N] [-c N] [-q N] [-v N] [-s] [-h] [--structure-export json]\n"); return 2; }
	core::config cfg;
	std::string task = argv[1];
	bool have_out = false;
	for (int i = 2; i < argc; i++) {
		std::string a = argv[i];
		auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
		if (a == "-i" || a == "--input-gfa") cfg.set_input_gfa(val());
		else if (a == "-o" || a == "--output-dir") { cfg.set_output_dir(val()); have_out = true; }
		else if (a == "-f" || a == "--forest-dir") cfg.set_forest_dir(val());
		else if (a == "-P" || a == "--path-prefix") { cfg.set_ref_input_format(core::input_format_e::params); cfg.add_ref_name_prefix(val()); }
		else if (a == "-r" || a == "--prefix-list") { cfg.set_ref_input_format(core::input_format_e::file_path); cfg.set_references_txt(val()); }
		else if (a == "-t" || a == "--threads") cfg.set_thread_count((uint8_t)atoi(val().c_str()));
		else if (a == "-c" || a == "--chunk-size") cfg.set_chunk_size((size_t)atol(val().c_str()));
		else if (a == "-q" || a == "--queue-length") cfg.set_queue_len((size_t)atol(val().c_str()));
		else if (a == "-v" || a == "--verbosity") cfg.set_verbosity((unsigned char)atoi(val().c_str()));
		else if (a == "-g" || a == "--restrict") cfg.set_genomic_region(val());
		else if (a == "-s" || a == "--subflubbles") cfg.set_subflubbles(true);
		else if (a == "-h" || a == "--hairpins") cfg.set_hairpins(true);
		else if (a == "--structure-export") cfg.set_structure_export_path(val());
		else if (a == "--stdout") { }
		else { cfg.set_ref_input_format(core::input_format_e::params); cfg.add_ref_name_prefix(a); }
	}
	/* bench.py --impl reference: `timed-gfa2vcf` runs the same two halves as do_gfa2vcf (gfa2vcf.cpp:49-70) with a
	 * clock around each, plus one stand-alone GFA load with the call configuration, so that the compute time of the
	 * call half can be reported without the text parse (which is not the reference's own code here: liteseq stand-in). */
	if (task == "timed-gfa2vcf") {
		auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
		cfg.set_inc_vtx_labels(true); cfg.set_inc_refs(true); cfg.set_stdout_vcf(true);
		char tmpl[] = "/tmp/povu_ref_timed_XXXXXX";
		char *d = mkdtemp(tmpl);
		if (!d) { perror("mkdtemp"); return 1; }
		core::config dc = cfg; dc.set_task(core::task_e::decompose); dc.set_inc_vtx_labels(false); dc.set_inc_refs(false); dc.set_output_dir(d);
		double t0 = now();
		povu::subcommands::decompose::do_decompose(dc);
		double t1 = now();
		{ povu::bidirected::VG *g = mto::from_gfa::to_bd(cfg); delete g; }
		double t2 = now();
		core::config cc = cfg; cc.set_task(core::task_e::call); cc.set_forest_dir(d);
		povu::subcommands::call::do_call(cc);
		double t3 = now();
		std::filesystem::remove_all(d);
		fprintf(stderr, "povu-ref-timing decompose_total_s=%.6f call_load_s=%.6f call_total_s=%.6f\n", t1 - t0, t2 - t1, t3 - t2);
		return 0;
	}
	if (task == "gfa2vcf") {
		cfg.set_task(core::task_e::gfa2vcf);
		cfg.set_inc_vtx_labels(true); cfg.set_inc_refs(true);
		if (!have_out) cfg.set_stdout_vcf(true);
		povu::subcommands::gfa2vcf::do_gfa2vcf(cfg);
	} else if (task == "decompose") {
		cfg.set_task(core::task_e::decompose);
		povu::subcommands::decompose::do_decompose(cfg);
	} else if (task == "call") {
		cfg.set_task(core::task_e::call);
		cfg.set_inc_vtx_labels(true); cfg.set_inc_refs(true);
		if (!have_out) cfg.set_stdout_vcf(true);
		povu::subcommands::call::do_call(cfg);
	} else { fprintf(stderr, "unknown task %s\n", task.c_str()); return 2; }
	return 0;
}
