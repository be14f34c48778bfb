/* povu_b200 -- C ABI of the B200-native decompose -> call hot path.
 *
 * Plain pointers and sizes only (no torch / C++ types).  This is the device layer that sits behind
 * the reference's in-process operators
 *     pst::Tree::from_bd + pfl::find_flubbles   (/root/reference/include/povu/algorithms/flubbles.hpp:62,
 *                                                /root/reference/src/povu/graph/spanning_tree.cpp:262)
 *     bd::VG::componetize                       (/root/reference/src/povu/graph/bidirected.cpp:477)
 * and behind the reference's own FFI entry point
 *     povu_graph_find_flubbles                  (/root/reference/povu-rs/povu-ffi/povu_ffi.h:330-345)
 * (declared, with identical signatures, in include/povu_ffi.h and implemented on top of these calls).
 *
 * Graph encoding: vertices are indices 0..V-1 in ascending segment-id order (the order
 * mto::from_gfa::to_bd adds them, /root/reference/src/mto/from_gfa.cpp:191-215).  A "port" is
 * 2*vertex + side with side 0 = left, 1 = right.  Edge e joins ports ep1[e] and ep2[e]; edges are in
 * L-line order (from_gfa.cpp:223-243).  `L a + b +` is (2a+1, 2b); a '-' flips the respective side.
 *
 * Every function returns 0 on success, non-zero on failure (message via pvgpu_last_error).  There is no
 * CPU fallback: pvgpu_create fails when no CUDA device is usable.
 */
#ifndef POVU_B200_H
#define POVU_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvgpu_ctx pvgpu_ctx;
typedef struct pvgpu_flubbles pvgpu_flubbles;

/* device = CUDA ordinal; stream = an existing cudaStream_t (e.g. torch's current stream) or NULL to let
 * the library create its own non-blocking stream. */
int pvgpu_create(int device, void *stream, pvgpu_ctx **out);
void pvgpu_destroy(pvgpu_ctx *ctx);
const char *pvgpu_last_error(const pvgpu_ctx *ctx);
/* kernels launched through this context so far (bench.py: gpu_launches) */
uint64_t pvgpu_launch_count(const pvgpu_ctx *ctx);
uint64_t pvgpu_peak_device_bytes(const pvgpu_ctx *ctx);
/* times the host has waited for the context's stream so far (each wait drains the device queue; bench.py: host_syncs_per_step) */
uint64_t pvgpu_host_sync_count(const pvgpu_ctx *ctx);
/* per-kernel CUDA-event timing (bench.py roofline leg).  While on, every kernel launch is bracketed by
 * events on the context's stream; pvgpu_profile_json synchronises and writes a JSON array of
 * {name, launches, items, bytes (algorithmic), ms} (returns the size needed incl. NUL). */
void pvgpu_set_profiling(pvgpu_ctx *ctx, int on);
size_t pvgpu_profile_json(pvgpu_ctx *ctx, char *buf, size_t cap);
/* 1 = sm_100a CUDA build, 2 = host debug rehearsal (tests only; never shipped) */
int pvgpu_build_kind(void);

/* test hook (tests/test_prims_gpu.py): runs the device primitives (stable pair sort on `bits` key bits, the scans,
 * range-minimum / previous-smaller queries) on n seeded elements and counts disagreements with the host's std::
 * algorithms.  Not part of the reference-facing surface. */
int pvgpu_selftest_prims(pvgpu_ctx *ctx, uint64_t n, uint64_t seed, int bits, uint64_t *mismatches);
/* fired-arc fixpoint (decompose, D6) on a random class sequence with crossing arcs against the reference's stack machine run on
 * the host; *rounds = fixpoint rounds (-1: not settled within 64, nothing compared) */
int pvgpu_selftest_arcs(pvgpu_ctx *ctx, uint32_t n, uint32_t n_classes, uint64_t seed, uint64_t *mismatches, int *rounds);

/* test hook (tests/test_text_emu.py, test_text_gpu.py): formats AF = ac / an for every 0 <= ac <= an <= max_an (and some
 * integers) with the device text writer and counts disagreements with the C library's "%.1f" / "%u". */
int pvgpu_selftest_text(pvgpu_ctx *ctx, uint32_t max_an, uint64_t *mismatches);

/* decompose with HOST inputs (pinned or pageable): copies ep1/ep2 to the device on the context's
 * stream, runs the pipeline, copies the flubble table back.  flags bit0: also return the class stack. */
int pvgpu_decompose(pvgpu_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *ep1, const uint32_t *ep2,
		    uint32_t flags, pvgpu_flubbles **out);
/* same, inputs already resident in device memory (bench.py `value`) */
int pvgpu_decompose_dev(pvgpu_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *d_ep1,
			const uint32_t *d_ep2, uint32_t flags, pvgpu_flubbles **out);

/* result accessors.  Flubbles are in creation order, components concatenated in component order;
 * parent / PVST vertex indices are local to the component (0 = dummy root, flubble k of the component
 * is PVST vertex k+1), exactly the indices the reference writes into <component_id>.pvst. */
uint64_t pvgpu_flubbles_count(const pvgpu_flubbles *f);
uint32_t pvgpu_flubbles_component_count(const pvgpu_flubbles *f);
uint32_t pvgpu_flubbles_dfs_rounds(const pvgpu_flubbles *f);
uint32_t pvgpu_flubbles_hash_collisions(const pvgpu_flubbles *f);
/* rounds of the fired-arc fixpoint that decides which class re-encounters still find their previous occurrence on the reference's
 * stack (0 = every component laminar; more than 64 = the sequential replay ran instead) */
uint32_t pvgpu_flubbles_arc_rounds(const pvgpu_flubbles *f);
/* > 0: the tree has vertices with three or more children reaching above them, where the reference's (top bracket, list size) class
 * names can differ from cycle equivalence (its capping edge, flubbles.cpp:564-573); the classes were then derived the reference's way
 * (pv_refnum.cuh) instead of from the XOR labels.  The value is the number of such vertices. */
uint32_t pvgpu_flubbles_faithful_classes(const pvgpu_flubbles *f);
/* any pointer may be NULL.  a_v/z_v are vertex indices; a_fwd/z_fwd 1 = '>' 0 = '<'. */
void pvgpu_flubbles_copy(const pvgpu_flubbles *f, uint32_t *comp, uint32_t *a_v, uint8_t *a_fwd, uint32_t *z_v,
			 uint8_t *z_fwd, uint32_t *parent, uint32_t *depth);
/* per component (index k -> component id k+1): lowest vertex index, vertex count, flubble count,
 * has-tips flag, non-laminar flag */
void pvgpu_flubbles_copy_components(const pvgpu_flubbles *f, uint32_t *root_v, uint32_t *n_vertices,
				    uint32_t *n_flubbles, uint32_t *has_tips, uint32_t *nonlaminar);
/* class stack (only when flags bit0 was set): n_vertices entries */
uint64_t pvgpu_flubbles_stack_size(const pvgpu_flubbles *f);
void pvgpu_flubbles_copy_stack(const pvgpu_flubbles *f, uint32_t *v, uint32_t *class_id, uint32_t *fwd);
void pvgpu_flubbles_free(pvgpu_flubbles *f);

/* ---------------------------------------------------------------------------------------------------
 * Sessions: one loaded GFA on one device.  This is the layer behind the `povu` CLI clone
 * (decompose / call / gfa2vcf: /root/reference/app/subcommand/{decompose,call,gfa2vcf}.cpp) and behind
 * include/povu_ffi.h.  It replaces the reference's in-process operators
 *     mto::from_gfa::to_bd        (/root/reference/src/mto/from_gfa.cpp:141)     -> pvgpu_session_open
 *     do_decompose + write_pvst   (decompose.cpp:94-160, to_pvst.cpp:30)         -> pvgpu_session_decompose / _write_pvst
 *     read_pvst                   (from_pvst.cpp:162)                            -> pvgpu_session_read_pvst
 *     ig::gen_vcf_rec_map         (/root/reference/src/ita/genomics/genomics.cpp:69) -> pvgpu_session_call
 *     to_vcf::init_vcfs/write_vcfs (/root/reference/src/mto/to_vcf.cpp:54,121)   -> pvgpu_session_vcf_text
 * Errors: non-zero return, message via pvgpu_last_error(ctx) (e.g. "Invalid GFA '<path>': ...").
 */
typedef struct pvgpu_session pvgpu_session;
int pvgpu_session_open(pvgpu_ctx *ctx, const char *gfa_path, int inc_vtx_labels, int inc_refs, pvgpu_session **out);
/* pvgpu_session_open parses the GFA text ON THE DEVICE (the file is streamed through page-locked staging
 * buffers into HBM and split / validated / converted there; replaces the liteseq parse + the pre-scan of
 * /root/reference/src/mto/from_gfa.cpp:28-113,141-280, same "Invalid GFA '<path>': ..." messages incl. the line number).
 * Links, walks and sequences then live only in HBM until something on the host asks for them.
 * text_bytes = size of the file, ms[0] = file -> HBM, ms[1] = device parse (wall, of the last open). */
void pvgpu_session_gfa_stats(const pvgpu_session *s, uint64_t *text_bytes, double ms[2]);
/* test hook: the same graph through the host loader (pv_host.hpp), the checker of the device parser's tests */
int pvgpu_session_open_hostparse(pvgpu_ctx *ctx, const char *gfa_path, int inc_vtx_labels, int inc_refs, pvgpu_session **out);
/* the loaded graph as flat arrays (any pointer may be NULL): ids[V] ascending segment ids, seg_len[V], ep1/ep2[E] ports in
 * L-line order, step_off[H+1], steps[S] packed (vertex<<1)|strand in P-line order, hap_len[H] contig lengths,
 * label_off[V+1] into labels[pvgpu_session_label_bytes] (the concatenated sequences). */
int pvgpu_session_copy_graph(pvgpu_session *s, uint32_t *ids, uint32_t *seg_len, uint32_t *ep1, uint32_t *ep2, uint32_t *step_off,
			     uint32_t *steps, uint32_t *hap_len, uint64_t *label_off, char *labels);
uint64_t pvgpu_session_label_bytes(pvgpu_session *s);
const char *pvgpu_session_path_name(const pvgpu_session *s, uint32_t k);
/* same, from arrays already in memory (the builder path of povu_ffi.h: povu_graph_new / add_vertex / add_edge /
 * add_path, /root/reference/povu-rs/povu-ffi/povu_ffi.h:147-262).  ids ascending; ep1/ep2 ports as above;
 * labels optional (label_off[V+1] into labels) -- without them seg_len[V] gives the segment lengths and VCF text is
 * unavailable; path k has name path_names[k] and packed steps steps[step_off[k] .. step_off[k+1]) = (vertex<<1)|strand. */
int pvgpu_session_from_arrays(pvgpu_ctx *ctx, uint32_t n_vertices, const uint32_t *ids, const uint32_t *seg_len, const uint64_t *label_off,
			      const char *labels, uint64_t n_edges, const uint32_t *ep1, const uint32_t *ep2, uint32_t n_paths,
			      const char *const *path_names, const uint32_t *step_off, const uint32_t *steps, pvgpu_session **out);
void pvgpu_session_close(pvgpu_session *s);
/* the whole hot path once: decompose -> device index -> call against the prefixes.  keep_resident != 0: the raw
 * arrays (edge ports, steps, segment lengths) stay in HBM between calls; 0: they are copied from the host inside the
 * call.  ms[4] receives the wall time of {upload, decompose, index build, call}. */
int pvgpu_session_hot_path(pvgpu_session *s, const char *const *prefixes, size_t n_prefixes, int keep_resident, double ms[4]);
/* graph sizes: vertices, edges, paths, total path steps */
void pvgpu_session_sizes(const pvgpu_session *s, uint64_t *n_vertices, uint64_t *n_edges, uint64_t *n_paths, uint64_t *n_steps);
int pvgpu_session_decompose(pvgpu_session *s);
uint64_t pvgpu_session_flubble_count(const pvgpu_session *s);
uint32_t pvgpu_session_tree_count(const pvgpu_session *s);
/* component id of PVST tree k, and its .pvst v0.0.3 text (malloc'ed; release with pvgpu_free) */
uint32_t pvgpu_session_tree_component(const pvgpu_session *s, uint32_t k);
char *pvgpu_session_pvst_text(const pvgpu_session *s, uint32_t k, size_t *len);
int pvgpu_session_write_pvst(const pvgpu_session *s, const char *dir);
/* The reference's two debug surfaces.  With the debug numbering on, decompose also derives -- in parallel, from
 * the tree it already holds (pv_refnum.cuh) -- the reference's internal numbering: tree-edge indices / ids in creation order, the
 * new_class() ids of simple_cycle_equiv, capping back edges, top brackets.
 *   pvgpu_session_hairpin_text           `povu decompose -h`: the "Boundary: b1 b2" lines the reference prints to stderr per component
 *                                        (/root/reference/src/povu/algorithms/flubbles.cpp:503-719), component order
 *   pvgpu_session_flubble_debug_frames   one "povu.flubble-debug.frame.v1" JSON line per component (flubbles.cpp:108-213): the content of
 *                                        the side file <json>.flubble-debug.jsonl that `gfa2vcf --structure-export json` leaves behind
 *   pvgpu_session_structure_export       after call: writes the "povu.lean4.structure.v1" document of `--structure-export`
 *                                        (/root/reference/src/mto/to_structure_export.cpp:148-640); frames_jsonl = the side file's content
 * Texts are malloc'ed (pvgpu_free). */
int pvgpu_session_set_debug_numbering(pvgpu_session *s, int on);
char *pvgpu_session_hairpin_text(const pvgpu_session *s, size_t *len);
char *pvgpu_session_flubble_debug_frames(const pvgpu_session *s, size_t *len);
int pvgpu_session_structure_export(pvgpu_session *s, const char *json_path, const char *input_name, size_t chunk_size, const char *frames_jsonl);
int pvgpu_session_read_pvst(pvgpu_session *s, const char *dir);
/* call against the paths selected by the prefixes (refs.hpp:278-301).  Records stay in the session. */
int pvgpu_session_call(pvgpu_session *s, const char *const *prefixes, size_t n_prefixes);
/* selects the reference paths without calling (records of an earlier call are dropped): the VCF header of those prefixes
 * can then be written on its own (what the reference has printed when its producer thread fails, to_vcf.cpp:54-82) */
int pvgpu_session_set_prefixes(pvgpu_session *s, const char *const *prefixes, size_t n_prefixes);
/* -g/--restrict "ref:start-end" of `povu call` (/root/reference/src/ita/variation/rov.cpp:51-183, genomics.cpp:76-88): the next
 * calls keep only regions with a boundary vertex whose step on path `ref` starts inside [start, end).  NULL clears it.
 * A malformed string fails with the reference's message (the reference exits with EXIT_FAILURE there). */
int pvgpu_session_set_restrict(pvgpu_session *s, const char *region);
uint64_t pvgpu_session_record_count(const pvgpu_session *s);
/* counters of the last call: [0] regions selected, [1] regions with >= 3 vertices, [2] tangled regions,
 * [3] inversion pin pairs, [4] regions whose BFS/Kahn order failed, [5] hard errors (the reference would
 * have exited), [6] device chunks, [7] regions dropped as too small, [8] vertices of the largest region, [9] regions ordered
 * by a global-arena BFS tier, [10] SUBR records whose merged length ran past the end of a walk (undefined in the
 * reference; clamped here), [11] slice trees of the SNE pass replayed sequentially (0 = all took the sorted fast path) */
void pvgpu_session_call_stats(const pvgpu_session *s, uint64_t out[12]);
/* combined VCF text (what `povu gfa2vcf` / `povu call --stdout` print); with_date = 0 drops ##fileDate.
 * prefix = NULL: combined stream; otherwise the per-prefix file content of `-o dir` (to_vcf.hpp:83-101). */
char *pvgpu_session_vcf_text(const pvgpu_session *s, size_t chunk_size, int with_date, const char *prefix, size_t *len);
/* same text written straight to a file descriptor from the page-locked buffer the device text landed in (what the
 * CLI uses for stdout / <prefix>.vcf; replaces to_vcf::write_vcfs, /root/reference/src/mto/to_vcf.cpp:121-139).
 * The record lines themselves are produced on the device (pv_text.cuh); only the '#' header is formatted on the host. */
int pvgpu_session_vcf_write(const pvgpu_session *s, size_t chunk_size, int with_date, const char *prefix, int fd, uint64_t *bytes_written);
/* ---- multi-GPU (one process per GPU): keep only this rank's share of the weakly connected components, balanced by
 * edges + path steps (greedy LPT).  Component ids, path ids and vertex / edge order stay those of the whole graph, so
 * the per-rank outputs concatenate into exactly the single-GPU output.  No exchange happens while computing; the only
 * communication is the final gather of record lines (povu_b200/dist.py does it with torch.distributed). */
int pvgpu_session_shard(pvgpu_session *s, uint32_t rank, uint32_t world, uint32_t *n_components_total);
/* The same, repeatable: the first call sets the whole graph aside (host arrays + a copy in HBM); every call cuts this rank's share
 * out of it ON THE DEVICE (component labels, per-component weights, compaction of segments / links / walks; only the handful of
 * component weights visit the host for the greedy balance).  from_host != 0 drops the device copy of the whole graph first, so
 * that the call also pays its host -> HBM transfer from the page-locked arrays (the end-to-end step of bench.py).  ms = wall time. */
/* The binary records of the last call as one self-describing stream written into caller-provided DEVICE memory (layout:
 * pv_engine.hpp "records_pack"; context steps carry segment ids, so streams of different shards can be concatenated).  This is
 * what a rank hands to the final NCCL gather of a multi-GPU run. */
uint64_t pvgpu_session_records_packed_bytes(const pvgpu_session *s);
int pvgpu_session_records_pack(pvgpu_session *s, void *dst_device, uint64_t capacity, uint64_t *bytes);
int pvgpu_session_reshard(pvgpu_session *s, uint32_t rank, uint32_t world, int from_host, uint32_t *n_components_total, double *ms);
/* after call: per PVST tree of this session its component id and the number of regions that reached the overlay */
void pvgpu_session_tree_regions(const pvgpu_session *s, uint32_t *comp_id, uint64_t *n_regions);
/* after call: all records as VCF text lines (malloc'ed, pvgpu_free) with line_off[n+1] and keys[3n] =
 * (component id, ordinal of the region inside its component, ref path id) per line */
char *pvgpu_session_record_lines(const pvgpu_session *s, size_t *len, uint64_t **line_off, uint32_t **keys, size_t *n);
void pvgpu_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
