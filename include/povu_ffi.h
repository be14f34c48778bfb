/* povu_ffi.h -- the povu-ffi C API, served by the B200-native library (libpovu_b200.so).
 *
 * Declarations are the ones povu-rs binds with bindgen
 * (/root/reference/povu-rs/povu-ffi/povu_ffi.h:147-437, implementation povu_ffi.cpp:56-536): same names,
 * argument meaning, ownership and error behaviour, so `povu-rs/build.rs` can link this library in place of
 * `povu_ffi + mto + povulib + liteseq`.  What differs is that the entry points the reference
 * leaves as stubs are real here: povu_graph_add_path (reference: returns false, povu_ffi.cpp:172-177),
 * povu_graph_find_flubbles (reference: PVST with the dummy root only, :369-391), povu_flubbles_get (:405-412),
 * povu_flubbles_call_variants and povu_gfa_to_vcf (reference: error 1 "not yet implemented", :441-462,:505-527).
 *
 * Conventions (unchanged): opaque handles are freed by the matching *_free (NULL-safe); a PovuFlubbles keeps a
 * non-owning pointer to its PovuGraph; arrays / strings returned by the library go back through
 * povu_*_free / povu_string_free; failure = NULL / false / (size_t)-1 and, if `error` is not NULL, code 1 and a
 * message the caller releases with povu_error_free.  No exception crosses the boundary.  Distinct graphs may be
 * used from different threads; one handle is not synchronised.
 *
 * Device selection: environment variable POVU_B200_DEVICE (CUDA ordinal, default 0).  There is no CPU
 * fallback: the compute entry points fail with an error when no CUDA device is usable.
 */
#ifndef POVU_FFI_H
#define POVU_FFI_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct PovuGraph PovuGraph;
typedef struct PovuFlubbles PovuFlubbles;
typedef struct PovuPvstTree PovuPvstTree;
typedef struct PovuVcfOutput PovuVcfOutput;

/* povu_graph_add_edge maps FORWARD to the left end and REVERSE to the right end of BOTH endpoints
 * (povu_ffi.cpp:148-156); povu_graph_get_edges returns the ends the same way (:250-252). */
typedef enum { POVU_ORIENTATION_FORWARD = 0, POVU_ORIENTATION_REVERSE = 1 } PovuOrientation;

typedef struct { uint64_t id; const char *sequence; size_t sequence_len; } PovuVertex;
typedef struct { uint64_t from_id; PovuOrientation from_orientation; uint64_t to_id; PovuOrientation to_orientation; } PovuEdge;
typedef struct { uint64_t vertex_id; PovuOrientation orientation; } PovuStep;
typedef struct { const char *name; size_t name_len; PovuStep *steps; size_t steps_count; } PovuPath;
typedef struct {
	uint64_t id; const char *type_name; uint64_t start_vertex_id; uint64_t end_vertex_id;
	PovuStep **walks; size_t *walk_lengths; size_t walks_count;
} PovuFlubble;
typedef struct { int code; char *message; } PovuError;

PovuGraph *povu_graph_new(size_t vertex_capacity, size_t edge_capacity, size_t path_capacity);
PovuGraph *povu_graph_from_gfa(const char *gfa_path, PovuError *error);
void povu_graph_free(PovuGraph *graph);
size_t povu_graph_add_vertex(PovuGraph *graph, uint64_t id, const char *sequence);
size_t povu_graph_add_edge(PovuGraph *graph, uint64_t from_id, PovuOrientation from_orientation, uint64_t to_id, PovuOrientation to_orientation);
bool povu_graph_add_path(PovuGraph *graph, const char *name, const PovuStep *steps, size_t steps_count);
void povu_graph_finalize(PovuGraph *graph);
size_t povu_graph_vertex_count(const PovuGraph *graph);
size_t povu_graph_edge_count(const PovuGraph *graph);
size_t povu_graph_path_count(const PovuGraph *graph);
PovuVertex *povu_graph_get_vertices(const PovuGraph *graph, size_t *count);
PovuEdge *povu_graph_get_edges(const PovuGraph *graph, size_t *count);
PovuPath *povu_graph_get_paths(const PovuGraph *graph, size_t *count);
void povu_vertices_free(PovuVertex *vertices, size_t count);
void povu_edges_free(PovuEdge *edges, size_t count);
void povu_paths_free(PovuPath *paths, size_t count);
bool povu_graph_set_references_from_file(PovuGraph *graph, const char *ref_file, PovuError *error);
bool povu_graph_set_references_from_prefixes(PovuGraph *graph, const char **prefixes, size_t count, PovuError *error);
PovuFlubbles *povu_graph_find_flubbles(PovuGraph *graph, PovuError *error);
void povu_flubbles_free(PovuFlubbles *flubbles);
size_t povu_flubbles_count(const PovuFlubbles *flubbles);
PovuFlubble *povu_flubbles_get(const PovuFlubbles *flubbles, size_t index);
void povu_flubble_free(PovuFlubble *flubble);
PovuPvstTree *povu_flubbles_get_pvst_tree(const PovuFlubbles *flubbles);
void povu_pvst_tree_free(PovuPvstTree *tree);
size_t povu_pvst_tree_vertex_count(const PovuPvstTree *tree);
PovuVcfOutput *povu_flubbles_call_variants(PovuFlubbles *flubbles, PovuError *error);
bool povu_vcf_write_to_file(const PovuVcfOutput *vcf, const char *path, PovuError *error);
char *povu_vcf_to_string(const PovuVcfOutput *vcf, size_t *length);
void povu_vcf_free(PovuVcfOutput *vcf);
void povu_string_free(char *str);
bool povu_gfa_to_vcf(const char *gfa_path, const char *vcf_path, const char *ref_file, PovuError *error);
void povu_error_free(PovuError *error);

#ifdef __cplusplus
}
#endif
#endif
