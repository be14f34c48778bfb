/* TEST INFRASTRUCTURE ONLY (oracle) -- liteseq stand-in, see types.h */
#ifndef LITESEQ_STANDIN_GFA_H
#define LITESEQ_STANDIN_GFA_H
#include "types.h"
#include "refs.h"
namespace liteseq {
struct vtx { id_t id; char *seq; };
struct edge { id_t v1_id; vtx_side_e v1_side; id_t v2_id; vtx_side_e v2_side; };
struct gfa_config { const char *fp; bool inc_vtx_labels; bool inc_refs; };
struct gfa_config_cpp : gfa_config {
	gfa_config_cpp(const char *fp_, bool l, bool r) { fp = fp_; inc_vtx_labels = l; inc_refs = r; }
};
struct gfa_props {
	idx_t vtx_arr_size, l_line_count, ref_count;
	vtx **v; edge *e; ref **refs; int status;
};
gfa_props *gfa_new(gfa_config *conf);
void gfa_free(gfa_props *g);
vtx *get_vtx(gfa_props *g, idx_t i);
ref *get_ref(gfa_props *g, idx_t i);
}
#endif
