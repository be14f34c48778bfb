/* TEST INFRASTRUCTURE ONLY (oracle) -- liteseq stand-in, see types.h */
#ifndef LITESEQ_STANDIN_REFS_H
#define LITESEQ_STANDIN_REFS_H
#include "types.h"
namespace liteseq {
struct ref_walk { id_t *v_ids; strand *strands; idx_t *loci; idx_t step_count; };
struct ref { ref_walk *walk; char *tag; char *sample_name; idx_t hap_id; idx_t hap_len; ref_id_type id_type; };
const char *get_tag(const ref *r);
const char *get_sample_name(const ref *r);
idx_t get_hap_id(const ref *r);
idx_t get_hap_len(const ref *r);
idx_t get_step_count(const ref *r);
ref_id_type get_ref_id_type(const ref *r);
const id_t *get_walk_v_ids(const ref *r);
const strand *get_walk_strands(const ref *r);
}
#endif
