/* TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for the third-party
 * `liteseq` GFA parser (github.com/urbanslug/liteseq @ 3678c8b7, pinned in the
 * reference's cmake/deps.cmake:26-33; its source is NOT under /root/reference).
 * Declares only the API surface the reference's call sites use
 * (SURVEY.md Appendix A).  Not product code: never linked into povu_b200. */
#ifndef LITESEQ_STANDIN_TYPES_H
#define LITESEQ_STANDIN_TYPES_H
#include <stdint.h>
namespace liteseq {
typedef uint32_t id_t;
typedef uint32_t idx_t;
enum strand { STRAND_FWD = 0, STRAND_REV = 1 };
enum vtx_side_e { LEFT = 0, RIGHT = 1 };
enum ref_id_type { REF_ID_PANSN = 0, REF_ID_RAW = 1 };
}
#endif
