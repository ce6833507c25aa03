/* Tuning / debug entry points of libtclight_tuning.so (csrc built with -DTCL_ATTN_TUNING: `make -C tclight_b200/csrc tuning`).
 * NOT part of the product ABI: libtclight.so ships one attention configuration per head-dim class and exports none of
 * these.  Used by tools/bench_attn_variants.py, tools/prof_one.py, tools/trace_attn.py and the variant sweep in
 * tests/test_attention_gpu.py. */
#ifndef TCLIGHT_TUNING_H
#define TCLIGHT_TUNING_H
#include "tclight.h"
#ifdef __cplusplus
extern "C" {
#endif
/* kernel variant used by tcl_attention (-1 = shipped configuration; see the dispatch in csrc/attn.cu) and trimming of the
 * MMA shapes to the live head-dim columns; both return the previous value */
int tcl_debug_attention_variant(int variant);
int tcl_debug_attention_trim(int on);
/* active only in -DTCL_ATTN_TRACE builds: device int64[192] event log (see attn.cu) */
void tcl_debug_attention_trace(long long* buf);
#ifdef __cplusplus
}
#endif
#endif
