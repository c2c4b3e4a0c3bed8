/* TEST INFRASTRUCTURE ONLY.
 *
 * sb_oracle.c -- the parity oracle: a plain-C, single-threaded restatement of the
 * reference's CPU algorithm (sparcityeu/SparseBase v0.3.1) for the preprocessing hot path
 * (edge list -> COO, fused degree
 * features, ReorderHeatmap, BOBAReorder).  Built by oracle/Makefile into oracle/liboracle.so.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the reported CPU baseline.
 * Nothing under sparsebase_b200/ or include/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every entry point against
 *  (1) the golden vectors of the reference's own tests
 *      (tests/suites/sparsebase/converter/common.inc:5-16, format/common.inc:4-12,
 *       functionality_common.inc:6-56, format/csr_tests.cc:80-115, coo_tests.cc:77-115)
 *      transcribed in tests/golden/reference_vectors.json, and
 *  (2) outputs of the reference itself: oracle/_ref/libsbref.so (the unmodified reference
 *      compiled from /root/reference/src by `make -C oracle ref`) run on seeded random
 *      graphs here, and the committed fixtures under tests/golden/ generated from it by
 *      tests/golden/make_golden.py (DegreeReorder tie order and the exact RCM permutation
 *      are pinned by no reference test, only by executing the reference).
 *
 * Symbols: sbo_<op>_<IDType>_<NNZType>_<ValueType>; the per-type bodies are in
 * sb_oracle_impl.h, each citing the reference file:line it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define I int32_t
#define N int32_t
#define V float
#define F float
#define HAS_V 1
#define TAG i32_i32_f32
#include "sb_oracle_impl.h"
#undef I
#undef N
#undef V
#undef F
#undef HAS_V
#undef TAG

#define I int32_t
#define N int64_t
#define V float
#define F float
#define HAS_V 1
#define TAG i32_i64_f32
#include "sb_oracle_impl.h"
#undef I
#undef N
#undef V
#undef F
#undef HAS_V
#undef TAG

#define I int64_t
#define N int64_t
#define V double
#define F double
#define HAS_V 1
#define TAG i64_i64_f64
#include "sb_oracle_impl.h"
#undef I
#undef N
#undef V
#undef F
#undef HAS_V
#undef TAG

#define I int32_t
#define N int32_t
#define V int32_t
#define F float
#define HAS_V 1
#define TAG i32_i32_i32
#include "sb_oracle_impl.h"
#undef I
#undef N
#undef V
#undef F
#undef HAS_V
#undef TAG

#define I int32_t
#define N int32_t
#define V char
#define F float
#define HAS_V 0
#define TAG i32_i32_void
#include "sb_oracle_impl.h"
#undef I
#undef N
#undef V
#undef F
#undef HAS_V
#undef TAG

/* permute/permute_order_one.cc:17-37 -- inv[order[i]] = i (:27-29); out[i] = vals[inv[i]]
 * (:31-33). */
#define PERMUTE1D(TAG, IT, VT)                                                  \
  int sbo_permute1d_##TAG(int64_t len, void *vals_, void *order_, void *out_) { \
    const VT *vals = (const VT *)vals_;                                         \
    const IT *order = (const IT *)order_;                                       \
    VT *out = (VT *)out_;                                                       \
    IT *inv = (IT *)malloc((size_t)(len ? len : 1) * sizeof(IT));               \
    if (!inv) return 1;                                                         \
    for (int64_t i = 0; i < len; i++) inv[order[i]] = (IT)i;                    \
    for (int64_t i = 0; i < len; i++) out[i] = vals[inv[i]];                    \
    free(inv);                                                                  \
    return 0;                                                                   \
  }
PERMUTE1D(i32_f32, int32_t, float)
PERMUTE1D(i64_f64, int64_t, double)

/* bases/reorder_base.h:662-671 -- inv[perm[i]] = i */
#define INVPERM(TAG, IT)                                                    \
  int sbo_inverse_permutation_##TAG(int64_t len, void *perm_, void *out_) { \
    const IT *perm = (const IT *)perm_;                                     \
    IT *out = (IT *)out_;                                                   \
    for (int64_t i = 0; i < len; i++) out[perm[i]] = (IT)i;                 \
    return 0;                                                               \
  }
INVPERM(i32, int32_t)
INVPERM(i64, int64_t)

int sbo_abi_version(void) { return 1; }
