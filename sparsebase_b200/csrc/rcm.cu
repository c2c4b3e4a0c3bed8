// rcm.cu -- placeholder until the order-exact BFS lands (next commit).
#include "common.cuh"
using namespace sb200;
extern "C" int sb200_rcm_reorder(int device, int64_t n, int64_t nnz, const void *row_ptr,
                                 const void *col, void *out_inv, int id_type, int nnz_type,
                                 void *stream) {
  set_error("sb200_rcm_reorder: not implemented yet");
  return SB200_ERR_INTERNAL;
}
