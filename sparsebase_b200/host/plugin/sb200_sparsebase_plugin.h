// sb200_sparsebase_plugin.h -- binds libsb200.so into an UNMODIFIED SparseBase build.
//
// Include this header after the SparseBase headers in a translation unit compiled with
// SparseBase's CUDA support (USE_CUDA: format::CUDACSR, format::CUDAArray and
// context::CUDAContext exist) and link against libsb200.so.  It adds, through SparseBase's own
// extension points and without touching its sources:
//
//   sb200_plugin::RegisterConversions(converter)          Converter::RegisterConversionFunction
//       COO  -> CUDACSR   H2D of the (constructed, sorted) COO + sb200_coo_to_csr
//       CUDACSR -> CSC    sb200_csr_to_csc on the device + D2H       (to_context = CPU)
//       CUDACSR -> COO    sb200_csr_to_coo on the device + D2H       (to_context = CPU)
//   sb200_plugin::RegisterTransfers(converter)            replaces SparseBase's own blocking,
//       pageable cudaMemcpy functions for CSR <-> CUDACSR and CUDACSR -> CUDACSR (peer) by staged
//       transfers (two pinned buffers filled / drained by all host threads while the previous
//       chunk is on the bus); they also keep the column count m and size row_ptr with n + 1
//       entries (converter_order_two_cuda.cu:37-38, :55-57 lose m and under-allocate).
//   sb200_plugin::Register(DegreeReorder&) / (RCMReorder&) / (PermuteOrderTwo&) /
//                 (PermuteOrderOne&) / (DegreeDistribution&) / (Degrees&)
//     FunctionMatcherMixin::RegisterFunction({CUDACSR id} or {CUDAArray id}, fn) -- the pattern
//     of feature/jaccard_weights.cc:17-36.  A call such as
//         reorder.GetReorder(csr, {&gpu0}, /*convert_input=*/true)
//     then converts CSR -> CUDACSR (SparseBase's function) and runs the sb200 kernel; results
//     follow SparseBase's conventions (host IDType* from new[], inv[old] = new; permuters return
//     a CUDACSR / CUDAArray owning cudaMalloc'ed arrays, released by utils::CUDADeleter).
//
// INTEGRATION.md shows the ten-line change a maintainer would make to register these by
// default (inside ConverterOrderTwo::ResetConverterOrderTwo and the operators' constructors).
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../../include/sb200.h"
#include "sparsebase/context/cuda_context_cuda.cuh"
#include "sparsebase/converter/converter_order_two.h"
#include "sparsebase/feature/degree_distribution.h"
#include "sparsebase/feature/degrees.h"
#include "sparsebase/format/array.h"
#include "sparsebase/format/coo.h"
#include "sparsebase/format/csc.h"
#include "sparsebase/format/csr.h"
#include "sparsebase/format/cuda_array_cuda.cuh"
#include "sparsebase/format/cuda_csr_cuda.cuh"
#include "sparsebase/permute/permute_order_one.h"
#include "sparsebase/permute/permute_order_two.h"
#include "sparsebase/reorder/boba_reorder.h"
#include "sparsebase/reorder/degree_reorder.h"
#include "sparsebase/reorder/rcm_reorder.h"
#include "sparsebase/reorder/reorder_heatmap.h"

namespace sb200_plugin {
namespace sbase = ::sparsebase;

class Error : public sbase::utils::Exception {
 public:
  Error(int code, std::string msg) : msg_("libsb200 error " + std::to_string(code) + ": " + msg) {}
  const char *what() const noexcept override { return msg_.c_str(); }

 private:
  std::string msg_;
};

inline void check(int rc, int device) {
  if (rc == SB200_OK) return;
  if (rc == SB200_ERR_BAD_DEVICE) {
    int cnt = 0;
    sb200_device_count(&cnt);
    throw sbase::utils::CUDADeviceException(cnt, device);
  }
  if (rc == SB200_ERR_ALLOC) throw sbase::utils::AllocationException();
  throw Error(rc, sb200_last_error());
}

template <typename T>
constexpr int dtype_of() {
  if constexpr (std::is_void_v<T>) {
    return SB200_VOID;
  } else {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8,
                  "libsb200 moves 4- and 8-byte index / value types only");
    if constexpr (std::is_same_v<T, float>)
      return SB200_F32;
    else if constexpr (std::is_same_v<T, double>)
      return SB200_F64;
    else if constexpr (sizeof(T) == 4)
      return std::is_signed_v<T> ? SB200_I32 : SB200_U32;
    else
      return std::is_signed_v<T> ? SB200_I64 : SB200_U64;
  }
}

// ------------------------------------------------------------------ staged transfers
// SparseBase's host arrays are pageable (new[]).  A plain cudaMemcpy of pageable memory runs at
// a fraction of the PCIe rate; here every transfer goes through two pinned buffers: all host
// threads copy chunk k+1 between the user's array and a pinned buffer while chunk k is on the
// bus.  One Stager per device and thread; sync() once after the last array of a call.
class Stager {
 public:
  static Stager &get(int dev) {
    static thread_local Stager *per_dev[64] = {nullptr};
    if (dev < 0 || dev >= 64) throw Error(SB200_ERR_BAD_DEVICE, "device out of range");
    if (!per_dev[dev]) per_dev[dev] = new Stager(dev);
    return *per_dev[dev];
  }
  void h2d(void *dst, const void *src, size_t bytes) {
    select();
    for (size_t off = 0; off < bytes; off += kChunk) {
      const size_t len = bytes - off < kChunk ? bytes - off : kChunk;
      const int k = turn_++ & 1;
      cuda(cudaEventSynchronize(ev_[k]));
      host_copy(buf_[k], (const char *)src + off, len);
      cuda(cudaMemcpyAsync((char *)dst + off, buf_[k], len, cudaMemcpyHostToDevice, st_));
      cuda(cudaEventRecord(ev_[k], st_));
    }
  }
  void d2h(void *dst, const void *src, size_t bytes) {
    select();
    size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};
    for (size_t off = 0; off < bytes; off += kChunk) {
      const size_t len = bytes - off < kChunk ? bytes - off : kChunk;
      const int k = turn_++ & 1;
      drain(dst, k, pend_off, pend_len);  // the buffer's previous chunk goes to the user first
      cuda(cudaMemcpyAsync(buf_[k], (const char *)src + off, len, cudaMemcpyDeviceToHost, st_));
      cuda(cudaEventRecord(ev_[k], st_));
      pend_off[k] = off;
      pend_len[k] = len;
    }
    drain(dst, turn_ & 1, pend_off, pend_len);
    drain(dst, (turn_ + 1) & 1, pend_off, pend_len);
  }
  void sync() {
    select();
    cuda(cudaStreamSynchronize(st_));
  }
  void *stream() const { return st_; }

 private:
  static constexpr size_t kChunk = 16u << 20;
  explicit Stager(int dev) : dev_(dev) {
    select();
    cuda(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      cuda(cudaMallocHost(&buf_[k], kChunk));
      cuda(cudaEventCreateWithFlags(&ev_[k], cudaEventDisableTiming));
    }
  }
  void select() { cuda(cudaSetDevice(dev_)); }
  void cuda(cudaError_t e) {
    if (e != cudaSuccess) throw Error(SB200_ERR_CUDA, cudaGetErrorString(e));
  }
  void drain(void *dst, int k, size_t *off, size_t *len) {
    if (!len[k]) return;
    cuda(cudaEventSynchronize(ev_[k]));
    host_copy((char *)dst + off[k], buf_[k], len[k]);
    len[k] = 0;
  }
  static void host_copy(void *dst, const void *src, size_t bytes) {
    constexpr size_t kPiece = 1u << 20;
    const long pieces = (long)((bytes + kPiece - 1) / kPiece);
#pragma omp parallel for schedule(static)
    for (long p = 0; p < pieces; p++) {
      const size_t at = (size_t)p * kPiece, len = bytes - at < kPiece ? bytes - at : kPiece;
      std::memcpy((char *)dst + at, (const char *)src + at, len);
    }
  }
  int dev_;
  cudaStream_t st_ = nullptr;
  void *buf_[2] = {nullptr, nullptr};
  cudaEvent_t ev_[2] = {nullptr, nullptr};
  unsigned turn_ = 0;
};

template <typename T>
T *dev_alloc(int dev, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    void *p = nullptr;
    check(sb200_malloc(dev, count * sizeof(T), &p), dev);
    return static_cast<T *>(p);
  }
}
template <typename T>
T *to_device(int dev, const T *h, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!h) return nullptr;
    T *d = dev_alloc<T>(dev, count);
    Stager::get(dev).h2d(d, h, count * sizeof(T));
    Stager::get(dev).sync();  // (the kernels run on the default stream)
    return d;
  }
}
template <typename T>
T *to_host(int dev, const T *d, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!d) return nullptr;
    T *h = new T[count ? count : 1];
    check(sb200_stream_synchronize(dev, nullptr), dev);  // the producer ran on the default stream
    Stager::get(dev).d2h(h, d, count * sizeof(T));
    Stager::get(dev).sync();
    return h;
  }
}
inline void dev_free(int dev, void *p) {
  if (p) sb200_free(dev, p);
}

// ------------------------------------------------------------------ conversion functions
template <typename I, typename N, typename V>
sbase::format::Format *CooCUDACsr(sbase::format::Format *source, sbase::context::Context *to) {
  auto *coo = source->AsAbsolute<sbase::format::COO<I, N, V>>();
  auto *ctx = static_cast<sbase::context::CUDAContext *>(to);
  const int dev = ctx->device_id;
  const auto dims = coo->get_dimensions();
  const size_t nnz = coo->get_num_nnz();
  I *row = to_device(dev, coo->get_row(), nnz), *col = to_device(dev, coo->get_col(), nnz);
  V *vals = to_device(dev, coo->get_vals(), nnz);
  N *o_ptr = dev_alloc<N>(dev, dims[0] + 1);
  I *o_col = dev_alloc<I>(dev, nnz);
  V *o_val = vals ? dev_alloc<V>(dev, nnz) : nullptr;
  const int rc = sb200_coo_to_csr(dev, dims[0], dims[1], nnz, row, col, vals, o_ptr, o_col, o_val,
                                  dtype_of<I>(), dtype_of<N>(), dtype_of<V>(), nullptr);
  sb200_stream_synchronize(dev, nullptr);
  dev_free(dev, row), dev_free(dev, col), dev_free(dev, vals);
  if (rc != SB200_OK) dev_free(dev, o_ptr), dev_free(dev, o_col), dev_free(dev, o_val);
  check(rc, dev);
  return new sbase::format::CUDACSR<I, N, V>(dims[0], dims[1], nnz, o_ptr, o_col, o_val, *ctx,
                                             sbase::format::kOwned);
}

// ---- CSR <-> CUDACSR and CUDACSR -> CUDACSR (peer): staged transfers that keep m and move the
//      n + 1 row pointers (replacing converter_order_two_cuda.cu:11-105)
template <typename I, typename N, typename V>
sbase::format::Format *CsrCUDACsr(sbase::format::Format *source, sbase::context::Context *to) {
  auto *csr = source->AsAbsolute<sbase::format::CSR<I, N, V>>();
  auto *ctx = static_cast<sbase::context::CUDAContext *>(to);
  const int dev = ctx->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  N *d_ptr = dev_alloc<N>(dev, dims[0] + 1);
  I *d_col = dev_alloc<I>(dev, nnz);
  V *d_val = csr->get_vals() ? dev_alloc<V>(dev, nnz) : nullptr;
  Stager &sg = Stager::get(dev);
  sg.h2d(d_ptr, csr->get_row_ptr(), (dims[0] + 1) * sizeof(N));
  sg.h2d(d_col, csr->get_col(), nnz * sizeof(I));
  if constexpr (!std::is_void_v<V>) {
    if (d_val) sg.h2d(d_val, csr->get_vals(), nnz * sizeof(V));
  }
  sg.sync();
  return new sbase::format::CUDACSR<I, N, V>(dims[0], dims[1], nnz, d_ptr, d_col, d_val, *ctx,
                                             sbase::format::kOwned);
}
template <typename I, typename N, typename V>
sbase::format::Format *CUDACsrCsr(sbase::format::Format *source, sbase::context::Context *) {
  auto *csr = source->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  N *h_ptr = new N[dims[0] + 1];
  I *h_col = new I[nnz ? nnz : 1];
  check(sb200_stream_synchronize(dev, nullptr), dev);
  Stager &sg = Stager::get(dev);
  sg.d2h(h_ptr, csr->get_row_ptr(), (dims[0] + 1) * sizeof(N));
  sg.d2h(h_col, csr->get_col(), nnz * sizeof(I));
  V *h_val = nullptr;
  if constexpr (!std::is_void_v<V>) {
    if (csr->get_vals()) {
      h_val = new V[nnz ? nnz : 1];
      sg.d2h(h_val, csr->get_vals(), nnz * sizeof(V));
    }
  }
  sg.sync();
  return new sbase::format::CSR<I, N, V>(dims[0], dims[1], h_ptr, h_col, h_val,
                                         sbase::format::kOwned, /*ignore_sort=*/true);
}
template <typename I, typename N, typename V>
sbase::format::Format *CUDACsrCUDACsr(sbase::format::Format *source, sbase::context::Context *to) {
  auto *csr = source->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  auto *ctx = static_cast<sbase::context::CUDAContext *>(to);
  const int src_dev = csr->get_cuda_context()->device_id, dev = ctx->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  N *d_ptr = dev_alloc<N>(dev, dims[0] + 1);
  I *d_col = dev_alloc<I>(dev, nnz);
  V *d_val = csr->get_vals() ? dev_alloc<V>(dev, nnz) : nullptr;
  int rc = sb200_memcpy_d2d(dev, d_ptr, src_dev, csr->get_row_ptr(), (dims[0] + 1) * sizeof(N),
                            nullptr);
  if (rc == SB200_OK)
    rc = sb200_memcpy_d2d(dev, d_col, src_dev, csr->get_col(), nnz * sizeof(I), nullptr);
  if constexpr (!std::is_void_v<V>) {
    if (rc == SB200_OK && d_val)
      rc = sb200_memcpy_d2d(dev, d_val, src_dev, csr->get_vals(), nnz * sizeof(V), nullptr);
  }
  sb200_stream_synchronize(dev, nullptr);
  if (rc != SB200_OK) dev_free(dev, d_ptr), dev_free(dev, d_col), dev_free(dev, d_val);
  check(rc, dev);
  return new sbase::format::CUDACSR<I, N, V>(dims[0], dims[1], nnz, d_ptr, d_col, d_val, *ctx,
                                             sbase::format::kOwned);
}

template <typename I, typename N, typename V>
sbase::format::Format *CUDACsrCsc(sbase::format::Format *source, sbase::context::Context *) {
  auto *csr = source->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  N *o_ptr = dev_alloc<N>(dev, dims[0] + 1);
  I *o_row = dev_alloc<I>(dev, nnz);
  V *o_val = csr->get_vals() ? dev_alloc<V>(dev, nnz) : nullptr;
  const int rc = sb200_csr_to_csc(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(),
                                  csr->get_vals(), o_ptr, o_row, o_val, dtype_of<I>(),
                                  dtype_of<N>(), dtype_of<V>(), nullptr);
  N *h_ptr = nullptr;
  I *h_row = nullptr;
  V *h_val = nullptr;
  if (rc == SB200_OK) {
    h_ptr = to_host(dev, o_ptr, dims[0] + 1);
    h_row = to_host(dev, o_row, nnz);
    h_val = to_host(dev, o_val, nnz);
  }
  dev_free(dev, o_ptr), dev_free(dev, o_row), dev_free(dev, o_val);
  check(rc, dev);
  return new sbase::format::CSC<I, N, V>(dims[0], dims[1], h_ptr, h_row, h_val,
                                         sbase::format::kOwned, /*ignore_sort=*/true);
}

template <typename I, typename N, typename V>
sbase::format::Format *CUDACsrCoo(sbase::format::Format *source, sbase::context::Context *) {
  auto *csr = source->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  I *o_row = dev_alloc<I>(dev, nnz), *o_col = dev_alloc<I>(dev, nnz);
  V *o_val = csr->get_vals() ? dev_alloc<V>(dev, nnz) : nullptr;
  const int rc = sb200_csr_to_coo(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(),
                                  csr->get_vals(), o_row, o_col, o_val, dtype_of<I>(),
                                  dtype_of<N>(), dtype_of<V>(), nullptr);
  I *h_row = nullptr, *h_col = nullptr;
  V *h_val = nullptr;
  if (rc == SB200_OK) {
    h_row = to_host(dev, o_row, nnz);
    h_col = to_host(dev, o_col, nnz);
    h_val = to_host(dev, o_val, nnz);
  }
  dev_free(dev, o_row), dev_free(dev, o_col), dev_free(dev, o_val);
  check(rc, dev);
  return new sbase::format::COO<I, N, V>(dims[0], dims[1], nnz, h_row, h_col, h_val,
                                         sbase::format::kOwned, /*ignore_sort=*/true);
}

inline bool CPUToCUDA(sbase::context::Context *from, sbase::context::Context *to) {
  return from->get_id() == sbase::context::CPUContext::get_id_static() &&
         to->get_id() == sbase::context::CUDAContext::get_id_static();
}
inline bool CUDAToCPU(sbase::context::Context *from, sbase::context::Context *to) {
  return from->get_id() == sbase::context::CUDAContext::get_id_static() &&
         to->get_id() == sbase::context::CPUContext::get_id_static();
}

template <typename I, typename N, typename V>
void RegisterConversions(sbase::converter::Converter &conv) {
  using namespace sbase::format;
  for (bool mv : {false, true}) {
    conv.RegisterConversionFunction(COO<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(),
                                    CooCUDACsr<I, N, V>, CPUToCUDA, mv);
    conv.RegisterConversionFunction(CUDACSR<I, N, V>::get_id_static(), CSC<I, N, V>::get_id_static(),
                                    CUDACsrCsc<I, N, V>, CUDAToCPU, mv);
    conv.RegisterConversionFunction(CUDACSR<I, N, V>::get_id_static(), COO<I, N, V>::get_id_static(),
                                    CUDACsrCoo<I, N, V>, CUDAToCPU, mv);
  }
}

inline bool CUDAToCUDA(sbase::context::Context *from, sbase::context::Context *to) {
  return from->get_id() == sbase::context::CUDAContext::get_id_static() &&
         to->get_id() == sbase::context::CUDAContext::get_id_static();
}

// Replaces SparseBase's own CSR <-> CUDACSR / CUDACSR -> CUDACSR functions in `conv` (a converter
// takes the FIRST registered function whose condition holds, so the old ones are cleared).
template <typename I, typename N, typename V>
void RegisterTransfers(sbase::converter::Converter &conv) {
  using namespace sbase::format;
  for (bool mv : {false, true}) {
    conv.ClearConversionFunctions(CSR<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(), mv);
    conv.ClearConversionFunctions(CUDACSR<I, N, V>::get_id_static(), CSR<I, N, V>::get_id_static(), mv);
    conv.ClearConversionFunctions(CUDACSR<I, N, V>::get_id_static(),
                                  CUDACSR<I, N, V>::get_id_static(), mv);
    conv.RegisterConversionFunction(CSR<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(),
                                    CsrCUDACsr<I, N, V>, CPUToCUDA, mv);
    conv.RegisterConversionFunction(CUDACSR<I, N, V>::get_id_static(), CSR<I, N, V>::get_id_static(),
                                    CUDACsrCsr<I, N, V>, CUDAToCPU, mv);
    conv.RegisterConversionFunction(CUDACSR<I, N, V>::get_id_static(),
                                    CUDACSR<I, N, V>::get_id_static(), CUDACsrCUDACsr<I, N, V>,
                                    CUDAToCUDA, mv);
  }
}

// ------------------------------------------------------------------ implementation functions
template <typename I, typename N, typename V>
I *DegreeReorderCUDACSR(std::vector<sbase::format::Format *> formats,
                        sbase::utils::Parameters *params) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  auto *p = static_cast<sbase::reorder::DegreeReorderParams *>(params);
  const int dev = csr->get_cuda_context()->device_id;
  const size_t n = csr->get_dimensions()[0];
  I *inv = dev_alloc<I>(dev, n);
  const int rc = sb200_degree_reorder(dev, n, csr->get_row_ptr(), p->ascending ? 1 : 0, inv,
                                      dtype_of<I>(), dtype_of<N>(), nullptr);
  I *h = rc == SB200_OK ? to_host(dev, inv, n) : nullptr;
  dev_free(dev, inv);
  check(rc, dev);
  return h;
}

// reorder::BOBAReorder works on a COO (reorder/boba_reorder.cc:18-21); the device format of the
// reference is CUDACSR, so the rows are expanded on the device first (sb200_csr_to_coo).
template <typename I, typename N, typename V>
I *BOBAReorderCUDACSR(std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz(), nodes = dims[0] > dims[1] ? dims[0] : dims[1];
  I *d_row = dev_alloc<I>(dev, nnz ? nnz : 1), *d_col = dev_alloc<I>(dev, nnz ? nnz : 1);
  I *inv = dev_alloc<I>(dev, nodes ? nodes : 1);
  int rc = sb200_csr_to_coo(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(), nullptr,
                            d_row, d_col, nullptr, dtype_of<I>(), dtype_of<N>(), SB200_VOID,
                            nullptr);
  if (rc == SB200_OK)
    rc = sb200_boba_reorder(dev, dims[0], dims[1], nnz, d_row, d_col, inv, dtype_of<I>(), nullptr);
  I *h = rc == SB200_OK ? to_host(dev, inv, nodes) : nullptr;
  dev_free(dev, d_row), dev_free(dev, d_col), dev_free(dev, inv);
  check(rc, dev);
  return h;
}

template <typename I, typename N, typename V>
I *RCMReorderCUDACSR(std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const size_t n = csr->get_dimensions()[0];
  I *inv = dev_alloc<I>(dev, n);
  const int rc = sb200_rcm_reorder(dev, n, csr->get_num_nnz(), csr->get_row_ptr(), csr->get_col(),
                                   inv, dtype_of<I>(), dtype_of<N>(), nullptr);
  I *h = rc == SB200_OK ? to_host(dev, inv, n) : nullptr;
  dev_free(dev, inv);
  check(rc, dev);
  return h;
}

template <typename I, typename N, typename V>
sbase::format::FormatOrderTwo<I, N, V> *PermuteOrderTwoCUDACSR(
    std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *params) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  auto *p = static_cast<sbase::permute::PermuteOrderTwoParams<I> *>(params);
  auto *ctx = csr->get_cuda_context();
  const int dev = ctx->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  I *d_row = to_device(dev, p->row_order, dims[0]), *d_col = to_device(dev, p->col_order, dims[1]);
  N *o_ptr = dev_alloc<N>(dev, dims[0] + 1);
  I *o_col = dev_alloc<I>(dev, nnz);
  V *o_val = csr->get_vals() ? dev_alloc<V>(dev, nnz) : nullptr;
  const int rc = sb200_permute2d(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(),
                                 csr->get_vals(), d_row, d_col, o_ptr, o_col, o_val, dtype_of<I>(),
                                 dtype_of<N>(), dtype_of<V>(), nullptr);
  sb200_stream_synchronize(dev, nullptr);
  dev_free(dev, d_row), dev_free(dev, d_col);
  if (rc != SB200_OK) dev_free(dev, o_ptr), dev_free(dev, o_col), dev_free(dev, o_val);
  check(rc, dev);
  return new sbase::format::CUDACSR<I, N, V>(dims[0], dims[1], nnz, o_ptr, o_col, o_val, *ctx,
                                             sbase::format::kOwned);
}

template <typename I, typename V>
sbase::format::FormatOrderOne<V> *PermuteOrderOneCUDAArray(
    std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *params) {
  auto *arr = formats[0]->AsAbsolute<sbase::format::CUDAArray<V>>();
  auto *p = static_cast<sbase::permute::PermuteOrderOneParams<I> *>(params);
  auto *ctx = static_cast<sbase::context::CUDAContext *>(arr->get_context());
  const int dev = ctx->device_id;
  const size_t len = arr->get_num_nnz();
  I *d_order = to_device(dev, p->order, len);
  V *out = dev_alloc<V>(dev, len);
  const int rc = sb200_permute1d(dev, len, arr->get_vals(), d_order, out, dtype_of<I>(),
                                 dtype_of<V>(), nullptr);
  sb200_stream_synchronize(dev, nullptr);
  dev_free(dev, d_order);
  if (rc != SB200_OK) dev_free(dev, out);
  check(rc, dev);
  return new sbase::format::CUDAArray<V>(len, out, *ctx, sbase::format::kOwned);
}

template <typename I, typename N, typename V, typename F>
F *DegreeDistributionCUDACSR(std::vector<sbase::format::Format *> formats,
                             sbase::utils::Parameters *) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const size_t n = csr->get_dimensions()[0];
  F *dist = dev_alloc<F>(dev, n);
  const int rc = sb200_degree_distribution(dev, n, csr->get_num_nnz(), csr->get_row_ptr(), dist,
                                           dtype_of<N>(), dtype_of<F>(), nullptr);
  F *h = rc == SB200_OK ? to_host(dev, dist, n) : nullptr;
  dev_free(dev, dist);
  check(rc, dev);
  return h;
}

template <typename I, typename N, typename V>
I *DegreesCUDACSR(std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  const int dev = csr->get_cuda_context()->device_id;
  const size_t n = csr->get_dimensions()[0];
  I *deg = dev_alloc<I>(dev, n);
  const int rc = sb200_degrees(dev, n, csr->get_row_ptr(), deg, dtype_of<I>(), dtype_of<N>(), nullptr);
  I *h = rc == SB200_OK ? to_host(dev, deg, n) : nullptr;
  dev_free(dev, deg);
  check(rc, dev);
  return h;
}

// reorder::ReorderHeatmap on {CUDACSR, CUDAArray, CUDAArray} (the reference registers
// {CSR, Array, Array}, reorder/reorder_heatmap.cc:11-15).  The grid is small: it comes back as
// a host Array, the type the reference's own function returns.
template <typename I, typename N, typename V, typename F>
sbase::format::FormatOrderOne<F> *ReorderHeatmapCUDACSR(
    std::vector<sbase::format::Format *> formats, sbase::utils::Parameters *params) {
  auto *csr = formats[0]->AsAbsolute<sbase::format::CUDACSR<I, N, V>>();
  auto *pr = formats[1]->AsAbsolute<sbase::format::CUDAArray<I>>();
  auto *pc = formats[2]->AsAbsolute<sbase::format::CUDAArray<I>>();
  auto *p = static_cast<sbase::reorder::ReorderHeatmapParams *>(params);
  const int dev = csr->get_cuda_context()->device_id;
  const auto dims = csr->get_dimensions();
  const int b = p->num_parts;
  if (b < 1 || (size_t)b > dims[0] || (size_t)b > dims[1])
    throw sbase::utils::ReorderException(
        "Cannot generate heatmap for matrix when num_parts > number of rows or columns");
  F *d_heat = dev_alloc<F>(dev, (size_t)b * b);
  const int rc = sb200_reorder_heatmap(dev, dims[0], dims[1], csr->get_num_nnz(),
                                       csr->get_row_ptr(), csr->get_col(), pr->get_vals(),
                                       pc->get_vals(), b, d_heat, dtype_of<I>(), dtype_of<N>(),
                                       dtype_of<F>(), nullptr);
  F *h = rc == SB200_OK ? to_host(dev, d_heat, (size_t)b * b) : nullptr;
  dev_free(dev, d_heat);
  check(rc, dev);
  return new sbase::format::Array<F>((size_t)b * b, h, sbase::format::kOwned);
}

// ------------------------------------------------------------------ registration helpers
template <typename I, typename N, typename V, typename F>
void Register(sbase::reorder::ReorderHeatmap<I, N, V, F> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static(),
                       sbase::format::CUDAArray<I>::get_id_static(),
                       sbase::format::CUDAArray<I>::get_id_static()},
                      ReorderHeatmapCUDACSR<I, N, V, F>);
}
template <typename I, typename N, typename V>
void Register(sbase::reorder::DegreeReorder<I, N, V> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()}, DegreeReorderCUDACSR<I, N, V>);
}
template <typename I, typename N, typename V>
void Register(sbase::reorder::BOBAReorder<I, N, V> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()}, BOBAReorderCUDACSR<I, N, V>);
}
template <typename I, typename N, typename V>
void Register(sbase::reorder::RCMReorder<I, N, V> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()}, RCMReorderCUDACSR<I, N, V>);
}
template <typename I, typename N, typename V>
void Register(sbase::permute::PermuteOrderTwo<I, N, V> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()},
                      PermuteOrderTwoCUDACSR<I, N, V>);
}
template <typename I, typename V>
void Register(sbase::permute::PermuteOrderOne<I, V> &op) {
  op.RegisterFunction({sbase::format::CUDAArray<V>::get_id_static()}, PermuteOrderOneCUDAArray<I, V>);
}
template <typename I, typename N, typename V, typename F>
void Register(sbase::feature::DegreeDistribution<I, N, V, F> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()},
                      DegreeDistributionCUDACSR<I, N, V, F>);
}
template <typename I, typename N, typename V>
void Register(sbase::feature::Degrees<I, N, V> &op) {
  op.RegisterFunction({sbase::format::CUDACSR<I, N, V>::get_id_static()}, DegreesCUDACSR<I, N, V>);
}

}  // namespace sb200_plugin
