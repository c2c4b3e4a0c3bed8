// sb200/format.h -- the sparse containers of the hot path and the conversion functions
// registered between them.
//
// Interfaces mirrored (reference paths relative to src/sparsebase/):
//   format::FormatImplementation              format/format_implementation.h:22-51
//   format::FormatOrderTwo / FormatOrderOne   format/format_order_two.h:23-157, format_order_one.h
//   format::CSR / COO / CSC / Array           format/csr.h:27-60, coo.h:25-58, csc.h, array.h:15-36
//   format::CUDACSR / CUDAArray               format/cuda_csr_cuda.cuh:20-59, cuda_array_cuda.cuh:10-32
//   converter::ConverterOrderTwo / One        converter/converter_order_two.cc:257-340,
//                                             converter/converter_order_one.cc:25-42
//   conversion functions                      converter/converter_order_two.cc:20-246,
//                                             converter/converter_order_two_cuda.cu:11-105,
//                                             converter/converter_order_one_cuda.cu:10-43,
//                                             converter/converter_cuda.cu:12-21 (CUDAPeerToPeer)
// New in the same style (SURVEY.md 8b "gaps"): format::CUDACOO and format::CUDACSC, so that
// whole pipelines stay in HBM.
//
// Where the arithmetic runs.  The reference's constructors sort on the CPU (coo.cc:96-157,
// csr.cc:99-157, csc.cc:99-157).  Here the same check + sort runs on the GPU: device formats
// call sb200_coo_sort / sb200_compressed_sort on their own arrays; host formats stage their
// arrays through device sb200::default_device() and copy the result back in place only when
// something had to be sorted.  There is no CPU implementation: with `ignore_sort = false` a
// host constructor needs a GPU and throws otherwise.
#pragma once
#include <cstring>

#include "converter.h"

namespace sparsebase {

namespace converter {
template <typename IDType, typename NNZType, typename ValueType>
class ConverterOrderTwo;
template <typename ValueType>
class ConverterOrderOne;
}  // namespace converter

namespace format {

// ---------------------------------------------------------------- array ownership
template <typename T>
using ArrayHandle = std::unique_ptr<T, std::function<void(T *)>>;

template <typename T>
ArrayHandle<T> host_handle(T *p, Ownership own) {
  if (own == kOwned)
    return ArrayHandle<T>(p, [](T *q) {
      if constexpr (!std::is_void_v<T>) delete[] q;
    });
  return ArrayHandle<T>(p, [](T *) {});
}
template <typename T>
ArrayHandle<T> device_handle(T *p, Ownership own, int device) {
  if (own == kOwned) return ArrayHandle<T>(p, [device](T *q) { sb200::device_free(device, q); });
  return ArrayHandle<T>(p, [](T *) {});
}
template <typename T>
T *release_handle(ArrayHandle<T> &h) {
  T *p = h.release();
  h = ArrayHandle<T>(p, [](T *) {});  // keep pointing at it, no longer responsible
  return p;
}
template <typename T>
T *host_copy(const T *src, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!src) return nullptr;
    T *dst = new T[count ? count : 1];
    std::memcpy(dst, src, count * sizeof(T));
    return dst;
  }
}
template <typename T>
T *device_copy(int dst_device, int src_device, const T *src, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!src) return nullptr;
    T *dst = sb200::device_alloc<T>(dst_device, count);
    sb200::check(sb200_memcpy_d2d(dst_device, dst, src_device, src, count * sizeof(T), nullptr),
                 dst_device);
    sb200::check(sb200_stream_synchronize(dst_device, nullptr), dst_device);
    return dst;
  }
}

// ---------------------------------------------------------------- shared plumbing
class FormatImplementation : public Format {
 public:
  std::vector<DimensionType> get_dimensions() const override { return dimension_; }
  DimensionType get_num_nnz() const override { return nnz_; }
  DimensionType get_order() const override { return order_; }
  context::Context *get_context() const override { return context_.get(); }
  std::shared_ptr<converter::Converter const> get_converter() const override { return converter_; }
  void set_converter(std::shared_ptr<converter::Converter> c) { converter_ = std::move(c); }

 protected:
  DimensionType order_ = 0;
  std::vector<DimensionType> dimension_;
  DimensionType nnz_ = 0;
  std::unique_ptr<context::Context> context_;
  std::shared_ptr<converter::Converter> converter_;

  int cuda_device() const { return static_cast<context::CUDAContext *>(context_.get())->device_id; }
};

inline context::CPUContext *shared_cpu_context() {
  static context::CPUContext cpu;
  return &cpu;
}

template <typename IDType, typename NNZType, typename ValueType>
class FormatOrderTwo : public FormatImplementation {
 public:
  FormatOrderTwo() {
    this->set_converter(converter::ConverterStore::GetStore()
                            .get_converter<converter::ConverterOrderTwo<IDType, NNZType, ValueType>>());
  }

  // Convert to ToType<IDType, NNZType, ValueType>.  `to_context == nullptr` means "where this
  // object lives, or the host" (the reference uses the object's own context only; the host is
  // added so that device results can be brought back with a bare Convert<CSR>()).
  template <template <typename, typename, typename> class ToType>
  ToType<IDType, NNZType, ValueType> *Convert(context::Context *to_context = nullptr,
                                              bool is_move_conversion = false) {
    std::vector<context::Context *> ctxs;
    if (to_context)
      ctxs = {to_context};
    else
      ctxs = {this->get_context(), shared_cpu_context()};
    return Convert<ToType>(ctxs, is_move_conversion);
  }
  template <template <typename, typename, typename> class ToType>
  ToType<IDType, NNZType, ValueType> *Convert(const std::vector<context::Context *> &to_contexts,
                                              bool is_move_conversion = false) {
    static_assert(std::is_base_of_v<FormatOrderTwo<IDType, NNZType, ValueType>,
                                    ToType<IDType, NNZType, ValueType>>,
                  "T must be an order two format");
    std::vector<context::Context *> ctxs = to_contexts;
    if (ctxs.empty()) ctxs = {this->get_context()};
    return this->get_converter()
        ->Convert(this, ToType<IDType, NNZType, ValueType>::get_id_static(), ctxs,
                  is_move_conversion)
        ->template AsAbsolute<ToType<IDType, NNZType, ValueType>>();
  }
  template <template <typename, typename, typename> typename T>
  T<IDType, NNZType, ValueType> *As() {
    using TBase = T<IDType, NNZType, ValueType>;
    static_assert(std::is_base_of_v<FormatOrderTwo<IDType, NNZType, ValueType>, TBase>,
                  "Cannot cast to a non-FormatOrderTwo class");
    if (this->get_id() == std::type_index(typeid(TBase))) return static_cast<TBase *>(this);
    throw utils::TypeException(this->get_name(), utils::demangle(typeid(TBase).name()));
  }
  template <template <typename, typename, typename> typename T>
  bool Is() {
    return this->get_id() == std::type_index(typeid(T<IDType, NNZType, ValueType>));
  }
};

template <typename ValueType>
class FormatOrderOne : public FormatImplementation {
 public:
  FormatOrderOne() {
    this->set_converter(converter::ConverterStore::GetStore()
                            .get_converter<converter::ConverterOrderOne<ValueType>>());
  }
  template <template <typename> class ToType>
  ToType<ValueType> *Convert(context::Context *to_context = nullptr,
                             bool is_move_conversion = false) {
    std::vector<context::Context *> ctxs;
    if (to_context)
      ctxs = {to_context};
    else
      ctxs = {this->get_context(), shared_cpu_context()};
    return Convert<ToType>(ctxs, is_move_conversion);
  }
  template <template <typename> class ToType>
  ToType<ValueType> *Convert(const std::vector<context::Context *> &to_contexts,
                             bool is_move_conversion = false) {
    static_assert(std::is_base_of_v<FormatOrderOne<ValueType>, ToType<ValueType>>,
                  "T must be an order one format");
    std::vector<context::Context *> ctxs = to_contexts;
    if (ctxs.empty()) ctxs = {this->get_context()};
    return this->get_converter()
        ->Convert(this, ToType<ValueType>::get_id_static(), ctxs, is_move_conversion)
        ->template AsAbsolute<ToType<ValueType>>();
  }
  template <template <typename> typename T>
  T<ValueType> *As() {
    using TBase = T<ValueType>;
    if (this->get_id() == std::type_index(typeid(TBase))) return static_cast<TBase *>(this);
    throw utils::TypeException(this->get_name(), utils::demangle(typeid(TBase).name()));
  }
  template <template <typename> typename T>
  bool Is() {
    return this->get_id() == std::type_index(typeid(T<ValueType>));
  }
};

// =====================================================================================
// Compressed (CSR / CSC) and coordinate layouts, host and device flavours.  One class
// template per storage place keeps the eight public classes below short.
// =====================================================================================
namespace detail {

// ptr[n_seg+1] + idx[nnz] + vals[nnz], host memory
template <typename Self, typename I, typename N, typename V>
class HostCompressed : public utils::IdentifiableImplementation<Self, FormatOrderTwo<I, N, V>> {
 protected:
  ArrayHandle<N> ptr_;
  ArrayHandle<I> idx_;
  ArrayHandle<V> vals_;

  // n_seg segments (rows for CSR; the reference's CSC also uses n, csc.cc:87), indices < n_idx
  void init(I n, I m, N *ptr, I *idx, V *vals, Ownership own, bool ignore_sort) {
    this->order_ = 2;
    this->dimension_ = {(DimensionType)n, (DimensionType)m};
    this->nnz_ = ptr ? (DimensionType)ptr[n] : 0;
    this->context_ = std::make_unique<context::CPUContext>();
    ptr_ = host_handle(ptr, own);
    idx_ = host_handle(idx, own);
    vals_ = host_handle(vals, own);
    if (!ignore_sort && this->nnz_ > 0) sort_through_device(n, m);
  }
  void sort_through_device(I n, I m) {
    const int dev = sb200::default_device();
    const size_t nnz = this->nnz_;
    sb200::DeviceScratch<N> d_ptr(dev, ptr_.get(), (size_t)n + 1);
    sb200::DeviceScratch<I> d_idx(dev, idx_.get(), nnz);
    sb200::DeviceScratch<V> d_val(dev, vals_.get(), nnz);
    int was_sorted = 1;
    sb200::check(sb200_compressed_sort(dev, n, m, nnz, d_ptr.get(), d_idx.get(), d_val.get(),
                                       sb200::dtype_of<I>(), sb200::dtype_of<N>(),
                                       sb200::dtype_of<V>(), &was_sorted, nullptr),
                 dev);
    if (was_sorted) return;
    sb200::check(sb200_memcpy_d2h(dev, idx_.get(), d_idx.get(), nnz * sizeof(I), nullptr), dev);
    if constexpr (!std::is_void_v<V>)
      if (vals_) sb200::check(sb200_memcpy_d2h(dev, vals_.get(), d_val.get(), nnz * sizeof(V), nullptr), dev);
    sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
  }
  void copy_from(const HostCompressed &rhs) {
    this->order_ = 2;
    this->dimension_ = rhs.dimension_;
    this->nnz_ = rhs.nnz_;
    this->context_ = std::make_unique<context::CPUContext>();
    ptr_ = host_handle(host_copy(rhs.ptr_.get(), (size_t)rhs.dimension_[0] + 1), kOwned);
    idx_ = host_handle(host_copy(rhs.idx_.get(), rhs.nnz_), kOwned);
    vals_ = host_handle(host_copy(rhs.vals_.get(), rhs.nnz_), kOwned);
  }

 public:
  Format *Clone() const override { return new Self(static_cast<const Self &>(*this)); }
  V *get_vals() const { return vals_.get(); }
  V *release_vals() { return release_handle(vals_); }
  void set_vals(V *p, Ownership own = kNotOwned) { vals_ = host_handle(p, own); }
};

// row[nnz] + col[nnz] + vals[nnz], host memory
template <typename Self, typename I, typename N, typename V>
class HostCoordinate : public utils::IdentifiableImplementation<Self, FormatOrderTwo<I, N, V>> {
 protected:
  ArrayHandle<I> row_;
  ArrayHandle<I> col_;
  ArrayHandle<V> vals_;
};

}  // namespace detail

// ---------------------------------------------------------------- host CSR
template <typename IDType, typename NNZType, typename ValueType>
class CSR : public detail::HostCompressed<CSR<IDType, NNZType, ValueType>, IDType, NNZType, ValueType> {
 public:
  // csr.cc:78-159: nnz = row_ptr[n]; unless ignore_sort, every row is checked and, if any row
  // is unsorted, every row is sorted by (col, val) IN PLACE (on the GPU here).
  CSR(IDType n, IDType m, NNZType *row_ptr, IDType *col, ValueType *vals,
      Ownership own = kNotOwned, bool ignore_sort = false) {
    this->init(n, m, row_ptr, col, vals, own, ignore_sort);
  }
  CSR(const CSR &rhs) { this->copy_from(rhs); }
  NNZType *get_row_ptr() const { return this->ptr_.get(); }
  IDType *get_col() const { return this->idx_.get(); }
  NNZType *release_row_ptr() { return release_handle(this->ptr_); }
  IDType *release_col() { return release_handle(this->idx_); }
  void set_row_ptr(NNZType *p, Ownership own = kNotOwned) { this->ptr_ = host_handle(p, own); }
  void set_col(IDType *p, Ownership own = kNotOwned) { this->idx_ = host_handle(p, own); }
};

// ---------------------------------------------------------------- host CSC
template <typename IDType, typename NNZType, typename ValueType>
class CSC : public detail::HostCompressed<CSC<IDType, NNZType, ValueType>, IDType, NNZType, ValueType> {
 public:
  // csc.cc:79-159.  col_ptr has n+1 entries -- the reference's square-matrix convention.
  CSC(IDType n, IDType m, NNZType *col_ptr, IDType *row, ValueType *vals,
      Ownership own = kNotOwned, bool ignore_sort = false) {
    this->init(n, m, col_ptr, row, vals, own, ignore_sort);
  }
  CSC(const CSC &rhs) { this->copy_from(rhs); }
  NNZType *get_col_ptr() const { return this->ptr_.get(); }
  IDType *get_row() const { return this->idx_.get(); }
  NNZType *release_col_ptr() { return release_handle(this->ptr_); }
  IDType *release_row() { return release_handle(this->idx_); }
  void set_col_ptr(NNZType *p, Ownership own = kNotOwned) { this->ptr_ = host_handle(p, own); }
  void set_row(IDType *p, Ownership own = kNotOwned) { this->idx_ = host_handle(p, own); }
};

// ---------------------------------------------------------------- host COO
template <typename IDType, typename NNZType, typename ValueType>
class COO : public utils::IdentifiableImplementation<COO<IDType, NNZType, ValueType>,
                                                     FormatOrderTwo<IDType, NNZType, ValueType>> {
 public:
  // coo.cc:76-158: unless ignore_sort, the (row, col) sequence is checked and, on the first
  // inversion, row/col/vals are sorted by (row, col) IN PLACE (on the GPU here).
  COO(IDType n, IDType m, NNZType nnz, IDType *row, IDType *col, ValueType *vals,
      Ownership own = kNotOwned, bool ignore_sort = false) {
    this->order_ = 2;
    this->dimension_ = {(DimensionType)n, (DimensionType)m};
    this->nnz_ = (DimensionType)nnz;
    this->context_ = std::make_unique<context::CPUContext>();
    row_ = host_handle(row, own);
    col_ = host_handle(col, own);
    vals_ = host_handle(vals, own);
    if (!ignore_sort && nnz > 1) {
      const int dev = sb200::default_device();
      const size_t cnt = (size_t)nnz;
      sb200::DeviceScratch<IDType> d_row(dev, row, cnt), d_col(dev, col, cnt);
      sb200::DeviceScratch<ValueType> d_val(dev, vals, cnt);
      int was_sorted = 1;
      sb200::check(sb200_coo_sort(dev, n, m, nnz, d_row.get(), d_col.get(), d_val.get(),
                                  sb200::dtype_of<IDType>(), sb200::dtype_of<ValueType>(),
                                  &was_sorted, nullptr),
                   dev);
      if (!was_sorted) {
        sb200::check(sb200_memcpy_d2h(dev, row, d_row.get(), cnt * sizeof(IDType), nullptr), dev);
        sb200::check(sb200_memcpy_d2h(dev, col, d_col.get(), cnt * sizeof(IDType), nullptr), dev);
        if constexpr (!std::is_void_v<ValueType>)
          if (vals)
            sb200::check(sb200_memcpy_d2h(dev, vals, d_val.get(), cnt * sizeof(ValueType), nullptr), dev);
        sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
      }
    }
  }
  COO(const COO &rhs) {
    this->order_ = 2;
    this->dimension_ = rhs.dimension_;
    this->nnz_ = rhs.nnz_;
    this->context_ = std::make_unique<context::CPUContext>();
    row_ = host_handle(host_copy(rhs.row_.get(), rhs.nnz_), kOwned);
    col_ = host_handle(host_copy(rhs.col_.get(), rhs.nnz_), kOwned);
    vals_ = host_handle(host_copy(rhs.vals_.get(), rhs.nnz_), kOwned);
  }
  Format *Clone() const override { return new COO(*this); }
  IDType *get_row() const { return row_.get(); }
  IDType *get_col() const { return col_.get(); }
  ValueType *get_vals() const { return vals_.get(); }
  IDType *release_row() { return release_handle(row_); }
  IDType *release_col() { return release_handle(col_); }
  ValueType *release_vals() { return release_handle(vals_); }
  void set_row(IDType *p, Ownership own = kNotOwned) { row_ = host_handle(p, own); }
  void set_col(IDType *p, Ownership own = kNotOwned) { col_ = host_handle(p, own); }
  void set_vals(ValueType *p, Ownership own = kNotOwned) { vals_ = host_handle(p, own); }

 protected:
  ArrayHandle<IDType> row_, col_;
  ArrayHandle<ValueType> vals_;
};

// ---------------------------------------------------------------- host Array
template <typename ValueType>
class Array : public utils::IdentifiableImplementation<Array<ValueType>, FormatOrderOne<ValueType>> {
 public:
  Array(DimensionType nnz, ValueType *vals, Ownership own = kNotOwned) {
    this->order_ = 1;
    this->dimension_ = {nnz};
    this->nnz_ = nnz;
    this->context_ = std::make_unique<context::CPUContext>();
    vals_ = host_handle(vals, own);
  }
  Array(const Array &rhs) : Array(rhs.nnz_, host_copy(rhs.vals_.get(), rhs.nnz_), kOwned) {}
  Format *Clone() const override { return new Array(*this); }
  ValueType *get_vals() const { return vals_.get(); }
  ValueType *release_vals() { return release_handle(vals_); }
  void set_vals(ValueType *p, Ownership own = kNotOwned) { vals_ = host_handle(p, own); }

 protected:
  ArrayHandle<ValueType> vals_;
};

// =====================================================================================
// Device formats.  Arrays are cudaMalloc-compatible allocations on context.device_id; with
// kOwned (the default, as in the reference) they are released with sb200_free (= cudaFree,
// the reference's utils::CUDADeleter).  Unlike the reference's CUDACSR these keep `m`
// (converter_order_two_cuda.cu:38 passes n for m).
// =====================================================================================
namespace detail {
template <typename Self, typename I, typename N, typename V>
class DeviceCompressed : public utils::IdentifiableImplementation<Self, FormatOrderTwo<I, N, V>> {
 protected:
  ArrayHandle<N> ptr_;
  ArrayHandle<I> idx_;
  ArrayHandle<V> vals_;
  void init(I n, I m, N nnz, N *ptr, I *idx, V *vals, const context::CUDAContext &ctx,
            Ownership own) {
    this->order_ = 2;
    this->dimension_ = {(DimensionType)n, (DimensionType)m};
    this->nnz_ = (DimensionType)nnz;
    this->context_ = std::make_unique<context::CUDAContext>(ctx);
    ptr_ = device_handle(ptr, own, ctx.device_id);
    idx_ = device_handle(idx, own, ctx.device_id);
    vals_ = device_handle(vals, own, ctx.device_id);
  }
  void copy_from(const DeviceCompressed &rhs) {
    const int dev = rhs.cuda_device();
    init((I)rhs.dimension_[0], (I)rhs.dimension_[1], (N)rhs.nnz_,
         device_copy(dev, dev, rhs.ptr_.get(), (size_t)rhs.dimension_[0] + 1),
         device_copy(dev, dev, rhs.idx_.get(), rhs.nnz_),
         device_copy(dev, dev, rhs.vals_.get(), rhs.nnz_), *rhs.get_cuda_context(), kOwned);
  }

 public:
  Format *Clone() const override { return new Self(static_cast<const Self &>(*this)); }
  context::CUDAContext *get_cuda_context() const {
    return static_cast<context::CUDAContext *>(this->get_context());
  }
  V *get_vals() const { return vals_.get(); }
  V *release_vals() { return release_handle(vals_); }
  void set_vals(V *p, context::CUDAContext ctx, Ownership own = kNotOwned) {
    vals_ = device_handle(p, own, ctx.device_id);
  }
};
}  // namespace detail

template <typename IDType, typename NNZType, typename ValueType>
class CUDACSR : public detail::DeviceCompressed<CUDACSR<IDType, NNZType, ValueType>, IDType,
                                                NNZType, ValueType> {
 public:
  CUDACSR(IDType n, IDType m, NNZType nnz, NNZType *row_ptr, IDType *col, ValueType *vals,
          context::CUDAContext context, Ownership own = kOwned) {
    this->init(n, m, nnz, row_ptr, col, vals, context, own);
  }
  CUDACSR(const CUDACSR &rhs) { this->copy_from(rhs); }
  NNZType *get_row_ptr() const { return this->ptr_.get(); }
  IDType *get_col() const { return this->idx_.get(); }
  NNZType *release_row_ptr() { return release_handle(this->ptr_); }
  IDType *release_col() { return release_handle(this->idx_); }
  void set_row_ptr(NNZType *p, context::CUDAContext ctx, Ownership own = kNotOwned) {
    this->ptr_ = device_handle(p, own, ctx.device_id);
  }
  void set_col(IDType *p, context::CUDAContext ctx, Ownership own = kNotOwned) {
    this->idx_ = device_handle(p, own, ctx.device_id);
  }
};

template <typename IDType, typename NNZType, typename ValueType>
class CUDACSC : public detail::DeviceCompressed<CUDACSC<IDType, NNZType, ValueType>, IDType,
                                                NNZType, ValueType> {
 public:
  CUDACSC(IDType n, IDType m, NNZType nnz, NNZType *col_ptr, IDType *row, ValueType *vals,
          context::CUDAContext context, Ownership own = kOwned) {
    this->init(n, m, nnz, col_ptr, row, vals, context, own);
  }
  CUDACSC(const CUDACSC &rhs) { this->copy_from(rhs); }
  NNZType *get_col_ptr() const { return this->ptr_.get(); }
  IDType *get_row() const { return this->idx_.get(); }
  NNZType *release_col_ptr() { return release_handle(this->ptr_); }
  IDType *release_row() { return release_handle(this->idx_); }
};

template <typename IDType, typename NNZType, typename ValueType>
class CUDACOO : public utils::IdentifiableImplementation<CUDACOO<IDType, NNZType, ValueType>,
                                                         FormatOrderTwo<IDType, NNZType, ValueType>> {
 public:
  // The COO constructor semantics of coo.cc:96-157 on device arrays: unless ignore_sort, the
  // arrays are checked and sorted in place by (row, col) (sb200_coo_sort).
  CUDACOO(IDType n, IDType m, NNZType nnz, IDType *row, IDType *col, ValueType *vals,
          context::CUDAContext context, Ownership own = kOwned, bool ignore_sort = false) {
    this->order_ = 2;
    this->dimension_ = {(DimensionType)n, (DimensionType)m};
    this->nnz_ = (DimensionType)nnz;
    this->context_ = std::make_unique<context::CUDAContext>(context);
    const int dev = context.device_id;
    row_ = device_handle(row, own, dev);
    col_ = device_handle(col, own, dev);
    vals_ = device_handle(vals, own, dev);
    if (!ignore_sort && nnz > 1)
      sb200::check(sb200_coo_sort(dev, n, m, nnz, row, col, vals, sb200::dtype_of<IDType>(),
                                  sb200::dtype_of<ValueType>(), nullptr, nullptr),
                   dev);
  }
  CUDACOO(const CUDACOO &rhs)
      : CUDACOO((IDType)rhs.dimension_[0], (IDType)rhs.dimension_[1], (NNZType)rhs.nnz_,
                device_copy(rhs.cuda_device(), rhs.cuda_device(), rhs.row_.get(), rhs.nnz_),
                device_copy(rhs.cuda_device(), rhs.cuda_device(), rhs.col_.get(), rhs.nnz_),
                device_copy(rhs.cuda_device(), rhs.cuda_device(), rhs.vals_.get(), rhs.nnz_),
                *rhs.get_cuda_context(), kOwned, true) {}
  Format *Clone() const override { return new CUDACOO(*this); }
  context::CUDAContext *get_cuda_context() const {
    return static_cast<context::CUDAContext *>(this->get_context());
  }
  IDType *get_row() const { return row_.get(); }
  IDType *get_col() const { return col_.get(); }
  ValueType *get_vals() const { return vals_.get(); }
  IDType *release_row() { return release_handle(row_); }
  IDType *release_col() { return release_handle(col_); }
  ValueType *release_vals() { return release_handle(vals_); }

 protected:
  ArrayHandle<IDType> row_, col_;
  ArrayHandle<ValueType> vals_;
};

template <typename ValueType>
class CUDAArray
    : public utils::IdentifiableImplementation<CUDAArray<ValueType>, FormatOrderOne<ValueType>> {
 public:
  CUDAArray(DimensionType nnz, ValueType *vals, context::CUDAContext context,
            Ownership own = kOwned) {
    this->order_ = 1;
    this->dimension_ = {nnz};
    this->nnz_ = nnz;
    this->context_ = std::make_unique<context::CUDAContext>(context);
    vals_ = device_handle(vals, own, context.device_id);
  }
  // The reference's copy constructor mixes host and device pointers (cuda_array_cuda.cu:20-59);
  // this one is a device-to-device copy.
  CUDAArray(const CUDAArray &rhs)
      : CUDAArray(rhs.nnz_,
                  device_copy(rhs.cuda_device(), rhs.cuda_device(), rhs.vals_.get(), rhs.nnz_),
                  *rhs.get_cuda_context(), kOwned) {}
  Format *Clone() const override { return new CUDAArray(*this); }
  context::CUDAContext *get_cuda_context() const {
    return static_cast<context::CUDAContext *>(this->get_context());
  }
  ValueType *get_vals() const { return vals_.get(); }
  ValueType *release_vals() { return release_handle(vals_); }
  void set_vals(ValueType *p, Ownership own = kNotOwned) {
    vals_ = device_handle(p, own, this->cuda_device());
  }

 protected:
  ArrayHandle<ValueType> vals_;
};

}  // namespace format

// =====================================================================================
// Conversion functions and their registration
// =====================================================================================
namespace converter {

// ---- edge conditions ----
inline bool IsCUDA(context::Context *c) {
  return c && c->get_id() == context::CUDAContext::get_id_static();
}
inline bool IsCPU(context::Context *c) {
  return c && c->get_id() == context::CPUContext::get_id_static();
}
inline bool CPUToCUDA(context::Context *from, context::Context *to) { return IsCPU(from) && IsCUDA(to); }
inline bool CUDAToCPU(context::Context *from, context::Context *to) { return IsCUDA(from) && IsCPU(to); }
inline bool SameCUDADevice(context::Context *from, context::Context *to) {
  return IsCUDA(from) && IsCUDA(to) && from->IsEquivalent(to);
}
// converter_cuda.cu:12-21
inline bool CUDAPeerToPeer(context::Context *from, context::Context *to) {
  if (!IsCUDA(from) || !IsCUDA(to)) return false;
  int can = 0;
  sb200_can_access_peer(static_cast<context::CUDAContext *>(to)->device_id,
                        static_cast<context::CUDAContext *>(from)->device_id, &can);
  return can != 0;
}

namespace fn {
using namespace format;

inline int device_of(context::Context *c) { return static_cast<context::CUDAContext *>(c)->device_id; }

// ---- transfers: converter_order_two_cuda.cu:11-105 (every return code checked, m kept) ----
template <typename I, typename N, typename V>
Format *CsrCUDACsr(Format *source, context::Context *to) {
  auto *csr = source->AsAbsolute<CSR<I, N, V>>();
  auto *ctx = static_cast<context::CUDAContext *>(to);
  const int dev = ctx->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  return new CUDACSR<I, N, V>((I)dims[0], (I)dims[1], (N)nnz,
                              sb200::upload(dev, csr->get_row_ptr(), dims[0] + 1),
                              sb200::upload(dev, csr->get_col(), nnz),
                              sb200::upload(dev, csr->get_vals(), nnz), *ctx, kOwned);
}
template <typename I, typename N, typename V>
Format *CUDACsrCsr(Format *source, context::Context *) {
  auto *d = source->AsAbsolute<CUDACSR<I, N, V>>();
  const int dev = d->get_cuda_context()->device_id;
  const auto dims = d->get_dimensions();
  const size_t nnz = d->get_num_nnz();
  return new CSR<I, N, V>((I)dims[0], (I)dims[1], sb200::download(dev, d->get_row_ptr(), dims[0] + 1),
                          sb200::download(dev, d->get_col(), nnz),
                          sb200::download(dev, d->get_vals(), nnz), kOwned, /*ignore_sort=*/true);
}
template <typename I, typename N, typename V>
Format *CUDACsrCUDACsr(Format *source, context::Context *to) {
  auto *d = source->AsAbsolute<CUDACSR<I, N, V>>();
  auto *ctx = static_cast<context::CUDAContext *>(to);
  const int src = d->get_cuda_context()->device_id, dst = ctx->device_id;
  sb200::check(sb200_enable_peer_access(dst, src), dst);
  const auto dims = d->get_dimensions();
  const size_t nnz = d->get_num_nnz();
  return new CUDACSR<I, N, V>((I)dims[0], (I)dims[1], (N)nnz,
                              device_copy(dst, src, d->get_row_ptr(), dims[0] + 1),
                              device_copy(dst, src, d->get_col(), nnz),
                              device_copy(dst, src, d->get_vals(), nnz), *ctx, kOwned);
}
template <typename I, typename N, typename V>
Format *CscCUDACsc(Format *source, context::Context *to) {
  auto *csc = source->AsAbsolute<CSC<I, N, V>>();
  auto *ctx = static_cast<context::CUDAContext *>(to);
  const int dev = ctx->device_id;
  const auto dims = csc->get_dimensions();
  const size_t nnz = csc->get_num_nnz();
  return new CUDACSC<I, N, V>((I)dims[0], (I)dims[1], (N)nnz,
                              sb200::upload(dev, csc->get_col_ptr(), dims[0] + 1),
                              sb200::upload(dev, csc->get_row(), nnz),
                              sb200::upload(dev, csc->get_vals(), nnz), *ctx, kOwned);
}
template <typename I, typename N, typename V>
Format *CUDACscCsc(Format *source, context::Context *) {
  auto *d = source->AsAbsolute<CUDACSC<I, N, V>>();
  const int dev = d->get_cuda_context()->device_id;
  const auto dims = d->get_dimensions();
  const size_t nnz = d->get_num_nnz();
  return new CSC<I, N, V>((I)dims[0], (I)dims[1], sb200::download(dev, d->get_col_ptr(), dims[0] + 1),
                          sb200::download(dev, d->get_row(), nnz),
                          sb200::download(dev, d->get_vals(), nnz), kOwned, /*ignore_sort=*/true);
}
template <typename I, typename N, typename V>
Format *CooCUDACoo(Format *source, context::Context *to) {
  auto *coo = source->AsAbsolute<COO<I, N, V>>();
  auto *ctx = static_cast<context::CUDAContext *>(to);
  const int dev = ctx->device_id;
  const auto dims = coo->get_dimensions();
  const size_t nnz = coo->get_num_nnz();
  // the host object was constructed already (sorted unless its creator said ignore_sort)
  return new CUDACOO<I, N, V>((I)dims[0], (I)dims[1], (N)nnz, sb200::upload(dev, coo->get_row(), nnz),
                              sb200::upload(dev, coo->get_col(), nnz),
                              sb200::upload(dev, coo->get_vals(), nnz), *ctx, kOwned,
                              /*ignore_sort=*/true);
}
template <typename I, typename N, typename V>
Format *CUDACooCoo(Format *source, context::Context *) {
  auto *d = source->AsAbsolute<CUDACOO<I, N, V>>();
  const int dev = d->get_cuda_context()->device_id;
  const auto dims = d->get_dimensions();
  const size_t nnz = d->get_num_nnz();
  return new COO<I, N, V>((I)dims[0], (I)dims[1], (N)nnz, sb200::download(dev, d->get_row(), nnz),
                          sb200::download(dev, d->get_col(), nnz),
                          sb200::download(dev, d->get_vals(), nnz), kOwned, /*ignore_sort=*/true);
}

// ---- the conversions proper: converter_order_two.cc:20-246 on the device ----
template <typename I, typename N, typename V>
Format *CUDACooCUDACsr(Format *source, context::Context *) {
  auto *coo = source->AsAbsolute<CUDACOO<I, N, V>>();
  auto *ctx = coo->get_cuda_context();
  const int dev = ctx->device_id;
  const auto dims = coo->get_dimensions();
  const size_t nnz = coo->get_num_nnz();
  sb200::DeviceScratch<N> row_ptr(dev, dims[0] + 1);
  sb200::DeviceScratch<I> col(dev, nnz);
  sb200::DeviceScratch<V> vals(dev, coo->get_vals() ? nnz : 0);
  sb200::check(sb200_coo_to_csr(dev, dims[0], dims[1], nnz, coo->get_row(), coo->get_col(),
                                coo->get_vals(), row_ptr.get(), col.get(),
                                coo->get_vals() ? vals.get() : nullptr, sb200::dtype_of<I>(),
                                sb200::dtype_of<N>(), sb200::dtype_of<V>(), nullptr),
               dev);
  sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
  V *v = coo->get_vals() ? vals.release() : nullptr;
  return new CUDACSR<I, N, V>((I)dims[0], (I)dims[1], (N)nnz, row_ptr.release(), col.release(), v,
                              *ctx, kOwned);
}
template <typename I, typename N, typename V>
Format *CUDACsrCUDACoo(Format *source, context::Context *) {
  auto *csr = source->AsAbsolute<CUDACSR<I, N, V>>();
  auto *ctx = csr->get_cuda_context();
  const int dev = ctx->device_id;
  const auto dims = csr->get_dimensions();
  const size_t nnz = csr->get_num_nnz();
  sb200::DeviceScratch<I> row(dev, nnz), col(dev, nnz);
  sb200::DeviceScratch<V> vals(dev, csr->get_vals() ? nnz : 0);
  sb200::check(sb200_csr_to_coo(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(),
                                csr->get_vals(), row.get(), col.get(),
                                csr->get_vals() ? vals.get() : nullptr, sb200::dtype_of<I>(),
                                sb200::dtype_of<N>(), sb200::dtype_of<V>(), nullptr),
               dev);
  sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
  V *v = csr->get_vals() ? vals.release() : nullptr;
  // sb200_csr_to_coo already applied the COO-constructor sort
  return new CUDACOO<I, N, V>((I)dims[0], (I)dims[1], (N)nnz, row.release(), col.release(), v, *ctx,
                              kOwned, /*ignore_sort=*/true);
}
template <typename I, typename N, typename V, bool FromCsr>
Format *ToCUDACsc(Format *source, context::Context *) {
  context::CUDAContext *ctx;
  std::vector<DimensionType> dims;
  size_t nnz;
  const void *a, *b;
  const V *sv;
  if constexpr (FromCsr) {
    auto *csr = source->AsAbsolute<CUDACSR<I, N, V>>();
    ctx = csr->get_cuda_context(), dims = csr->get_dimensions(), nnz = csr->get_num_nnz();
    a = csr->get_row_ptr(), b = csr->get_col(), sv = csr->get_vals();
  } else {
    auto *coo = source->AsAbsolute<CUDACOO<I, N, V>>();
    ctx = coo->get_cuda_context(), dims = coo->get_dimensions(), nnz = coo->get_num_nnz();
    a = coo->get_row(), b = coo->get_col(), sv = coo->get_vals();
  }
  const int dev = ctx->device_id;
  sb200::DeviceScratch<N> col_ptr(dev, dims[0] + 1);
  sb200::DeviceScratch<I> row(dev, nnz);
  sb200::DeviceScratch<V> vals(dev, sv ? nnz : 0);
  auto call = FromCsr ? sb200_csr_to_csc : sb200_coo_to_csc;
  sb200::check(call(dev, dims[0], dims[1], nnz, a, b, sv, col_ptr.get(), row.get(),
                    sv ? vals.get() : nullptr, sb200::dtype_of<I>(), sb200::dtype_of<N>(),
                    sb200::dtype_of<V>(), nullptr),
               dev);
  sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
  V *v = sv ? vals.release() : nullptr;
  return new CUDACSC<I, N, V>((I)dims[0], (I)dims[1], (N)nnz, col_ptr.release(), row.release(), v,
                              *ctx, kOwned);
}

// ---- order one: converter_order_one_cuda.cu:10-43 ----
template <typename V>
Format *ArrayCUDAArray(Format *source, context::Context *to) {
  auto *arr = source->AsAbsolute<Array<V>>();
  auto *ctx = static_cast<context::CUDAContext *>(to);
  return new CUDAArray<V>(arr->get_num_nnz(),
                          sb200::upload(ctx->device_id, arr->get_vals(), arr->get_num_nnz()), *ctx,
                          kOwned);
}
template <typename V>
Format *CUDAArrayArray(Format *source, context::Context *) {
  auto *d = source->AsAbsolute<CUDAArray<V>>();
  return new Array<V>(d->get_num_nnz(),
                      sb200::download(d->get_cuda_context()->device_id, d->get_vals(), d->get_num_nnz()),
                      kOwned);
}
}  // namespace fn

// The conversion graph of one <IDType, NNZType, ValueType> triple:
//
//   COO  <-> CUDACOO --> CUDACSR <-> CSR          host <-> device edges are copies,
//             |   ^         |  ^                  device -> device edges are the kernels of
//             v   +---------+  | (peer copy)      libsb200.so; there is NO host -> host edge
//           CUDACSC <-> CSC  CUDACSR'             (no CPU implementation): host conversions
//                                                 are found as chains through a CUDAContext.
template <typename IDType, typename NNZType, typename ValueType>
class ConverterOrderTwo : public ConverterImpl<ConverterOrderTwo<IDType, NNZType, ValueType>> {
 public:
  ConverterOrderTwo() { ResetConverterOrderTwo(); }
  Converter *Clone() const override { return new ConverterOrderTwo(*this); }
  void Reset() override { ResetConverterOrderTwo(); }
  void ResetConverterOrderTwo() {
    using namespace format;
    using I = IDType;
    using N = NNZType;
    using V = ValueType;
    this->ClearConversionFunctions(false);
    this->ClearConversionFunctions(true);
    for (bool mv : {false, true}) {
      auto reg = [&](std::type_index f, std::type_index t, ConversionFunction fn,
                     ConversionCondition c) {
        this->RegisterConversionFunction(f, t, std::move(fn), std::move(c), mv);
      };
      reg(CUDACOO<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(),
          fn::CUDACooCUDACsr<I, N, V>, SameCUDADevice);
      reg(CUDACSR<I, N, V>::get_id_static(), CUDACOO<I, N, V>::get_id_static(),
          fn::CUDACsrCUDACoo<I, N, V>, SameCUDADevice);
      reg(CUDACOO<I, N, V>::get_id_static(), CUDACSC<I, N, V>::get_id_static(),
          fn::ToCUDACsc<I, N, V, false>, SameCUDADevice);
      reg(CUDACSR<I, N, V>::get_id_static(), CUDACSC<I, N, V>::get_id_static(),
          fn::ToCUDACsc<I, N, V, true>, SameCUDADevice);
      reg(CSR<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(), fn::CsrCUDACsr<I, N, V>,
          CPUToCUDA);
      reg(CUDACSR<I, N, V>::get_id_static(), CSR<I, N, V>::get_id_static(), fn::CUDACsrCsr<I, N, V>,
          CUDAToCPU);
      reg(COO<I, N, V>::get_id_static(), CUDACOO<I, N, V>::get_id_static(), fn::CooCUDACoo<I, N, V>,
          CPUToCUDA);
      reg(CUDACOO<I, N, V>::get_id_static(), COO<I, N, V>::get_id_static(), fn::CUDACooCoo<I, N, V>,
          CUDAToCPU);
      reg(CSC<I, N, V>::get_id_static(), CUDACSC<I, N, V>::get_id_static(), fn::CscCUDACsc<I, N, V>,
          CPUToCUDA);
      reg(CUDACSC<I, N, V>::get_id_static(), CSC<I, N, V>::get_id_static(), fn::CUDACscCsc<I, N, V>,
          CUDAToCPU);
      reg(CUDACSR<I, N, V>::get_id_static(), CUDACSR<I, N, V>::get_id_static(),
          fn::CUDACsrCUDACsr<I, N, V>, CUDAPeerToPeer);
    }
  }
};

template <typename ValueType>
class ConverterOrderOne : public ConverterImpl<ConverterOrderOne<ValueType>> {
 public:
  ConverterOrderOne() { ResetConverterOrderOne(); }
  Converter *Clone() const override { return new ConverterOrderOne(*this); }
  void Reset() override { ResetConverterOrderOne(); }
  void ResetConverterOrderOne() {
    using namespace format;
    this->ClearConversionFunctions(false);
    this->ClearConversionFunctions(true);
    for (bool mv : {false, true}) {
      this->RegisterConversionFunction(Array<ValueType>::get_id_static(),
                                       CUDAArray<ValueType>::get_id_static(),
                                       fn::ArrayCUDAArray<ValueType>, CPUToCUDA, mv);
      this->RegisterConversionFunction(CUDAArray<ValueType>::get_id_static(),
                                       Array<ValueType>::get_id_static(),
                                       fn::CUDAArrayArray<ValueType>, CUDAToCPU, mv);
    }
  }
};

}  // namespace converter
}  // namespace sparsebase
