// sb200/ops.h -- the preprocessing operators of the hot path: reorderings, permutations,
// degree features and the ReorderBase facade.  Every implementation function is keyed on a
// DEVICE format (CUDACSR / CUDAArray) and ends in one C-ABI call; host inputs reach them
// through the converter when the caller lists a CUDAContext and allows conversion -- the
// registration pattern the reference uses for its only CUDA operator
// (feature/jaccard_weights.cc:17-36).
//
// Interfaces mirrored (reference paths relative to src/sparsebase/):
//   reorder::Reorderer<ID>                 reorder/reorderer.h:36-119, reorderer.cc:21-52
//   reorder::DegreeReorder (+Params)       reorder/degree_reorder.h:15-44, degree_reorder.cc:13-62
//   reorder::RCMReorder (+Params)          reorder/rcm_reorder.h:14-45, rcm_reorder.cc:9-166
//   permute::Permuter<In,Out>              permute/permuter.h:23-102
//   permute::PermuteOrderTwo (+Params)     permute/permute_order_two.h:16-50, permute_order_two.cc:8-79
//   permute::PermuteOrderOne (+Params)     permute/permute_order_one.h, permute_order_one.cc:7-37
//   feature::DegreeDistribution            feature/degree_distribution.h, degree_distribution.cc:110-162
//   feature::Degrees                       feature/degrees.h, degrees.cc:60-105
//   bases::ReorderBase                     bases/reorder_base.h:29-707 (Reorder, Permute2D*,
//                                          Permute1D, InversePermutation)
// Result conventions are the reference's: reorderers and features return host arrays from
// new[] (caller delete[]s), inv[old] = new; permuters return a new Format owned by the caller.
#pragma once
#include "matcher.h"

namespace sparsebase {

// ======================================================================= reorder
namespace reorder {

template <typename IDType>
class Reorderer : public utils::FunctionMatcherMixin<IDType *> {
 public:
  IDType *GetReorder(format::Format *format, std::vector<context::Context *> contexts,
                     bool convert_input) {
    return this->Execute(this->params_.get(), contexts, convert_input, format);
  }
  IDType *GetReorder(format::Format *format, utils::Parameters *params,
                     std::vector<context::Context *> contexts, bool convert_input) {
    return this->Execute(params, contexts, convert_input, format);
  }
  std::tuple<std::vector<std::vector<format::Format *>>, IDType *> GetReorderCached(
      format::Format *format, std::vector<context::Context *> contexts, bool convert_input) {
    return this->CachedExecute(this->params_.get(), contexts, convert_input, false, format);
  }
  std::tuple<std::vector<std::vector<format::Format *>>, IDType *> GetReorderCached(
      format::Format *format, utils::Parameters *params, std::vector<context::Context *> contexts,
      bool convert_input) {
    return this->CachedExecute(params, contexts, convert_input, false, format);
  }
  virtual ~Reorderer() = default;
};

struct DegreeReorderParams : utils::Parameters {
  bool ascending;
  DegreeReorderParams(bool ascending) : ascending(ascending) {}
};

template <typename IDType, typename NNZType, typename ValueType>
class DegreeReorder : public Reorderer<IDType> {
 public:
  typedef DegreeReorderParams ParamsType;
  DegreeReorder(bool ascending) : DegreeReorder(DegreeReorderParams(ascending)) {}
  DegreeReorder(DegreeReorderParams params) {
    this->RegisterFunction({format::CUDACSR<IDType, NNZType, ValueType>::get_id_static()},
                           CalculateReorderCUDACSR);
    this->params_ = std::make_unique<DegreeReorderParams>(params);
  }

 protected:
  // degree_reorder.cc:22-62 -> sb200_degree_reorder
  static IDType *CalculateReorderCUDACSR(std::vector<format::Format *> formats,
                                         utils::Parameters *params) {
    auto *csr = formats[0]->AsAbsolute<format::CUDACSR<IDType, NNZType, ValueType>>();
    auto *p = static_cast<DegreeReorderParams *>(params);
    const int dev = csr->get_cuda_context()->device_id;
    const size_t n = csr->get_dimensions()[0];
    sb200::DeviceScratch<IDType> inv(dev, n);
    sb200::check(sb200_degree_reorder(dev, n, csr->get_row_ptr(), p->ascending ? 1 : 0, inv.get(),
                                      sb200::dtype_of<IDType>(), sb200::dtype_of<NNZType>(),
                                      nullptr),
                 dev);
    return sb200::download(dev, inv.get(), n);
  }
};

struct RCMReorderParams : utils::Parameters {};

template <typename IDType, typename NNZType, typename ValueType>
class RCMReorder : public Reorderer<IDType> {
 public:
  typedef RCMReorderParams ParamsType;
  RCMReorder() : RCMReorder(RCMReorderParams{}) {}
  RCMReorder(RCMReorderParams params) {
    this->RegisterFunction({format::CUDACSR<IDType, NNZType, ValueType>::get_id_static()},
                           GetReorderCUDACSR);
    this->params_ = std::make_unique<RCMReorderParams>(params);
  }

 protected:
  // rcm_reorder.cc:22-166 -> sb200_rcm_reorder
  static IDType *GetReorderCUDACSR(std::vector<format::Format *> formats, utils::Parameters *) {
    auto *csr = formats[0]->AsAbsolute<format::CUDACSR<IDType, NNZType, ValueType>>();
    const int dev = csr->get_cuda_context()->device_id;
    const size_t n = csr->get_dimensions()[0];
    sb200::DeviceScratch<IDType> inv(dev, n);
    sb200::check(sb200_rcm_reorder(dev, n, csr->get_num_nnz(), csr->get_row_ptr(), csr->get_col(),
                                   inv.get(), sb200::dtype_of<IDType>(),
                                   sb200::dtype_of<NNZType>(), nullptr),
                 dev);
    return sb200::download(dev, inv.get(), n);
  }
};

}  // namespace reorder

// ======================================================================= permute
namespace permute {

template <typename InputFormatType, typename ReturnFormatType>
class Permuter : public utils::FunctionMatcherMixin<ReturnFormatType *> {
 public:
  Permuter() {
    static_assert(std::is_base_of<format::Format, InputFormatType>::value,
                  "Permuter must take as input a Format object");
    static_assert(std::is_base_of<format::Format, ReturnFormatType>::value,
                  "Permuter must return a Format object");
  }
  ReturnFormatType *GetPermutation(format::Format *format, std::vector<context::Context *> contexts,
                                   bool convert_input) {
    return this->Execute(this->params_.get(), contexts, convert_input, format);
  }
  ReturnFormatType *GetPermutation(format::Format *format, utils::Parameters *params,
                                   std::vector<context::Context *> contexts, bool convert_input) {
    return this->Execute(params, contexts, convert_input, format);
  }
  std::tuple<std::vector<std::vector<format::Format *>>, ReturnFormatType *> GetPermutationCached(
      format::Format *format, std::vector<context::Context *> contexts, bool convert_input) {
    return this->CachedExecute(this->params_.get(), contexts, convert_input, false, format);
  }
  std::tuple<std::vector<std::vector<format::Format *>>, ReturnFormatType *> GetPermutationCached(
      format::Format *format, utils::Parameters *params, std::vector<context::Context *> contexts,
      bool convert_input) {
    return this->CachedExecute(params, contexts, convert_input, false, format);
  }
  virtual ~Permuter() = default;
};

// row_order / col_order are inverse permutations (inv[old] = new) in HOST memory, or nullptr
// for the identity -- exactly what the reorderers return.
template <typename IDType>
struct PermuteOrderTwoParams : utils::Parameters {
  IDType *row_order;
  IDType *col_order;
  explicit PermuteOrderTwoParams(IDType *r, IDType *c) : row_order(r), col_order(c) {}
};

template <typename IDType, typename NNZType, typename ValueType>
class PermuteOrderTwo : public Permuter<format::FormatOrderTwo<IDType, NNZType, ValueType>,
                                        format::FormatOrderTwo<IDType, NNZType, ValueType>> {
 public:
  typedef PermuteOrderTwoParams<IDType> ParamsType;
  PermuteOrderTwo(IDType *row_order, IDType *col_order)
      : PermuteOrderTwo(ParamsType(row_order, col_order)) {}
  // (the reference's params-struct constructor is a no-op, permute_order_two.cc:17-20; this
  //  one registers the function like the pointer constructor does)
  explicit PermuteOrderTwo(ParamsType params) {
    this->RegisterFunction({format::CUDACSR<IDType, NNZType, ValueType>::get_id_static()},
                           PermuteOrderTwoCUDACSR);
    this->params_ = std::make_unique<ParamsType>(params);
  }

 protected:
  // permute_order_two.cc:21-79 (+ the CSR-constructor row sort it triggers) -> sb200_permute2d
  static format::FormatOrderTwo<IDType, NNZType, ValueType> *PermuteOrderTwoCUDACSR(
      std::vector<format::Format *> formats, utils::Parameters *params) {
    auto *csr = formats[0]->AsAbsolute<format::CUDACSR<IDType, NNZType, ValueType>>();
    auto *p = static_cast<ParamsType *>(params);
    auto *ctx = csr->get_cuda_context();
    const int dev = ctx->device_id;
    const auto dims = csr->get_dimensions();
    const size_t nnz = csr->get_num_nnz();
    sb200::DeviceScratch<IDType> d_row(dev, p->row_order, dims[0]);
    sb200::DeviceScratch<IDType> d_col(dev, p->col_order, dims[1]);
    sb200::DeviceScratch<NNZType> o_ptr(dev, dims[0] + 1);
    sb200::DeviceScratch<IDType> o_col(dev, nnz);
    sb200::DeviceScratch<ValueType> o_val(dev, csr->get_vals() ? nnz : 0);
    sb200::check(sb200_permute2d(dev, dims[0], dims[1], nnz, csr->get_row_ptr(), csr->get_col(),
                                 csr->get_vals(), d_row.get(), d_col.get(), o_ptr.get(),
                                 o_col.get(), csr->get_vals() ? o_val.get() : nullptr,
                                 sb200::dtype_of<IDType>(), sb200::dtype_of<NNZType>(),
                                 sb200::dtype_of<ValueType>(), nullptr),
                 dev);
    sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
    ValueType *v = csr->get_vals() ? o_val.release() : nullptr;
    return new format::CUDACSR<IDType, NNZType, ValueType>((IDType)dims[0], (IDType)dims[1],
                                                          (NNZType)nnz, o_ptr.release(),
                                                          o_col.release(), v, *ctx, format::kOwned);
  }
};

template <typename IDType>
struct PermuteOrderOneParams : utils::Parameters {
  IDType *order;
  explicit PermuteOrderOneParams(IDType *order) : order(order) {}
};

template <typename IDType, typename ValueType>
class PermuteOrderOne
    : public Permuter<format::FormatOrderOne<ValueType>, format::FormatOrderOne<ValueType>> {
 public:
  typedef PermuteOrderOneParams<IDType> ParamsType;
  PermuteOrderOne(IDType *order) : PermuteOrderOne(ParamsType(order)) {}
  explicit PermuteOrderOne(ParamsType params) {
    this->RegisterFunction({format::CUDAArray<ValueType>::get_id_static()}, PermuteCUDAArray);
    this->params_ = std::make_unique<ParamsType>(params);
  }

 protected:
  // permute_order_one.cc:17-37 -> sb200_permute1d
  static format::FormatOrderOne<ValueType> *PermuteCUDAArray(std::vector<format::Format *> formats,
                                                             utils::Parameters *params) {
    auto *arr = formats[0]->AsAbsolute<format::CUDAArray<ValueType>>();
    auto *p = static_cast<ParamsType *>(params);
    auto *ctx = arr->get_cuda_context();
    const int dev = ctx->device_id;
    const size_t len = arr->get_num_nnz();
    sb200::DeviceScratch<IDType> d_order(dev, p->order, len);
    sb200::DeviceScratch<ValueType> out(dev, len);
    sb200::check(sb200_permute1d(dev, len, arr->get_vals(), d_order.get(), out.get(),
                                 sb200::dtype_of<IDType>(), sb200::dtype_of<ValueType>(), nullptr),
                 dev);
    sb200::check(sb200_stream_synchronize(dev, nullptr), dev);
    return new format::CUDAArray<ValueType>(len, out.release(), *ctx, format::kOwned);
  }
};

}  // namespace permute

// ======================================================================= feature
namespace feature {

template <typename ReturnType>
class FeaturePreprocessType : public utils::FunctionMatcherMixin<ReturnType> {
 public:
  virtual ~FeaturePreprocessType() = default;
};

struct DegreeDistributionParams : utils::Parameters {};

template <typename IDType, typename NNZType, typename ValueType, typename FeatureType>
class DegreeDistribution : public FeaturePreprocessType<FeatureType *> {
 public:
  typedef DegreeDistributionParams ParamsType;
  DegreeDistribution() {
    this->RegisterFunction({format::CUDACSR<IDType, NNZType, ValueType>::get_id_static()},
                           GetDegreeDistributionCUDACSR);
    this->params_ = std::make_unique<ParamsType>();
  }
  explicit DegreeDistribution(DegreeDistributionParams) : DegreeDistribution() {}
  FeatureType *GetDistribution(format::Format *format, std::vector<context::Context *> contexts,
                               bool convert_input) {
    return this->Execute(this->params_.get(), contexts, convert_input, format);
  }
  std::tuple<std::vector<std::vector<format::Format *>>, FeatureType *> GetDistributionCached(
      format::Format *format, std::vector<context::Context *> contexts, bool convert_input) {
    return this->CachedExecute(this->params_.get(), contexts, convert_input, false, format);
  }

 protected:
  // degree_distribution.cc:146-162 -> sb200_degree_distribution
  static FeatureType *GetDegreeDistributionCUDACSR(std::vector<format::Format *> formats,
                                                   utils::Parameters *) {
    static_assert(std::is_same_v<FeatureType, float> || std::is_same_v<FeatureType, double>,
                  "FeatureType must be float or double");
    auto *csr = formats[0]->AsAbsolute<format::CUDACSR<IDType, NNZType, ValueType>>();
    const int dev = csr->get_cuda_context()->device_id;
    const size_t n = csr->get_dimensions()[0];
    sb200::DeviceScratch<FeatureType> dist(dev, n);
    sb200::check(sb200_degree_distribution(dev, n, csr->get_num_nnz(), csr->get_row_ptr(),
                                           dist.get(), sb200::dtype_of<NNZType>(),
                                           sb200::dtype_of<FeatureType>(), nullptr),
                 dev);
    return sb200::download(dev, dist.get(), n);
  }
};

struct DegreesParams : utils::Parameters {};

template <typename IDType, typename NNZType, typename ValueType>
class Degrees : public FeaturePreprocessType<IDType *> {
 public:
  typedef DegreesParams ParamsType;
  Degrees() {
    this->RegisterFunction({format::CUDACSR<IDType, NNZType, ValueType>::get_id_static()},
                           GetDegreesCUDACSR);
    this->params_ = std::make_unique<ParamsType>();
  }
  explicit Degrees(DegreesParams) : Degrees() {}
  IDType *GetDegrees(format::Format *format, std::vector<context::Context *> contexts,
                     bool convert_input) {
    return this->Execute(this->params_.get(), contexts, convert_input, format);
  }

 protected:
  // degrees.cc:93-105 -> sb200_degrees
  static IDType *GetDegreesCUDACSR(std::vector<format::Format *> formats, utils::Parameters *) {
    auto *csr = formats[0]->AsAbsolute<format::CUDACSR<IDType, NNZType, ValueType>>();
    const int dev = csr->get_cuda_context()->device_id;
    const size_t n = csr->get_dimensions()[0];
    sb200::DeviceScratch<IDType> deg(dev, n);
    sb200::check(sb200_degrees(dev, n, csr->get_row_ptr(), deg.get(), sb200::dtype_of<IDType>(),
                               sb200::dtype_of<NNZType>(), nullptr),
                 dev);
    return sb200::download(dev, deg.get(), n);
  }
};

}  // namespace feature

// ======================================================================= bases
namespace bases {

class ReorderBase {
 public:
  // reorder_base.h:48-68
  template <template <typename, typename, typename> typename Reordering, typename AutoIDType,
            typename AutoNNZType, typename AutoValueType>
  static AutoIDType *Reorder(
      typename Reordering<AutoIDType, AutoNNZType, AutoValueType>::ParamsType params,
      format::FormatOrderTwo<AutoIDType, AutoNNZType, AutoValueType> *format,
      std::vector<context::Context *> contexts, bool convert_input) {
    static_assert(std::is_base_of_v<reorder::Reorderer<AutoIDType>,
                                    Reordering<AutoIDType, AutoNNZType, AutoValueType>>,
                  "You must pass a reordering function (with base Reorderer) to "
                  "ReorderBase::Reorder");
    Reordering<AutoIDType, AutoNNZType, AutoValueType> reordering(params);
    return reordering.GetReorder(format, contexts, convert_input);
  }

  // reorder_base.h:142-166 -- one order for rows and columns
  template <template <typename, typename, typename> typename ReturnFormatType = format::FormatOrderTwo,
            typename AutoIDType, typename AutoNNZType, typename AutoValueType>
  static ReturnFormatType<AutoIDType, AutoNNZType, AutoValueType> *Permute2D(
      AutoIDType *ordering, format::FormatOrderTwo<AutoIDType, AutoNNZType, AutoValueType> *format,
      std::vector<context::Context *> contexts, bool convert_input, bool convert_output = false) {
    return Permute2DRowColumnWise<ReturnFormatType>(ordering, ordering, format, contexts,
                                                    convert_input, convert_output);
  }

  // reorder_base.h:313-338
  template <template <typename, typename, typename> typename ReturnFormatType = format::FormatOrderTwo,
            typename AutoIDType, typename AutoNNZType, typename AutoValueType>
  static ReturnFormatType<AutoIDType, AutoNNZType, AutoValueType> *Permute2DRowColumnWise(
      AutoIDType *row_ordering, AutoIDType *col_ordering,
      format::FormatOrderTwo<AutoIDType, AutoNNZType, AutoValueType> *format,
      std::vector<context::Context *> contexts, bool convert_input, bool convert_output = false) {
    permute::PermuteOrderTwo<AutoIDType, AutoNNZType, AutoValueType> perm(row_ordering,
                                                                          col_ordering);
    auto *out = perm.GetPermutation(format, contexts, convert_input);
    return Finish2D<ReturnFormatType>(out, convert_output);
  }
  // reorder_base.h:359-384
  template <template <typename, typename, typename> typename ReturnFormatType = format::FormatOrderTwo,
            typename AutoIDType, typename AutoNNZType, typename AutoValueType>
  static ReturnFormatType<AutoIDType, AutoNNZType, AutoValueType> *Permute2DRowWise(
      AutoIDType *ordering, format::FormatOrderTwo<AutoIDType, AutoNNZType, AutoValueType> *format,
      std::vector<context::Context *> contexts, bool convert_input, bool convert_output = false) {
    return Permute2DRowColumnWise<ReturnFormatType>(ordering, (AutoIDType *)nullptr, format,
                                                    contexts, convert_input, convert_output);
  }
  // reorder_base.h:449-474
  template <template <typename, typename, typename> typename ReturnFormatType = format::FormatOrderTwo,
            typename AutoIDType, typename AutoNNZType, typename AutoValueType>
  static ReturnFormatType<AutoIDType, AutoNNZType, AutoValueType> *Permute2DColWise(
      AutoIDType *ordering, format::FormatOrderTwo<AutoIDType, AutoNNZType, AutoValueType> *format,
      std::vector<context::Context *> contexts, bool convert_input, bool convert_output = false) {
    return Permute2DRowColumnWise<ReturnFormatType>((AutoIDType *)nullptr, ordering, format,
                                                    contexts, convert_input, convert_output);
  }

  // reorder_base.h:576-598
  template <template <typename> typename ReturnFormatType = format::FormatOrderOne,
            typename AutoIDType, typename AutoValueType>
  static ReturnFormatType<AutoValueType> *Permute1D(AutoIDType *ordering,
                                                    format::FormatOrderOne<AutoValueType> *format,
                                                    std::vector<context::Context *> contexts,
                                                    bool convert_inputs,
                                                    bool convert_output = false) {
    permute::PermuteOrderOne<AutoIDType, AutoValueType> perm(ordering);
    auto *out = perm.GetPermutation(format, contexts, convert_inputs);
    if constexpr (std::is_same_v<ReturnFormatType<AutoValueType>,
                                 format::FormatOrderOne<AutoValueType>>) {
      return out;
    } else {
      if (!convert_output) return out->template As<ReturnFormatType>();
      if (out->template Is<ReturnFormatType>()) return out->template As<ReturnFormatType>();
      auto *converted = out->template Convert<ReturnFormatType>();
      delete out;
      return converted;
    }
  }

  // reorder_base.h:662-671: inv_perm[perm[i]] = i, host arrays in and out.  Staged through
  // the default device (sb200_inverse_permutation) -- no CPU implementation here either.
  template <typename AutoIDType, typename AutoNumType>
  static AutoIDType *InversePermutation(AutoIDType *perm, AutoNumType length) {
    static_assert(std::is_integral_v<AutoNumType>,
                  "Length of the permutation array must be an integer");
    const int dev = sb200::default_device();
    sb200::DeviceScratch<AutoIDType> d_perm(dev, perm, (size_t)length);
    sb200::DeviceScratch<AutoIDType> d_inv(dev, (size_t)length);
    sb200::check(sb200_inverse_permutation(dev, (int64_t)length, d_perm.get(), d_inv.get(),
                                           sb200::dtype_of<AutoIDType>(), nullptr),
                 dev);
    return sb200::download(dev, d_inv.get(), (size_t)length);
  }

 private:
  template <template <typename, typename, typename> typename ReturnFormatType, typename I,
            typename N, typename V>
  static ReturnFormatType<I, N, V> *Finish2D(format::FormatOrderTwo<I, N, V> *out,
                                             bool convert_output) {
    if constexpr (std::is_same_v<ReturnFormatType<I, N, V>, format::FormatOrderTwo<I, N, V>>) {
      return out;
    } else {
      if (!convert_output) return out->template As<ReturnFormatType>();
      if (out->template Is<ReturnFormatType>()) return out->template As<ReturnFormatType>();
      auto *converted = out->template Convert<ReturnFormatType>();
      delete out;
      return converted;
    }
  }
};

}  // namespace bases
}  // namespace sparsebase
