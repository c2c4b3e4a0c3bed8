// host_api_test.cc -- exercises the C++ host layer (sparsebase_b200/host/include) the way the
// reference's own unit tests exercise the same interfaces, with the reference's golden vectors:
//
//   converter/common.inc:5-16 + converter_order_two_tests.cc:9-351     (COO/CSR/CSC conversions)
//   format/csr_tests.cc:80-115, format/coo_tests.cc:77-115             (constructor sorts)
//   converter_order_two_cuda_tests.cu:11-49, converter_order_one_cuda_tests.cu:15-106
//   functionality_common.inc:6-98 + permute_order_two_tests.cc:27-91,
//   permute_order_one_tests.cc:25-51, bases/reorder_base_tests.cc:151-430
//   reorder/degree_reorder_tests.cc:28-81, reorder/rcm_reorder_tests.cc:21-25
//   feature/degree_distribution_tests.cc:38-119, feature/degrees_tests.cc
//
// The include paths and class names below are the reference's; only the context passed to
// the operators is a CUDAContext, because this implementation has no CPU path.
//
//   host_api_test              run everything (needs a GPU)
//   host_api_test --no-device  only the checks that must hold without a GPU
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>
#include <set>
#include <string>

#include "sparsebase/bases/reorder_base.h"
#include "sparsebase/context/cpu_context.h"
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
#include "sparsebase/reorder/degree_reorder.h"
#include "sparsebase/reorder/rcm_reorder.h"

using namespace sparsebase;

// ------------------------------------------------------------------ a very small test harness
struct TestCase {
  const char *name;
  bool needs_device;
  std::function<void()> body;
};
static std::vector<TestCase> &registry() {
  static std::vector<TestCase> r;
  return r;
}
struct Registrar {
  Registrar(const char *name, bool dev, std::function<void()> body) {
    registry().push_back({name, dev, std::move(body)});
  }
};
static int g_failures = 0;
#define TEST_DEVICE(name) \
  static void name();     \
  static Registrar reg_##name(#name, true, name); \
  static void name()
#define TEST_HOST(name) \
  static void name();   \
  static Registrar reg_##name(#name, false, name); \
  static void name()
#define EXPECT_TRUE(cond)                                                      \
  do {                                                                         \
    if (!(cond)) {                                                             \
      std::printf("    FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      g_failures++;                                                            \
    }                                                                          \
  } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_THROW(stmt, ExType)                                                   \
  do {                                                                               \
    bool thrown__ = false;                                                           \
    try {                                                                            \
      stmt;                                                                          \
    } catch (const ExType &) {                                                       \
      thrown__ = true;                                                               \
    } catch (...) {                                                                  \
    }                                                                                \
    if (!thrown__) {                                                                 \
      std::printf("    FAILED %s:%d: expected %s from %s\n", __FILE__, __LINE__, #ExType, #stmt); \
      g_failures++;                                                                  \
    }                                                                                \
  } while (0)

template <typename T, typename U>
static bool same(const T *got, std::initializer_list<U> want) {
  size_t i = 0;
  for (const U &w : want)
    if (!(got[i++] == (T)w)) return false;
  return true;
}

// ------------------------------------------------------------------ golden vectors
// converter/common.inc:5-16 (12 x 9, 7 nonzeros, "converted using scipy")
static const int kN = 12, kM = 9, kNnz = 7;
#define COO_ROW {0, 0, 1, 3, 5, 10, 11}
#define COO_COL {0, 2, 1, 3, 3, 8, 7}
#define COO_VALS {3, 5, 7, 9, 15, 11, 13}
#define CSR_ROW_PTR {0, 2, 3, 3, 4, 4, 5, 5, 5, 5, 5, 6, 7}
#define CSC_COL_PTR {0, 1, 2, 3, 5, 5, 5, 5, 6, 7, 7, 7, 7}
#define CSC_ROW {0, 1, 0, 3, 5, 11, 10}
#define CSC_VALS {3, 7, 5, 9, 15, 13, 11}

static context::CPUContext cpu_context;
static context::CUDAContext &gpu() {
  static context::CUDAContext g(0);
  return g;
}
static std::vector<context::Context *> cpu_and_gpu() { return {&cpu_context, &gpu()}; }

// ================================================================== without a device
TEST_HOST(TypeIdentity) {
  using csr_t = format::CSR<int, int, int>;
  EXPECT_TRUE(csr_t::get_id_static() == std::type_index(typeid(csr_t)));
  EXPECT_TRUE((format::CSR<int, int, float>::get_id_static() != csr_t::get_id_static()));
  EXPECT_TRUE((format::CUDACSR<int, int, int>::get_id_static() != csr_t::get_id_static()));
  EXPECT_TRUE(csr_t::get_name_static().find("CSR") != std::string::npos);
  context::CPUContext other;
  EXPECT_TRUE(cpu_context.IsEquivalent(&other));
}

TEST_HOST(ContainersWithIgnoreSort) {
  int row_ptr[] = CSR_ROW_PTR, col[] = COO_COL, vals[] = COO_VALS;
  format::CSR<int, int, int> csr(kN, kM, row_ptr, col, vals, format::kNotOwned, true);
  EXPECT_EQ(csr.get_num_nnz(), (format::DimensionType)kNnz);
  EXPECT_EQ(csr.get_dimensions()[0], (format::DimensionType)kN);
  EXPECT_EQ(csr.get_dimensions()[1], (format::DimensionType)kM);
  EXPECT_EQ(csr.get_order(), (format::DimensionType)2);
  EXPECT_TRUE(csr.get_context()->get_id() == context::CPUContext::get_id_static());
  format::Format *f = &csr;
  EXPECT_TRUE((f->AsAbsolute<format::CSR<int, int, int>>() == &csr));
  EXPECT_THROW((f->AsAbsolute<format::COO<int, int, int>>()), utils::TypeException);
  EXPECT_TRUE(csr.Is<format::CSR>());
  EXPECT_TRUE(!csr.Is<format::CSC>());
  EXPECT_THROW(csr.As<format::CSC>(), utils::TypeException);
  // Clone is a deep copy
  auto *clone = csr.Clone()->AsAbsolute<format::CSR<int, int, int>>();
  EXPECT_TRUE(clone->get_col() != csr.get_col());
  EXPECT_TRUE(std::memcmp(clone->get_col(), col, sizeof(col)) == 0);
  delete clone;
  // release_* hands the array out; the format keeps pointing at it
  int *own = new int[3]{1, 2, 3};
  format::Array<int> arr(3, own, format::kOwned);
  int *back = arr.release_vals();
  EXPECT_TRUE(back == own && arr.get_vals() == own);
  delete[] back;
}

TEST_HOST(NoHostToHostConversion) {
  // There is no CPU implementation: with only a CPU context no chain exists (the reference
  // would run CooCsrFunctionConditional here).
  int row[] = COO_ROW, col[] = COO_COL, vals[] = COO_VALS;
  format::COO<int, int, int> coo(kN, kM, kNnz, row, col, vals, format::kNotOwned, true);
  EXPECT_THROW(coo.Convert<format::CSR>(&cpu_context), utils::ConversionException);
  auto conv = coo.get_converter();
  EXPECT_TRUE(!conv->CanConvert(coo.get_id(), coo.get_context(),
                                format::CSR<int, int, int>::get_id_static(), &cpu_context));
  // same type, same context: the source itself comes back (converter.cc:86-92)
  EXPECT_TRUE(coo.Convert<format::COO>(&cpu_context) == &coo);
}

TEST_HOST(ConverterStoreSharesOneInstance) {
  int row[] = COO_ROW, col[] = COO_COL;
  format::COO<int, int, void> a(kN, kM, kNnz, row, col, nullptr, format::kNotOwned, true);
  format::COO<int, int, void> b(kN, kM, kNnz, row, col, nullptr, format::kNotOwned, true);
  EXPECT_TRUE(a.get_converter().get() == b.get_converter().get());
  format::COO<int, int, float> c(kN, kM, kNnz, row, col, nullptr, format::kNotOwned, true);
  EXPECT_TRUE((const void *)a.get_converter().get() != (const void *)c.get_converter().get());
}

TEST_HOST(FunctionNotFoundWithoutCudaContext) {
  int row_ptr[] = {0, 2, 3, 4}, cols[] = {1, 2, 0, 0}, vals[] = {1, 2, 3, 4};
  format::CSR<int, int, int> csr(3, 3, row_ptr, cols, vals, format::kNotOwned, true);
  reorder::DegreeReorder<int, int, int> reorder(true);
  EXPECT_THROW(reorder.GetReorder(&csr, {&cpu_context}, true), utils::FunctionNotFoundException);
}

TEST_HOST(BadDeviceIdThrows) {
  int cnt = 0;
  sb200_device_count(&cnt);
  EXPECT_THROW(context::CUDAContext bad(cnt), utils::CUDADeviceException);
  EXPECT_THROW(context::CUDAContext neg(-1), utils::CUDADeviceException);
}

// ================================================================== conversions
template <typename I, typename N, typename V>
static void conversion_suite() {
  I row[] = COO_ROW, col[] = COO_COL;
  V vals[] = COO_VALS;
  format::COO<I, N, V> coo(kN, kM, kNnz, row, col, vals, format::kNotOwned);

  // COO -> CSR (converter_order_two_tests.cc:9-60)
  auto *csr = coo.template Convert<format::CSR>(cpu_and_gpu());
  EXPECT_TRUE(same(csr->get_row_ptr(), CSR_ROW_PTR));
  EXPECT_TRUE(same(csr->get_col(), COO_COL));
  EXPECT_TRUE(same(csr->get_vals(), COO_VALS));
  EXPECT_EQ(csr->get_dimensions()[1], (format::DimensionType)kM);
  EXPECT_EQ(csr->get_num_nnz(), (format::DimensionType)kNnz);

  // CSR -> COO (:62-110)
  auto *coo2 = csr->template Convert<format::COO>(cpu_and_gpu());
  EXPECT_TRUE(same(coo2->get_row(), COO_ROW));
  EXPECT_TRUE(same(coo2->get_col(), COO_COL));
  EXPECT_TRUE(same(coo2->get_vals(), COO_VALS));

  // COO -> CSC (:112-162), CSR -> CSC (:164-203); col_ptr has n+1 = 13 entries
  for (format::FormatOrderTwo<I, N, V> *src :
       {(format::FormatOrderTwo<I, N, V> *)&coo, (format::FormatOrderTwo<I, N, V> *)csr}) {
    auto *csc = src->template Convert<format::CSC>(cpu_and_gpu());
    EXPECT_TRUE(same(csc->get_col_ptr(), CSC_COL_PTR));
    EXPECT_TRUE(same(csc->get_row(), CSC_ROW));
    EXPECT_TRUE(same(csc->get_vals(), CSC_VALS));
    delete csc;
  }

  // move conversion gives the same arrays (:205-260)
  auto *csr_mv = coo.template Convert<format::CSR>(cpu_and_gpu(), true);
  EXPECT_TRUE(same(csr_mv->get_row_ptr(), CSR_ROW_PTR));
  EXPECT_TRUE(same(csr_mv->get_vals(), COO_VALS));
  delete csr_mv;

  // cached conversion returns every hop: COO -> CUDACOO -> CUDACSR -> CSR
  auto conv = coo.get_converter();
  auto hops = conv->ConvertCached(&coo, format::CSR<I, N, V>::get_id_static(), cpu_and_gpu());
  EXPECT_EQ(hops.size(), (size_t)3);
  EXPECT_TRUE(hops[0]->get_id() == (format::CUDACOO<I, N, V>::get_id_static()));
  EXPECT_TRUE(hops[1]->get_id() == (format::CUDACSR<I, N, V>::get_id_static()));
  EXPECT_TRUE(hops[2]->get_id() == (format::CSR<I, N, V>::get_id_static()));
  for (auto *h : hops) delete h;

  delete coo2;
  delete csr;
}

TEST_DEVICE(ConverterOrderTwoIntIntInt) { conversion_suite<int, int, int>(); }
TEST_DEVICE(ConverterOrderTwoUnsignedFloat) { conversion_suite<unsigned, unsigned, float>(); }
TEST_DEVICE(ConverterOrderTwoWideTypes) { conversion_suite<long long, long long, double>(); }
TEST_DEVICE(ConverterOrderTwoMixedWidth) { conversion_suite<int, long long, float>(); }

TEST_DEVICE(ConverterOrderTwoVoidValues) {
  int row[] = COO_ROW, col[] = COO_COL;
  format::COO<int, int, void> coo(kN, kM, kNnz, row, col, nullptr);
  auto *csr = coo.Convert<format::CSR>(cpu_and_gpu());
  EXPECT_TRUE(same(csr->get_row_ptr(), CSR_ROW_PTR));
  EXPECT_TRUE(same(csr->get_col(), COO_COL));
  EXPECT_TRUE(csr->get_vals() == nullptr);
  auto *csc = csr->Convert<format::CSC>(cpu_and_gpu());
  EXPECT_TRUE(same(csc->get_col_ptr(), CSC_COL_PTR));
  EXPECT_TRUE(same(csc->get_row(), CSC_ROW));
  delete csc;
  delete csr;
}

TEST_DEVICE(PrivateConverterAndClearedEdges) {
  // converter_order_two_tests.cc:262-351 + examples/custom_converter: a user-owned converter,
  // and multi-step conversion after the direct function was removed
  using I = int;
  int row[] = COO_ROW, col[] = COO_COL, vals[] = COO_VALS;
  format::COO<I, I, I> coo(kN, kM, kNnz, row, col, vals);
  converter::ConverterOrderTwo<I, I, I> mine;
  auto *dcoo = mine.Convert<format::CUDACOO<I, I, I>>(&coo, &gpu());
  // direct CUDACOO -> CUDACSC exists ...
  EXPECT_TRUE(mine.CanConvert(dcoo->get_id(), dcoo->get_context(),
                              format::CUDACSC<I, I, I>::get_id_static(), &gpu()));
  auto chain = mine.GetConversionChain(dcoo->get_id(), dcoo->get_context(),
                                       format::CUDACSC<I, I, I>::get_id_static(), {&gpu()});
  EXPECT_EQ(std::get<1>(*chain), 1u);
  // ... remove it: the chain goes through CUDACSR
  mine.ClearConversionFunctions(format::CUDACOO<I, I, I>::get_id_static(),
                                format::CUDACSC<I, I, I>::get_id_static());
  chain = mine.GetConversionChain(dcoo->get_id(), dcoo->get_context(),
                                  format::CUDACSC<I, I, I>::get_id_static(), {&gpu()});
  EXPECT_TRUE(chain.has_value());
  EXPECT_EQ(std::get<1>(*chain), 2u);
  auto *dcsc = mine.Convert<format::CUDACSC<I, I, I>>(dcoo, &gpu());
  auto *csc = dcsc->Convert<format::CSC>(&cpu_context);
  EXPECT_TRUE(same(csc->get_col_ptr(), CSC_COL_PTR));
  EXPECT_TRUE(same(csc->get_row(), CSC_ROW));
  EXPECT_TRUE(same(csc->get_vals(), CSC_VALS));
  // a user-registered function takes part in dispatch
  int calls = 0;
  mine.ClearConversionFunctions();
  mine.RegisterConversionFunction(
      format::COO<I, I, I>::get_id_static(), format::CUDACOO<I, I, I>::get_id_static(),
      [&calls](format::Format *src, context::Context *to) {
        calls++;
        return converter::fn::CooCUDACoo<I, I, I>(src, to);
      },
      converter::CPUToCUDA);
  delete mine.Convert<format::CUDACOO<I, I, I>>(&coo, &gpu());
  EXPECT_EQ(calls, 1);
  mine.Reset();
  EXPECT_TRUE(mine.CanConvert(coo.get_id(), coo.get_context(),
                              format::CSC<I, I, I>::get_id_static(), cpu_and_gpu()));
  delete csc;
  delete dcsc;
  delete dcoo;
}

// ================================================================== constructors
TEST_DEVICE(CsrConstructorSorts) {
  // format/csr_tests.cc:80-115: row 0 holds columns {2, 0}; the constructor sorts in place
  int row_ptr[] = {0, 2, 3, 3, 4}, col[] = {2, 0, 1, 3}, vals[] = {5, 4, 7, 9};
  format::CSR<int, int, int> csr(4, 4, row_ptr, col, vals, format::kNotOwned);
  EXPECT_TRUE(same(col, {0, 2, 1, 3}));
  EXPECT_TRUE(same(vals, {4, 5, 7, 9}));
  int col2[] = {2, 0, 1, 3}, vals2[] = {5, 4, 7, 9};
  format::CSR<int, int, int> keep(4, 4, row_ptr, col2, vals2, format::kNotOwned, true);
  EXPECT_TRUE(same(col2, {2, 0, 1, 3}));
  int col3[] = {2, 0, 1, 3};
  format::CSR<int, int, void> novals(4, 4, row_ptr, col3, nullptr, format::kNotOwned);
  EXPECT_TRUE(same(col3, {0, 2, 1, 3}));
}

TEST_DEVICE(CooConstructorSorts) {
  // format/coo_tests.cc:77-115
  int row[] = {0, 0, 3, 1}, col[] = {2, 0, 3, 1}, vals[] = {5, 4, 9, 7};
  format::COO<int, int, int> coo(4, 4, 4, row, col, vals, format::kNotOwned);
  EXPECT_TRUE(same(row, {0, 0, 1, 3}));
  EXPECT_TRUE(same(col, {0, 2, 1, 3}));
  EXPECT_TRUE(same(vals, {4, 5, 7, 9}));
  int row2[] = {0, 0, 3, 1}, col2[] = {2, 0, 3, 1};
  format::COO<int, int, void> novals(4, 4, 4, row2, col2, nullptr, format::kNotOwned);
  EXPECT_TRUE(same(row2, {0, 0, 1, 3}));
  EXPECT_TRUE(same(col2, {0, 2, 1, 3}));
}

// ================================================================== transfers
TEST_DEVICE(CsrCudaCsrRoundTrip) {
  // converter_order_two_cuda_tests.cu:11-49
  int row_ptr[] = CSR_ROW_PTR, col[] = COO_COL;
  float vals[] = COO_VALS;
  format::CSR<int, int, float> csr(kN, kM, row_ptr, col, vals, format::kNotOwned);
  auto *d = csr.Convert<format::CUDACSR>(&gpu());
  EXPECT_TRUE(d->get_context()->IsEquivalent(&gpu()));
  EXPECT_EQ(d->get_dimensions()[1], (format::DimensionType)kM);  // m is kept
  auto *clone = d->Clone()->AsAbsolute<format::CUDACSR<int, int, float>>();
  EXPECT_TRUE(clone->get_col() != d->get_col());
  auto *back = clone->Convert<format::CSR>(&cpu_context);
  EXPECT_TRUE(std::memcmp(back->get_row_ptr(), row_ptr, sizeof(row_ptr)) == 0);
  EXPECT_TRUE(std::memcmp(back->get_col(), col, sizeof(col)) == 0);
  EXPECT_TRUE(std::memcmp(back->get_vals(), vals, sizeof(vals)) == 0);
  // same device: the peer-copy edge is usable and yields an independent copy
  auto *peer = d->get_converter()->Convert(d, d->get_id(), &gpu());
  EXPECT_TRUE(peer == d);  // same type + equivalent context: returned as is
  delete back;
  delete clone;
  delete d;
}

TEST_DEVICE(ArrayCudaArrayRoundTrip) {
  // converter_order_one_cuda_tests.cu:15-106
  float v[] = {0.0f, 0.1f, 0.2f};
  format::Array<float> arr(3, v, format::kNotOwned);
  auto *d = arr.Convert<format::CUDAArray>(&gpu());
  EXPECT_EQ(d->get_num_nnz(), (format::DimensionType)3);
  auto *back = d->Convert<format::Array>(&cpu_context);
  EXPECT_TRUE(std::memcmp(back->get_vals(), v, sizeof(v)) == 0);
  delete back;
  delete d;
}

// ================================================================== functionality_common.inc
static const int n3 = 3, nnz3 = 4;
#define ROW_PTR3 {0, 2, 3, 4}
#define COLS3 {1, 2, 0, 0}
#define VALS3 {1, 2, 3, 4}

TEST_DEVICE(PermuteOrderTwoGoldens) {
  // permute_order_two_tests.cc:27-91, reorder_base_tests.cc:306-430
  int row_ptr[] = ROW_PTR3, cols[] = COLS3, vals[] = VALS3;
  format::CSR<int, int, int> csr(n3, n3, row_ptr, cols, vals, format::kNotOwned);
  int r_order[] = {1, 2, 0}, c_order[] = {2, 0, 1};
  {  // rows only
    auto *out = bases::ReorderBase::Permute2DRowWise<format::CSR>(r_order, &csr, {&gpu()}, true, true);
    EXPECT_TRUE(same(out->get_row_ptr(), {0, 1, 3, 4}));
    EXPECT_TRUE(same(out->get_col(), {0, 1, 2, 0}));
    EXPECT_TRUE(same(out->get_vals(), {4, 1, 2, 3}));
    delete out;
  }
  {  // columns only
    auto *out = bases::ReorderBase::Permute2DColWise<format::CSR>(c_order, &csr, {&gpu()}, true, true);
    EXPECT_TRUE(same(out->get_row_ptr(), {0, 2, 3, 4}));
    EXPECT_TRUE(same(out->get_col(), {0, 1, 2, 2}));
    EXPECT_TRUE(same(out->get_vals(), {1, 2, 3, 4}));
    delete out;
  }
  {  // rows and columns
    auto *out = bases::ReorderBase::Permute2DRowColumnWise<format::CSR>(r_order, c_order, &csr,
                                                                         {&gpu()}, true, true);
    EXPECT_TRUE(same(out->get_row_ptr(), {0, 1, 3, 4}));
    EXPECT_TRUE(same(out->get_col(), {2, 0, 1, 2}));
    EXPECT_TRUE(same(out->get_vals(), {4, 1, 2, 3}));
    delete out;
  }
  {  // the class itself, result left on the device, then applying the inverse restores A
    permute::PermuteOrderTwo<int, int, int> perm(r_order, c_order);
    auto *dev_out = perm.GetPermutation(&csr, {&gpu()}, true);
    EXPECT_TRUE(dev_out->Is<format::CUDACSR>());
    int *r_inv = bases::ReorderBase::InversePermutation(r_order, n3);
    int *c_inv = bases::ReorderBase::InversePermutation(c_order, n3);
    permute::PermuteOrderTwo<int, int, int> undo(r_inv, c_inv);
    auto *restored_dev = undo.GetPermutation(dev_out, {&gpu()}, false);
    auto *restored = restored_dev->Convert<format::CSR>(&cpu_context);
    EXPECT_TRUE(same(restored->get_row_ptr(), ROW_PTR3));
    EXPECT_TRUE(same(restored->get_col(), COLS3));
    EXPECT_TRUE(same(restored->get_vals(), VALS3));
    delete restored;
    delete restored_dev;
    delete dev_out;
    delete[] r_inv;
    delete[] c_inv;
  }
  // convert_input = false with a host CSR: no direct execution
  permute::PermuteOrderTwo<int, int, int> perm(r_order, r_order);
  EXPECT_THROW(perm.GetPermutation(&csr, {&gpu()}, false),
               utils::DirectExecutionNotAvailableException<std::vector<std::type_index>>);
  // cast (not convert) of a device result to a host class fails like the reference's As<>
  EXPECT_THROW((bases::ReorderBase::Permute2D<format::CSR>(r_order, &csr, {&gpu()}, true, false)),
               utils::TypeException);
}

TEST_DEVICE(Permute1DAndInversePermutation) {
  // permute_order_one_tests.cc:25-51, reorder_base_tests.cc:151-197, 299-305
  float original[] = {0.0f, 0.1f, 0.2f};
  int inverse_perm[] = {2, 0, 1};
  format::Array<float> arr(3, original, format::kNotOwned);
  auto *out = bases::ReorderBase::Permute1D<format::Array>(inverse_perm, &arr, {&gpu()}, true, true);
  EXPECT_TRUE(same(out->get_vals(), {0.1f, 0.2f, 0.0f}));
  delete out;
  int *perm = bases::ReorderBase::InversePermutation(inverse_perm, 3);
  EXPECT_TRUE(same(perm, {1, 2, 0}));
  delete[] perm;
}

template <typename I, typename N>
static void check_degree_ordering(const I *order, I n, const N *row_ptr, bool ascending) {
  // functionality_common.inc:67-90
  std::vector<I> perm(n);
  std::set<I> seen;
  for (I i = 0; i < n; i++) {
    EXPECT_TRUE(order[i] >= 0 && order[i] < n);
    seen.insert(order[i]);
    perm[order[i]] = i;
  }
  EXPECT_EQ(seen.size(), (size_t)n);
  for (I k = 0; k + 1 < n; k++) {
    const N d0 = row_ptr[perm[k] + 1] - row_ptr[perm[k]];
    const N d1 = row_ptr[perm[k + 1] + 1] - row_ptr[perm[k + 1]];
    EXPECT_TRUE(ascending ? d0 <= d1 : d0 >= d1);
  }
}

TEST_DEVICE(DegreeReorderLikeTheReferenceTests) {
  // degree_reorder_tests.cc:28-81
  int row_ptr[] = ROW_PTR3, cols[] = COLS3, vals[] = VALS3;
  format::CSR<int, int, int> csr(n3, n3, row_ptr, cols, vals, format::kNotOwned);
  reorder::DegreeReorder<int, int, int> asc(true);
  int *order = asc.GetReorder(&csr, {&gpu()}, true);
  check_degree_ordering(order, n3, row_ptr, true);
  EXPECT_TRUE(same(order, {2, 1, 0}));  // ties: descending id
  delete[] order;
  // parameters passed at call time override the constructor's
  reorder::DegreeReorderParams desc_params(false);
  order = asc.GetReorder(&csr, &desc_params, {&gpu()}, true);
  check_degree_ordering(order, n3, row_ptr, false);
  EXPECT_TRUE(same(order, {0, 1, 2}));
  delete[] order;
  // facade + COO input converted on the way (COO -> CUDACOO -> CUDACSR)
  int rows[] = {0, 0, 1, 2};
  format::COO<int, int, int> coo(n3, n3, nnz3, rows, cols, vals, format::kNotOwned);
  order = bases::ReorderBase::Reorder<reorder::DegreeReorder>({true}, &coo, {&gpu()}, true);
  check_degree_ordering(order, n3, row_ptr, true);
  delete[] order;
  EXPECT_THROW(asc.GetReorder(&coo, {&gpu()}, false),
               utils::DirectExecutionNotAvailableException<std::vector<std::type_index>>);
  // cached variant hands back the converted input
  auto cached = asc.GetReorderCached(&csr, {&gpu()}, true);
  EXPECT_EQ(std::get<0>(cached).size(), (size_t)1);
  EXPECT_EQ(std::get<0>(cached)[0].size(), (size_t)1);
  EXPECT_TRUE(std::get<0>(cached)[0][0]->get_id() ==
              (format::CUDACSR<int, int, int>::get_id_static()));
  delete std::get<0>(cached)[0][0];
  delete[] std::get<1>(cached);
}

// 5-point stencil on a w x h grid, as CSR (rows sorted)
static void grid_graph(int w, int h, std::vector<int> &row_ptr, std::vector<int> &col,
                       std::vector<float> &vals) {
  row_ptr.assign(1, 0);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int v = y * w + x;
      if (y > 0) col.push_back(v - w), vals.push_back(-1.f);
      if (x > 0) col.push_back(v - 1), vals.push_back(-1.f);
      col.push_back(v), vals.push_back(4.f);
      if (x + 1 < w) col.push_back(v + 1), vals.push_back(-1.f);
      if (y + 1 < h) col.push_back(v + w), vals.push_back(-1.f);
      row_ptr.push_back((int)col.size());
    }
}

TEST_DEVICE(RcmReorderPipelineOnAGrid) {
  // rcm_reorder_tests.cc:21-25 checks "is a permutation"; here additionally: Permute2D with the
  // RCM order renumbers the matrix consistently (confirm_renumbered_csr,
  // functionality_common.inc:138-157) and the bandwidth is the grid's short side + 1.
  const int w = 37, h = 23, n = w * h;
  std::vector<int> row_ptr, col;
  std::vector<float> vals;
  grid_graph(w, h, row_ptr, col, vals);
  format::CSR<int, int, float> csr(n, n, row_ptr.data(), col.data(), vals.data(), format::kNotOwned);
  int *order = bases::ReorderBase::Reorder<reorder::RCMReorder>({}, &csr, {&gpu()}, true);
  std::set<int> seen(order, order + n);
  EXPECT_EQ(seen.size(), (size_t)n);
  EXPECT_TRUE(*seen.begin() == 0 && *seen.rbegin() == n - 1);
  auto *out = bases::ReorderBase::Permute2D<format::CSR>(order, &csr, {&gpu()}, true, true);
  EXPECT_EQ(out->get_num_nnz(), csr.get_num_nnz());
  int bandwidth = 0;
  bool consistent = true;
  for (int u = 0; u < n && consistent; u++) {
    const int nu = order[u];
    const int len = row_ptr[u + 1] - row_ptr[u];
    consistent = out->get_row_ptr()[nu + 1] - out->get_row_ptr()[nu] == len;
    std::map<int, float> want;
    for (int k = row_ptr[u]; k < row_ptr[u + 1]; k++) want[order[col[k]]] = vals[k];
    int k = out->get_row_ptr()[nu];
    for (const auto &cv : want) {  // std::map iterates in ascending column order
      consistent = consistent && out->get_col()[k] == cv.first && out->get_vals()[k] == cv.second;
      bandwidth = std::max(bandwidth, std::abs(cv.first - nu));
      k++;
    }
  }
  EXPECT_TRUE(consistent);
  EXPECT_TRUE(bandwidth <= std::min(w, h) + 1);
  // ... and transposing the permuted matrix works on the device-resident result too
  auto *dcsr = out->Convert<format::CUDACSR>(&gpu());
  auto *dcsc = dcsr->Convert<format::CUDACSC>(&gpu());
  auto *csc = dcsc->Convert<format::CSC>(&cpu_context);
  // structurally symmetric matrix: CSC arrays equal the CSR arrays
  EXPECT_TRUE(std::memcmp(csc->get_col_ptr(), out->get_row_ptr(), sizeof(int) * (n + 1)) == 0);
  EXPECT_TRUE(std::memcmp(csc->get_row(), out->get_col(), sizeof(int) * col.size()) == 0);
  delete csc;
  delete dcsc;
  delete dcsr;
  delete out;
  delete[] order;
}

TEST_DEVICE(DegreeDistributionAndDegrees) {
  // degree_distribution_tests.cc:38-119, degrees_tests.cc
  int row_ptr[] = ROW_PTR3, cols[] = COLS3, vals[] = VALS3;
  format::CSR<int, int, int> csr(n3, n3, row_ptr, cols, vals, format::kNotOwned);
  feature::DegreeDistribution<int, int, int, float> f32;
  float *dist = f32.GetDistribution(&csr, {&gpu()}, true);
  EXPECT_TRUE(same(dist, {2.0f / 4.0f, 1.0f / 4.0f, 1.0f / 4.0f}));
  delete[] dist;
  feature::DegreeDistribution<int, int, int, double> f64;
  double *dist64 = f64.GetDistribution(&csr, {&gpu()}, true);
  EXPECT_TRUE(same(dist64, {2.0 / 4.0, 1.0 / 4.0, 1.0 / 4.0}));
  delete[] dist64;
  feature::Degrees<int, int, int> deg;
  int *d = deg.GetDegrees(&csr, {&gpu()}, true);
  EXPECT_TRUE(same(d, {2, 1, 1}));
  delete[] d;
  EXPECT_THROW(f32.GetDistribution(&csr, {&gpu()}, false),
               utils::DirectExecutionNotAvailableException<std::vector<std::type_index>>);
  auto *dcsr = csr.Convert<format::CUDACSR>(&gpu());
  dist = f32.GetDistribution(dcsr, {&gpu()}, false);  // direct execution on the device format
  EXPECT_TRUE(same(dist, {0.5f, 0.25f, 0.25f}));
  delete[] dist;
  delete dcsr;
}

int main(int argc, char **argv) {
  const bool no_device = argc > 1 && std::string(argv[1]) == "--no-device";
  int ran = 0;
  for (const TestCase &t : registry()) {
    if (no_device && t.needs_device) continue;
    const int before = g_failures;
    std::printf("[ RUN  ] %s\n", t.name);
    try {
      t.body();
    } catch (const std::exception &e) {
      std::printf("    FAILED: unexpected exception: %s\n", e.what());
      g_failures++;
    }
    std::printf("[ %s ] %s\n", g_failures == before ? " OK " : "FAIL", t.name);
    ran++;
  }
  std::printf("%d test(s) run, %d failure(s)\n", ran, g_failures);
  return g_failures == 0 ? 0 : 1;
}
