// TEST INFRASTRUCTURE ONLY -- see oracle/README.md.
//
// plugin_demo: the UNMODIFIED reference (compiled header-only with USE_CUDA from the sources
// under $(REF)/src, nothing copied) with sparsebase_b200/host/plugin/sb200_sparsebase_plugin.h
// registered into it.  For a few synthetic matrices it runs the hot path twice inside one
// process -- through the reference's CPU functions ({&cpu_context}) and through the sb200 CUDA
// functions reached by the reference's own dispatch ({&gpu_context}) -- and memcmp's the
// results.  Built into oracle/_ref/plugin_demo by `make -C oracle ref` (development container,
// where /root/reference is mounted); the binary travels to the GPU box with the snapshot and
// tests/test_plugin_demo.py runs it there.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <random>
#include <vector>

#include "sparsebase/bases/reorder_base.h"
#include "sparsebase/context/cpu_context.h"
#include "sparsebase/format/coo.h"
#include "sparsebase/format/csc.h"
#include "sparsebase/format/csr.h"
#include "sparsebase/utils/logger.h"
// the plugin (includes the reference's CUDA format headers)
#include "../sparsebase_b200/host/plugin/sb200_sparsebase_plugin.h"

// degree_reorder.cc:41-45 reads and writes mr[n], one element past `new IDType[n]()`.  Like
// oracle/ref_harness.cc, give every array allocation of this test program zeroed slack so that
// the stray slot has the value 0 deterministically instead of corrupting the heap.
void *operator new[](std::size_t sz) {
  void *p = std::calloc(1, sz + 64);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }

using namespace sparsebase;
using I = int;
using N = int;
using V = float;

static int g_fail = 0;
#define CHECK(cond, what)                                        \
  do {                                                           \
    if (!(cond)) {                                               \
      std::printf("  MISMATCH: %s (%s:%d)\n", what, __FILE__, __LINE__); \
      g_fail++;                                                  \
    }                                                            \
  } while (0)

#define STEP(what) std::fprintf(stderr, "  .. %s\n", what)

struct Coo {
  I n;
  std::vector<I> row, col;
  std::vector<V> val;
};

static Coo grid(int w, int h) {
  Coo c;
  c.n = w * h;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int v = y * w + x;
      auto add = [&](int u, float a) { c.row.push_back(v), c.col.push_back(u), c.val.push_back(a); };
      if (y > 0) add(v - w, -1.f);
      if (x > 0) add(v - 1, -1.f);
      add(v, 4.f);
      if (x + 1 < w) add(v + 1, -1.f);
      if (y + 1 < h) add(v + w, -1.f);
    }
  return c;
}

// symmetric random graph, de-duplicated, delivered in random order (exercises the COO sort)
static Coo random_graph(int n, int deg, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<std::pair<I, I>> e;
  for (long k = 0; k < (long)n * deg / 2; k++) {
    I a = rng() % n, b = rng() % n;
    if (a == b) continue;
    e.emplace_back(a, b), e.emplace_back(b, a);
  }
  std::sort(e.begin(), e.end());
  e.erase(std::unique(e.begin(), e.end()), e.end());
  std::shuffle(e.begin(), e.end(), rng);
  Coo c;
  c.n = n;
  for (size_t k = 0; k < e.size(); k++) {
    c.row.push_back(e[k].first), c.col.push_back(e[k].second);
    c.val.push_back((float)(rng() & 0xffff) + 0.5f);
  }
  return c;
}

template <typename T>
static bool same(const T *a, const T *b, size_t cnt) {
  return std::memcmp(a, b, cnt * sizeof(T)) == 0;
}

static void run_case(const char *name, Coo c, context::CPUContext &cpu, context::CUDAContext &gpu) {
  std::printf("[ %s ] n=%d nnz=%zu\n", name, c.n, c.row.size());
  const size_t nnz = c.row.size();
  const I n = c.n;
  // two identical host COOs (the constructor sorts in place)
  Coo c2 = c;
  format::COO<I, N, V> coo_cpu(n, n, nnz, c.row.data(), c.col.data(), c.val.data());
  format::COO<I, N, V> coo_gpu(n, n, nnz, c2.row.data(), c2.col.data(), c2.val.data());

  // ---- COO -> CSR: reference CPU function vs the plugin's COO -> CUDACSR (+ reference D2H)
  STEP("COO->CSR");
  auto *csr_ref = coo_cpu.Convert<format::CSR>(&cpu);
  auto *dcsr = coo_gpu.Convert<format::CUDACSR>(&gpu);
  auto *csr_got = dcsr->Convert<format::CSR>(&cpu);
  CHECK(same(csr_ref->get_row_ptr(), csr_got->get_row_ptr(), n + 1), "COO->CSR row_ptr");
  CHECK(same(csr_ref->get_col(), csr_got->get_col(), nnz), "COO->CSR col");
  CHECK(same(csr_ref->get_vals(), csr_got->get_vals(), nnz), "COO->CSR vals");

  // ---- CSR -> CSC
  STEP("CSR->CSC");
  auto *csc_ref = csr_ref->Convert<format::CSC>(&cpu);
  auto *csc_got = dcsr->Convert<format::CSC>(&cpu);
  CHECK(same(csc_ref->get_col_ptr(), csc_got->get_col_ptr(), n + 1), "CSR->CSC col_ptr");
  CHECK(same(csc_ref->get_row(), csc_got->get_row(), nnz), "CSR->CSC row");
  CHECK(same(csc_ref->get_vals(), csc_got->get_vals(), nnz), "CSR->CSC vals");

  // ---- DegreeReorder / RCMReorder: the same operator object, CPU context vs CUDA context
  STEP("DegreeReorder");
  for (bool asc : {true, false}) {
    reorder::DegreeReorder<I, N, V> deg(asc);
    sb200_plugin::Register(deg);
    I *ref = deg.GetReorder(csr_ref, {&cpu}, false);
    I *got = deg.GetReorder(dcsr, {&gpu}, false);
    CHECK(same(ref, got, n), asc ? "DegreeReorder asc" : "DegreeReorder desc");
    delete[] ref;
    delete[] got;
  }
  STEP("RCMReorder");
  reorder::RCMReorder<I, N, V> rcm;
  sb200_plugin::Register(rcm);
  I *rcm_ref = rcm.GetReorder(csr_ref, {&cpu}, false);
  // host CSR in, CUDA context allowed: the reference's matcher sees an identical key and keeps
  // the CPU function (function_matcher_mixin.h:366-369), so the device format is passed
  I *rcm_got = rcm.GetReorder(dcsr, {&gpu}, false);
  CHECK(same(rcm_ref, rcm_got, n), "RCMReorder");

  // ---- Permute2D with the RCM order
  STEP("Permute2D");
  permute::PermuteOrderTwo<I, N, V> perm(rcm_ref, rcm_ref);
  sb200_plugin::Register(perm);
  auto *p_ref = perm.GetPermutation(csr_ref, {&cpu}, false)->As<format::CSR>();
  auto *p_dev = perm.GetPermutation(dcsr, {&gpu}, false);
  auto *p_got = p_dev->Convert<format::CSR>(&cpu);
  CHECK(same(p_ref->get_row_ptr(), p_got->get_row_ptr(), n + 1), "Permute2D row_ptr");
  CHECK(same(p_ref->get_col(), p_got->get_col(), nnz), "Permute2D col");
  CHECK(same(p_ref->get_vals(), p_got->get_vals(), nnz), "Permute2D vals");

  // ---- Permute1D
  STEP("Permute1D");
  format::Array<V> arr(nnz < (size_t)n ? nnz : n, csr_ref->get_vals());
  std::vector<I> order(arr.get_num_nnz());
  for (size_t i = 0; i < order.size(); i++) order[i] = (I)((i * 7919) % order.size());
  if (order.size() % 7919 != 0) {  // 7919 prime: a permutation unless it divides the length
    permute::PermuteOrderOne<I, V> p1(order.data());
    sb200_plugin::Register(p1);
    auto *a_ref = p1.GetPermutation(&arr, {&cpu}, false)->As<format::Array>();
    auto *d_arr = arr.Convert<format::CUDAArray>(&gpu);  // the reference's own H2D function
    auto *a_dev = p1.GetPermutation(d_arr, {&gpu}, false);
    auto *a_got = a_dev->Convert<format::Array>(&cpu);
    CHECK(same(a_ref->get_vals(), a_got->get_vals(), order.size()), "Permute1D");
    delete a_got;
    delete a_dev;
    delete d_arr;
    delete a_ref;
  }

  // ---- DegreeDistribution / Degrees
  STEP("DegreeDistribution/Degrees");
  feature::DegreeDistribution<I, N, V, float> dd;
  sb200_plugin::Register(dd);
  float *d_ref = dd.GetDistribution(csr_ref, {&cpu}, false);
  float *d_got = dd.GetDistribution(dcsr, {&gpu}, false);
  CHECK(same(d_ref, d_got, n), "DegreeDistribution<float>");
  feature::Degrees<I, N, V> dg;
  sb200_plugin::Register(dg);
  I *g_ref = dg.GetDegrees(csr_ref, {&cpu}, false);
  I *g_got = dg.GetDegrees(dcsr, {&gpu}, false);
  CHECK(same(g_ref, g_got, n), "Degrees");

  // ---- BOBAReorder (the reference works on the host COO; the plugin on the CUDACSR)
  STEP("BOBAReorder");
  {
    reorder::BOBAReorder<I, N, V> boba(true);
    sb200_plugin::Register(boba);
    auto *coo_h = csr_ref->template Convert<format::COO>(&cpu);
    I *b_ref = boba.GetReorder(coo_h, {&cpu}, false);
    I *b_got = boba.GetReorder(dcsr, {&gpu}, false);
    CHECK(same(b_ref, b_got, n), "BOBAReorder");
    delete[] b_ref;
    delete[] b_got;
    delete coo_h;
  }

  // ---- ReorderHeatmap (rows and columns renumbered by the RCM permutation)
  STEP("ReorderHeatmap");
  {
    reorder::ReorderHeatmap<I, N, V, float> hm{reorder::ReorderHeatmapParams(7)};
    sb200_plugin::Register(hm);
    format::Array<I> perm_h(n, rcm_ref, format::kNotOwned);
    auto *h_ref = hm.Get(csr_ref, &perm_h, &perm_h, {&cpu}, false)->As<format::Array>();
    auto *perm_d = perm_h.Convert<format::CUDAArray>(&gpu);
    auto *h_got = hm.Get(dcsr, perm_d, perm_d, {&gpu}, false)->As<format::Array>();
    CHECK(same(h_ref->get_vals(), h_got->get_vals(), 49), "ReorderHeatmap<float> 7x7");
    delete h_got;
    delete perm_d;
    delete h_ref;
  }

  STEP("cleanup");
  delete[] d_ref;
  delete[] d_got;
  delete[] g_ref;
  delete[] g_got;
  delete[] rcm_ref;
  delete[] rcm_got;
  delete p_got;
  delete p_dev;
  delete csc_ref;
  delete csc_got;
  delete csr_got;
  delete dcsr;
  delete csr_ref;
}

int main() {
  utils::Logger::set_level(utils::LOG_LVL_NONE);
  context::CPUContext cpu;
  context::CUDAContext gpu(0);
  // the shared converter of this type triple: keep it alive and add the sb200 edges
  auto conv = converter::ConverterStore::GetStore()
                  .get_converter<converter::ConverterOrderTwo<I, N, V>>();
  sb200_plugin::RegisterConversions<I, N, V>(*conv);
  // CSR <-> CUDACSR through the plugin's staged transfers instead of the reference's own
  sb200_plugin::RegisterTransfers<I, N, V>(*conv);

  run_case("grid 61x47", grid(61, 47), cpu, gpu);
  run_case("grid 300x200", grid(300, 200), cpu, gpu);
  run_case("random n=20000 deg 8", random_graph(20000, 8, 7), cpu, gpu);
  run_case("random n=3000 deg 40", random_graph(3000, 40, 9), cpu, gpu);
  std::printf("%s: %d mismatch(es)\n", g_fail ? "FAILED" : "ALL EQUAL", g_fail);
  return g_fail ? 1 : 0;
}
