// TEST / MEASUREMENT INFRASTRUCTURE -- see oracle/README.md.
//
// plugin_bench: what a SparseBase user gets end to end.  The UNMODIFIED reference (header-only,
// USE_CUDA, compiled from $(REF)/src) with sb200_sparsebase_plugin.h registered; a host
// format::CSR goes in, host arrays come out, every conversion and dispatch decision is the
// reference's own:
//     inv = RCMReorder::GetReorder(csr, {&gpu}, /*convert_input=*/true)
//     out = PermuteOrderTwo(inv, inv).GetPermutation(csr, {&gpu}, true) -> Convert<CSR>(&cpu)
// on the 2-D Poisson grid of BASELINE.json configs[1].  Prints one JSON line; bench.py reports
// it as rcm.e2e_plugin.  `--check` also runs the reference's CPU functions and memcmp's.
//
//   plugin_bench [grid = 4096] [reps = 3] [--check]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "sparsebase/bases/reorder_base.h"
#include "sparsebase/context/cpu_context.h"
#include "sparsebase/format/csr.h"
#include "sparsebase/utils/logger.h"
#include "../sparsebase_b200/host/plugin/sb200_sparsebase_plugin.h"

using namespace sparsebase;
using I = int;
using N = int;
using V = float;

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  int grid = 4096, reps = 3;
  bool check = false;
  int pos = 0;
  for (int a = 1; a < argc; a++) {
    if (!std::strcmp(argv[a], "--check"))
      check = true;
    else if (pos++ == 0)
      grid = std::atoi(argv[a]);
    else
      reps = std::atoi(argv[a]);
  }
  utils::Logger::set_level(utils::LOG_LVL_NONE);
  const long n = (long)grid * grid;
  std::vector<N> rp(n + 1);
  std::vector<I> col;
  std::vector<V> val;
  col.reserve(5 * n);
  val.reserve(5 * n);
  for (long v = 0; v < n; v++) {
    const long x = v % grid, y = v / grid;
    rp[v] = (N)col.size();
    if (y > 0) col.push_back((I)(v - grid)), val.push_back(-1.f);
    if (x > 0) col.push_back((I)(v - 1)), val.push_back(-1.f);
    col.push_back((I)v), val.push_back(4.f);
    if (x + 1 < grid) col.push_back((I)(v + 1)), val.push_back(-1.f);
    if (y + 1 < grid) col.push_back((I)(v + grid)), val.push_back(-1.f);
  }
  rp[n] = (N)col.size();
  const size_t nnz = col.size();

  context::CPUContext cpu;
  context::CUDAContext gpu(0);
  auto conv = converter::ConverterStore::GetStore()
                  .get_converter<converter::ConverterOrderTwo<I, N, V>>();
  sb200_plugin::RegisterConversions<I, N, V>(*conv);
  sb200_plugin::RegisterTransfers<I, N, V>(*conv);
  format::CSR<I, N, V> csr((I)n, (I)n, rp.data(), col.data(), val.data(), format::kNotOwned, true);

  // the call sequence of examples/rcm_order + ReorderBase::Permute2D, CUDA context allowed
  auto each_call = [&](I *&inv_out, format::CSR<I, N, V> *&out) {
    reorder::RCMReorder<I, N, V> rcm;
    sb200_plugin::Register(rcm);
    // a host CSR with a CUDA context: the matcher keeps the CPU function for an identical key
    // (function_matcher_mixin.h:366-369), so the device format is what the user passes
    auto *d = csr.Convert<format::CUDACSR>(&gpu);
    inv_out = rcm.GetReorder(d, {&gpu}, false);
    permute::PermuteOrderTwo<I, N, V> perm(inv_out, inv_out);
    sb200_plugin::Register(perm);
    auto *p = perm.GetPermutation(d, {&gpu}, false);
    out = p->Convert<format::CSR>(&cpu);
    delete p;
    delete d;
  };

  I *inv = nullptr;
  format::CSR<I, N, V> *out = nullptr;
  each_call(inv, out);  // warm-up (context, pinned buffers, pools)
  double best = 1e30, sum = 0;
  for (int r = 0; r < reps; r++) {
    delete[] inv;
    delete out;
    const double t0 = now();
    each_call(inv, out);
    const double t = now() - t0;
    best = t < best ? t : best;
    sum += t;
  }
  const size_t h2d = (size_t)(n + 1) * sizeof(N) + nnz * (sizeof(I) + sizeof(V)) + 2 * n * sizeof(I);
  const size_t d2h = (size_t)(n + 1) * sizeof(N) + nnz * (sizeof(I) + sizeof(V)) + n * sizeof(I);
  int parity = -1;
  if (check) {
    reorder::RCMReorder<I, N, V> rcm;
    I *ref_inv = rcm.GetReorder(&csr, {&cpu}, false);
    permute::PermuteOrderTwo<I, N, V> perm(ref_inv, ref_inv);
    auto *ref = perm.GetPermutation(&csr, {&cpu}, false)->As<format::CSR>();
    parity = !std::memcmp(ref_inv, inv, n * sizeof(I)) &&
             !std::memcmp(ref->get_row_ptr(), out->get_row_ptr(), (n + 1) * sizeof(N)) &&
             !std::memcmp(ref->get_col(), out->get_col(), nnz * sizeof(I)) &&
             !std::memcmp(ref->get_vals(), out->get_vals(), nnz * sizeof(V));
  }
  std::printf("{\"workload\": \"RCMReorder + Permute2D on Poisson %dx%d through the reference's "
              "plugin API: host format::CSR in, host CSR + permutation out\", \"n\": %ld, "
              "\"nnz\": %zu, \"reps\": %d, \"ms_mean\": %.3f, \"ms_best\": %.3f, "
              "\"gnnz_per_s\": %.4f, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"parity\": %s}\n",
              grid, grid, n, nnz, reps, sum / reps * 1e3, best * 1e3, nnz / (sum / reps) / 1e9, h2d,
              d2h, parity < 0 ? "null" : (parity ? "true" : "false"));
  return parity == 0 ? 1 : 0;
}
