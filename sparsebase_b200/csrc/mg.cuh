// mg.cuh -- peer-memory communicator of the multi-GPU (row-block sharded) path.
//
// One rank per GPU.  Every rank owns a WINDOW in its HBM that all peers can address (CUDA IPC
// between processes, direct peer access inside one process; NVLink 5 / NVSwitch underneath).
// An exchange is a kernel that stores straight into the destination rank's window, followed by
// a flag barrier -- no host round trip, no staging copy on the sender, counts and offsets stay on
// the device.
//
//   window = [ control area | data area ]
//   control: flag[r]      one 64-bit epoch per peer (the barrier)
//            table[r][k]  kMgSlots 64-bit words per peer (small all-gathers: counts, extents)
//   data:    carved by every operator with MgLayout (all ranks carve identically)
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int kMgMaxRanks = 16;
constexpr int kMgSlots = 64;                       // 64-bit words per rank in the table
constexpr size_t kMgControlBytes = 64 * 1024;      // flags + tables, well inside

struct MgControl {
  unsigned long long flag[kMgMaxRanks];
  unsigned long long pad[16];
  unsigned long long table[kMgMaxRanks][kMgSlots];
};
static_assert(sizeof(MgControl) <= kMgControlBytes, "control area too small");

// What the kernels see (passed by value).
struct MgPeers {
  int rank, world;
  char *win[kMgMaxRanks];  // win[r] = rank r's window in THIS rank's address space
  __host__ __device__ MgControl *ctl(int r) const { return reinterpret_cast<MgControl *>(win[r]); }
  __host__ __device__ char *data(int r) const { return win[r] + kMgControlBytes; }
};

}  // namespace sb200

struct sb200_mg_comm {
  sb200::MgPeers peers;
  int device;
  size_t window_bytes;      // whole window
  unsigned long long epoch; // barriers issued so far (same on every rank)
  bool ipc;                 // peers opened through CUDA IPC (to be closed)
  bool owns_window;
  void *aux_stream = nullptr;  // cudaStream_t, created on first use (overlapped exchanges)
  void *aux_event = nullptr;   // cudaEvent_t
};

namespace sb200 {

// Carves the data area; every rank runs the same sequence and gets the same offsets.
class MgLayout {
 public:
  explicit MgLayout(const sb200_mg_comm *c) : c_(c), off_(0) {}
  // offset (bytes, inside the data area) of a region of `bytes`
  size_t take(size_t bytes) {
    const size_t at = off_;
    off_ += (bytes + 255) & ~size_t(255);
    SB_REQUIRE(kMgControlBytes + off_ <= c_->window_bytes, SB200_ERR_BAD_ARG,
               "multi-GPU window too small: need %zu bytes, have %zu (create the communicator "
               "with a larger window)",
               kMgControlBytes + off_, c_->window_bytes);
    return at;
  }
  size_t used() const { return off_; }

 private:
  const sb200_mg_comm *c_;
  size_t off_;
};

void mg_barrier(sb200_mg_comm *c, cudaStream_t st);

}  // namespace sb200
