// device.h -- device-side state of one engine: the published CSR snapshot in HBM,
// scratch buffers, and the launch sequences of the hot path.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/zgpu.h"
#include "schema.h"
#include "store.h"

namespace zg {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grow-only; returns false on allocation failure
  bool ensure(size_t bytes);
  void release();
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

// Immutable once published; shared_ptr keeps it alive for in-flight readers.
struct Snapshot {
  DevBuf row_ptr, col, exp, prog, rrow_ptr, rcol, type_bit_base;
  // incremental publish (build.cu gpu_apply_delta): the edge arrays are re-emitted into their alternates and
  // swapped; per-class relationship counts tell when a class becomes (non-)empty and the program changes
  DevBuf col_alt, exp_alt, rcol_alt;
  std::vector<uint64_t> cls_count;
  std::vector<uint64_t> type_base;   // per type: first row_ptr index of its objects
  std::vector<uint32_t> type_ncls;   // per type: row stride
  bool delta_ok = false;             // built by the GPU builder: counts are known
  bool resources_stale = false;      // resources[] predate an incremental publish: rebuilt on first use
  std::vector<uint32_t> n_objects;  // per type
  uint64_t total_bits = 0;          // visited bitmap size for the reverse BFS
  // Per slot: device array of FlatLookupClass when the slot is a flat union of direct
  // (non-expiring) relations; flat_n[slot] = -1 otherwise.
  std::vector<DevBuf> flat_cls;
  std::vector<int> flat_n;
  std::vector<DevBuf> resources;  // per type
  std::vector<uint64_t> n_resources;
  uint32_t prog_bytes = 0;
  uint32_t max_leaves = 1;
  bool has_nonpure = false;
  uint64_t n_tuples = 0, bytes = 0, revision = 0;
  ~Snapshot();
};

class Device {
 public:
  ~Device();
  std::string init(int device, uint64_t subq_cap, uint32_t budget);
  std::string publish(const HostSnapshot& h, const Schema& sc, uint64_t revision);  // host-built arrays
  // Builds the CSR on the GPU (build.cu); verify: also build on the host and compare every array.
  std::string publish_gpu(const Store& store, const Schema& sc, uint64_t revision, bool verify);
  // Merges the store's journal into the resident snapshot. Returns "" on success, "full" when only a rebuild
  // can publish this state (no snapshot, re-layout, bulk load, huge delta), else an error message.
  std::string publish_delta(const Store& store, const Schema& sc, uint64_t revision);
  uint64_t delta_publishes = 0, full_publishes = 0;
  // ZGPU_VERIFY_BUILD=1: every array of the resident snapshot against the host builder's
  std::string verify_against_host(const Store& store, const Schema& sc);
  std::string verify_snapshot(Snapshot& s, const Store& store, const Schema& sc, const HostSnapshot* lay);
  double last_build_ms = 0;

  // d_items / d_out are device pointers. count_bytes != nullptr selects the
  // instrumented kernel variant. Returns ZG_* and fills err.
  int check_device(const zg_check* d_items, uint64_t n, uint8_t* d_out, cudaStream_t st, bool raw_items,
                   uint64_t* count_bytes, std::string* err, bool top = true, const unsigned long long* ready = nullptr);
  int check_host(const zg_check* items, uint64_t n, uint8_t* out, std::string* err);
  // Several callers' requests answered by ONE launch sequence (see Batcher in capi.cu).
  struct HostReq {
    const zg_check* items;
    uint64_t n;
    uint8_t* out;
  };
  int check_host_multi(const std::vector<HostReq>& reqs, std::string* err);
  int lookup(uint16_t res_type, const zg_check& proto, std::vector<uint32_t>* ids, std::string* err);
  // K <= 64 LookupResources in one launch sequence (multi-source reverse walk, one verification launch, one
  // sort, one copy back). rcs[i] = ZG_OK / ZG_EDEPTH per lookup; the call's own code covers the batch.
  struct LookupReq {
    uint16_t res_type;
    zg_check proto;
  };
  int lookup_batch(const std::vector<LookupReq>& reqs, std::vector<std::vector<uint32_t>>* ids, std::vector<int>* rcs,
                   std::string* err);
  uint64_t lookup_batches = 0, lookups_batched = 0;

  // ---- object-hash sharded store: one pass at a time, sub-queries routed by the host ----
  uint32_t shard_count = 1, shard_rank = 0;
  // Evaluates `n` host queries as pass `level` (level 0: caller items; deeper: routed
  // sub-queries, depth in flags). *n_sub = sub-queries this pass raised.
  int shard_pass(const zg_check* queries, uint64_t n, int level, uint64_t* n_sub, std::string* err);
  int shard_subqueries(int level, zg_check* out, uint64_t n, std::string* err);
  // child_vals: one byte per raised sub-query (kValT|kValE bits) in emission order. out: one byte
  // per query of the level: v1 codes at level 0, value bits deeper.
  int shard_fold(int level, const uint8_t* child_vals, uint64_t n_sub, uint8_t* out, std::string* err);
  // ---- the same protocol with everything resident on the device (dist.DeviceShardedChecker) ----
  // Bucketises n check items by owner = res % n_dest into d_routed (destination-major) with their source
  // indices in d_src; counts[d] (host) = items for destination d. d_items == nullptr: the sub-queries raised by
  // pass `level` (shard_pass / shard_pass_dev).
  int route_by_owner(const zg_check* d_items, uint64_t n, int level, uint32_t n_dest, zg_check* d_routed, uint32_t* d_src,
                     uint64_t* counts, std::string* err);
  int shard_pass_dev(const zg_check* d_queries, uint64_t n, int level, uint64_t* n_sub, std::string* err);
  // d_child_vals: one byte per raised sub-query in ROUTED order, d_src = the source indices route_by_owner wrote.
  // d_out: one byte per query of the level (v1 codes when final_codes, else value bits).
  int shard_fold_dev(int level, const uint8_t* d_child_vals, const uint32_t* d_src, uint64_t n_sub, uint8_t* d_out,
                     bool final_codes, std::string* err);
  int unroute(const uint32_t* d_src, const uint8_t* d_val, uint64_t n, uint8_t* d_out, std::string* err);

  std::shared_ptr<Snapshot> snap;
  cudaStream_t stream = nullptr;
  int device = 0;
  uint64_t launches = 0, passes = 0, checks = 0;
  uint64_t coalesced_launches = 0, coalesced_requests = 0;
  uint64_t streamed_calls = 0;  // host calls whose items were streamed in behind the running kernel
  bool stream_h2d = true;       // ZGPU_NO_STREAM_H2D=1: one copy, then the kernel (A/B measurement)
  uint64_t split_batches = 0;  // batches answered in halves because their sub-queries overflowed the pass buffer
  // cumulative device-side event counters: stack spills to HBM, batches that switched the path memo on
  void read_events(uint64_t* spills, uint64_t* memo_batches);
  uint64_t lookups_rbfs = 0, lookups_exhaustive = 0, lookups_flat = 0;
  bool use_rbfs = true;  // ZGPU_NO_RBFS=1: LookupResources checks every resource of the type  // batcher: launches that served > 1 caller
  double last_ms = 0;
  bool invert = true;  // direction-optimised probes (ZG_FLAG_FORWARD_ONLY / ZGPU_NO_INVERT=1 disable)
  uint32_t now = 0;  // clock for expiration, set by the caller before each hot-path call
  uint64_t last_alg_bytes = 0;

 private:
  std::string finish_publish(std::shared_ptr<Snapshot> s, const HostSnapshot& lay, const Schema& sc, uint64_t revision);
  // One check_kernel launch over `nq` queries: val (may be null) receives the per-(query, leaf) value
  // bits, out (may be null) the per-query v1 codes / value bits.
  int run_pass(const Snapshot& s, const zg_check* queries, uint64_t nq, uint8_t* val, uint8_t* out, bool final_codes,
               bool raw, zg_check* subq, uint32_t* subq_parent, cudaStream_t st, bool count, std::string* err,
               const unsigned long long* ready = nullptr);
  int sm_count_ = 0, blocks_per_sm_ = 0, blocks_per_sm_count_ = 0;
  uint32_t spill_cap_ = 4096, budget_ = 1u << 20;
  uint64_t subq_cap_ = 1ull << 22;
  DevBuf spill_, ctrl_, memo_;
  uint32_t memo_entries_ = 8192, memo_after_ = 2048;  // ctrl: [0] next, [1] subq_count, [2] alg_bytes, then flags u32
  std::vector<DevBuf> q_, parent_, val_;  // per pass level: queries, raising (query, leaf), leaf values
  std::vector<uint64_t> shard_nq_, shard_nsub_;  // sharded mode: queries / raised sub-queries per level
  DevBuf shard_tmp_, route_ctrl_;
  DevBuf stage_in_, stage_out_, lk_jobs_, lk_codes_, lk_ids_;
  DevBuf rb_visited_, rb_front_[2], rb_cand_;
  DevBuf lb_params_, lb_owner_, lb_ctrl_, lb_keys_[2];  // batched lookups
  uint64_t lb_cap_ = 1ull << 24;  // frontier entries / candidates of one lookup batch (ZGPU_LOOKUP_BATCH_CAP)
  void* pin_lk_ = nullptr;
  size_t pin_lk_cap_ = 0;
  uint64_t rb_cap_ = 1ull << 22;  // frontier / candidate capacity before falling back to the exhaustive scan
  // Reverse-BFS candidates of type res_type for the subject in proto; *overflow -> use the exhaustive list.
  int lookup_candidates(const Snapshot& s, uint16_t res_type, const zg_check& proto, uint64_t* n_cand, bool* overflow,
                        std::string* err);
  // streamed admission (check_host_multi): copy stream, "reset done" event, pinned per-chunk counters
  cudaStream_t cstream_ = nullptr;
  cudaEvent_t ev_reset_ = nullptr;
  unsigned long long* pin_ready_ = nullptr;
  size_t pin_ready_cap_ = 0;
  static constexpr uint64_t kStreamChunkItems = 1ull << 17;  // 2 MB of items per chunk
  static constexpr uint64_t kStreamMinItems = 1ull << 18;    // smaller calls: one plain copy
  static constexpr uint64_t kParallelGatherItems = 1ull << 20;  // coalesced groups from here on are gathered by helpers
  static constexpr unsigned kGatherThreads = 8;
  void* pin_in_ = nullptr;
  void* pin_out_ = nullptr;
  size_t pin_in_cap_ = 0, pin_out_cap_ = 0;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, last_done_ = nullptr;
  bool have_last_ = false;
  bool timing_pending_ = false;

 public:
  void finish_timing();
};

// Build time and tuning macros of the kernels in this library.
const char* build_info();

// Device radix sort of n u32 keys (build.cu keeps the cub instantiations in one translation unit).
std::string sort_u32(const uint32_t* d_in, uint32_t* d_out, uint64_t n, cudaStream_t st);
std::string sort_u64(const unsigned long long* d_in, unsigned long long* d_out, uint64_t n, int end_bit, cudaStream_t st);
std::string gpu_build_snapshot(const Store& store, const Schema& sc, HostSnapshot* lay, cudaStream_t st, Snapshot* s);
// Applies store.journal to the arrays of `s` (same layout). cls_delta: per-class change of the relationship count.
// "relayout" / "corrupt": the caller must rebuild.
std::string gpu_apply_delta(const Store& store, const Schema& sc, const HostSnapshot& lay, cudaStream_t st, Snapshot* s,
                            std::vector<uint32_t>* cls_delta);
// (Re)builds s->resources[] / n_resources[] from the forward row table.
std::string gpu_resource_lists(Snapshot* s, cudaStream_t st);

}  // namespace zg
