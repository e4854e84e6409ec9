// device.cu -- snapshot upload and the launch sequences of the hot path.
//
// Call graph it replaces in the reference:
//   CheckBulkPermissions (pkg/authz/check.go:48, postfilter.go:134) -> check_*()
//   LookupResources      (pkg/authz/lookups.go:65)                    -> lookup()
#include "device.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "kernels.cuh"

namespace zg {

#define ZG_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(_e);               \
      return ZG_ECUDA;                                                                  \
    }                                                                                   \
  } while (0)

#define ZG_STR2(x) #x
#define ZG_STR(x) ZG_STR2(x)
const char* build_info() {
  return "libzgpu sm_100a built " __DATE__ " " __TIME__ " L2_MODE=" ZG_STR(ZG_L2_MODE) " PRESENCE=" ZG_STR(ZG_PRESENCE) " POP_FAST=" ZG_STR(ZG_POP_FAST) " L2_MATCH=" ZG_STR(ZG_L2_MATCH) " L2_FILTER=" ZG_STR(ZG_L2_FILTER) " L2_SPLIT=" ZG_STR(ZG_L2_SPLIT)
         " STACK_CAP=" ZG_STR(ZG_STACK_CAP) " RSET_CAP=" ZG_STR(ZG_RSET_CAP) " MIN_BLOCKS=" ZG_STR(ZG_MIN_BLOCKS);
}

bool DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return true;
  release();
  size_t want = std::max<size_t>(bytes, 256);
  if (cudaMalloc(&p, want) != cudaSuccess) {
    p = nullptr;
    cap = 0;
    cudaGetLastError();
    return false;
  }
  cap = want;
  return true;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}
Snapshot::~Snapshot() {
  row_ptr.release();
  col.release();
  exp.release();
  prog.release();
  rrow_ptr.release();
  rcol.release();
  col_alt.release();
  exp_alt.release();
  rcol_alt.release();
  type_bit_base.release();
  for (auto& r : resources) r.release();
  for (auto& f : flat_cls) f.release();
}

Device::~Device() {
  if (stream) cudaStreamSynchronize(stream);
  snap.reset();
  spill_.release();
  ctrl_.release();
  memo_.release();
  for (auto* v : {&q_, &parent_, &val_})
    for (auto& b : *v) b.release();
  stage_in_.release();
  stage_out_.release();
  shard_tmp_.release();
  route_ctrl_.release();
  lk_jobs_.release();
  lk_codes_.release();
  lk_ids_.release();
  rb_visited_.release();
  rb_front_[0].release();
  rb_front_[1].release();
  rb_cand_.release();
  for (auto* b : {&lb_params_, &lb_owner_, &lb_ctrl_, &lb_keys_[0], &lb_keys_[1]}) b->release();
  if (pin_in_) cudaFreeHost(pin_in_);
  if (pin_out_) cudaFreeHost(pin_out_);
  if (pin_ready_) cudaFreeHost(pin_ready_);
  if (pin_lk_) cudaFreeHost(pin_lk_);
  if (cstream_) {
    cudaStreamSynchronize(cstream_);
    cudaStreamDestroy(cstream_);
  }
  if (ev_reset_) cudaEventDestroy(ev_reset_);
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  if (last_done_) cudaEventDestroy(last_done_);
  if (stream) cudaStreamDestroy(stream);
}

static size_t smem_bytes(uint32_t prog_bytes) {
  return prog_bytes + static_cast<size_t>(kWarpsPerBlock) * kWarpSmem;
}

std::string Device::init(int dev, uint64_t subq_cap, uint32_t budget) {
  if (const char* v = std::getenv("ZGPU_NO_INVERT")) invert = !(*v && *v != '0');
  if (const char* v = std::getenv("ZGPU_NO_RBFS")) use_rbfs = !(*v && *v != '0');
  if (const char* v = std::getenv("ZGPU_LOOKUP_BATCH_CAP")) lb_cap_ = std::max<uint64_t>(1024, std::strtoull(v, nullptr, 10));
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
           " (libzgpu has no CPU fallback)";
  if (dev < 0) {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  }
  if (dev >= count) return "CUDA device ordinal out of range";
  if ((e = cudaSetDevice(dev)) != cudaSuccess) return std::string("cudaSetDevice: ") + cudaGetErrorString(e);
  device = dev;
  cudaDeviceProp prop{};
  if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
    return std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
  sm_count_ = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess)
    return std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
  cudaEventCreate(&ev0_);
  cudaEventCreate(&ev1_);
  cudaEventCreateWithFlags(&last_done_, cudaEventDisableTiming);
  if ((e = cudaStreamCreateWithFlags(&cstream_, cudaStreamNonBlocking)) != cudaSuccess)
    return std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
  cudaEventCreateWithFlags(&ev_reset_, cudaEventDisableTiming);
  if (const char* v = std::getenv("ZGPU_NO_STREAM_H2D")) stream_h2d = !(*v && *v != '0');
  if (subq_cap) subq_cap_ = subq_cap;
  if (budget) budget_ = budget;
  if (!ctrl_.ensure(128)) return "out of device memory";
  cudaMemset(ctrl_.p, 0, 128);
  return "";
}

std::string Device::publish(const HostSnapshot& h, const Schema& sc, uint64_t revision) {
  cudaSetDevice(device);
  auto s = std::make_shared<Snapshot>();
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> bool {
    if (!b.ensure(bytes + 64)) return false;  // + 64: rows are streamed with aligned 128-bit loads (build.cu)
    if (bytes && cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) return false;
    s->bytes += bytes;
    return true;
  };
  bool ok = up(s->row_ptr, h.row_ptr.data(), h.row_ptr.size() * 4) && up(s->col, h.col.data(), h.col.size() * 4) &&
            up(s->rrow_ptr, h.rrow_ptr.data(), h.rrow_ptr.size() * 4) && up(s->rcol, h.rcol.data(), h.rcol.size() * 4);
  if (ok && sc.has_expiry) ok = up(s->exp, h.exp.data(), h.exp.size() * 4);
  s->resources.resize(h.resources.size());
  s->n_resources.resize(h.resources.size());
  for (size_t t = 0; ok && t < h.resources.size(); ++t) {
    s->n_resources[t] = h.resources[t].size();
    ok = up(s->resources[t], h.resources[t].data(), h.resources[t].size() * 4);
  }
  if (!ok || cudaStreamSynchronize(stream) != cudaSuccess)
    return std::string("snapshot upload failed: ") + cudaGetErrorString(cudaGetLastError());
  s->n_tuples = h.n_tuples;
  return finish_publish(s, h, sc, revision);
}

std::string Device::publish_gpu(const Store& store, const Schema& sc, uint64_t revision, bool verify) {
  cudaSetDevice(device);
  HostSnapshot lay = store.layout();
  auto s = std::make_shared<Snapshot>();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  std::string err = gpu_build_snapshot(store, sc, &lay, stream, s.get());
  cudaEventRecord(e1, stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  last_build_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!err.empty()) return "GPU snapshot build failed: " + err;
  ++full_publishes;
  if (verify) {
    std::string verr = verify_snapshot(*s, store, sc, &lay);
    if (!verr.empty()) return verr;
  }
  return finish_publish(s, lay, sc, revision);
}

// test mode (ZGPU_VERIFY_BUILD=1): the host builder must produce byte-identical arrays
std::string Device::verify_snapshot(Snapshot& sn, const Store& store, const Schema& sc, const HostSnapshot* lay) {
  Snapshot* s = &sn;
  if (s->resources_stale) {
    std::string rerr = gpu_resource_lists(s, stream);
    if (!rerr.empty()) return rerr;
  }
  cudaStreamSynchronize(stream);
  HostSnapshot h = store.build();
  if (!h.err.empty()) return h.err;
  auto same = [&](const DevBuf& d, const std::vector<uint32_t>& v, const char* what) -> std::string {
    std::vector<uint32_t> got(v.size());
    if (!v.empty() && cudaMemcpy(got.data(), d.p, v.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
      return std::string("verify: cannot read ") + what;
    for (size_t i = 0; i < v.size(); ++i)
      if (got[i] != v[i])
        return std::string("verify: ") + what + " differs at " + std::to_string(i) + " (gpu " + std::to_string(got[i]) +
               ", host " + std::to_string(v[i]) + ")";
    return "";
  };
  if (h.n_tuples != s->n_tuples) return "verify: relationship count differs";
  for (auto m : {same(s->row_ptr, h.row_ptr, "row_ptr"), same(s->col, h.col, "col"), same(s->rrow_ptr, h.rrow_ptr, "rrow_ptr"),
                 same(s->rcol, h.rcol, "rcol"), sc.has_expiry ? same(s->exp, h.exp, "exp") : std::string()})
    if (!m.empty()) return m;
  for (size_t t = 0; t < h.resources.size(); ++t) {
    if (h.resources[t].size() != s->n_resources[t]) return "verify: resource list size differs";
    std::string m = same(s->resources[t], h.resources[t], "resources");
    if (!m.empty()) return m;
  }
  for (size_t c = 0; c < h.cls.size(); ++c) {
    const bool empty = lay ? (lay->cls[c].flags & CF_EMPTY) != 0 : s->cls_count[c] == 0;
    if (((h.cls[c].flags & CF_EMPTY) != 0) != empty) return "verify: class emptiness differs";
  }
  return "";
}
std::string Device::verify_against_host(const Store& store, const Schema& sc) {
  std::shared_ptr<Snapshot> s = snap;
  if (!s) return "verify: no snapshot";
  cudaSetDevice(device);
  return verify_snapshot(*s, store, sc, nullptr);
}

std::string Device::publish_delta(const Store& store, const Schema& sc, uint64_t revision) {
  constexpr size_t kMaxDelta = 1u << 16;
  std::shared_ptr<Snapshot> s = snap;
  if (!s || !s->delta_ok || !store.journal_ok || shard_count > 1 || store.journal.size() > kMaxDelta || !store.layout_stable())
    return "full";
  cudaSetDevice(device);
  HostSnapshot lay = store.layout();  // same capacities as the resident snapshot: same bases
  if (lay.n_objects != s->n_objects) return "full";
  // readers on other streams (zg_check_bulk_device) are ordered through last_done_
  if (have_last_ && cudaStreamWaitEvent(stream, last_done_, 0) != cudaSuccess) return "full";
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  std::vector<uint32_t> cls_delta;
  std::string err = gpu_apply_delta(store, sc, lay, stream, s.get(), &cls_delta);
  cudaEventRecord(e1, stream);
  cudaEventRecord(last_done_, stream);
  have_last_ = true;
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  last_build_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!err.empty()) {
    snap.reset();  // the arrays may be half merged: nobody may read them; the caller rebuilds
    return err == "relayout" || err == "corrupt" ? "full" : err;
  }
  bool structural = false;
  for (size_t c = 0; c < cls_delta.size(); ++c) {
    const uint64_t before = s->cls_count[c];
    s->cls_count[c] = before + static_cast<int32_t>(cls_delta[c]);
    if ((before == 0) != (s->cls_count[c] == 0)) structural = true;
  }
  s->resources_stale = true;
  s->revision = revision;
  ++delta_publishes;
  if (!structural) return "";
  // a class became empty or non-empty: the flattened steps of the program change
  for (size_t c = 0; c < lay.cls.size(); ++c)
    if (s->cls_count[c]) lay.cls[c].flags &= static_cast<uint16_t>(~CF_EMPTY);
  lay.n_tuples = s->n_tuples;
  return finish_publish(s, lay, sc, revision);
}

std::string Device::finish_publish(std::shared_ptr<Snapshot> s, const HostSnapshot& h, const Schema& sc, uint64_t revision) {
  std::vector<uint8_t> blob = sc.blob(h.rels, h.cls);
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> bool {
    if (!b.ensure(bytes ? bytes : 16)) return false;
    if (bytes && cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) return false;
    s->bytes += bytes;
    return true;
  };
  bool ok = up(s->prog, blob.data(), blob.size());
  {
    std::vector<unsigned long long> base(h.n_objects.size() + 1, 0);
    for (size_t t = 0; t < h.n_objects.size(); ++t) base[t + 1] = base[t] + ((uint64_t(h.n_objects[t]) + 31) & ~31ull);
    s->total_bits = base.back();
    s->n_objects = h.n_objects;
    if (ok) ok = up(s->type_bit_base, base.data(), base.size() * 8);
  }
  if (!ok || cudaStreamSynchronize(stream) != cudaSuccess)
    return std::string("snapshot upload failed: ") + cudaGetErrorString(cudaGetLastError());
  // flat-union slots: LookupResources is the union of the subject's reverse rows
  {
    const size_t ns = sc.slots.size();
    s->flat_cls.resize(ns);
    s->flat_n.assign(ns, -1);
    for (size_t sl = 0; ok && sl < ns; ++sl) {
      if (sc.d_slots[sl].kind == SK_NONPURE) continue;
      const DUnit& u = sc.d_units[sc.d_slots[sl].unit];
      bool flat = true;
      std::vector<FlatLookupClass> fc;
      for (int oi = u.op_begin; oi < u.op_end && flat; ++oi) {
        const DOp& op = sc.d_ops[oi];
        if (op.kind != OP_REL) { flat = false; break; }
        const DRel& r = h.rels[op.rel];
        for (uint16_t k = 0; k < r.ncls; ++k) {
          const DCls& c = h.cls[r.cls_begin + k];
          if (c.flags & CF_EMPTY) continue;
          if ((c.sslot != kNone && c.sslot != kWildcard) || (c.flags & CF_EXPIRY)) { flat = false; break; }
          fc.push_back(FlatLookupClass{c.rrow_base, c.nsubj, c.stype, static_cast<uint16_t>(c.sslot == kWildcard), c.rstride, 0});
        }
      }
      if (!flat) continue;
      s->flat_n[sl] = static_cast<int>(fc.size());
      ok = up(s->flat_cls[sl], fc.data(), fc.size() * sizeof(FlatLookupClass));
    }
    if (ok && cudaStreamSynchronize(stream) != cudaSuccess) ok = false;
    if (!ok) return std::string("snapshot upload failed (flat lookup tables)");
  }
  s->prog_bytes = static_cast<uint32_t>(blob.size());
  s->max_leaves = sc.max_leaves;
  s->has_nonpure = sc.has_nonpure;
  s->revision = revision;

  // occupancy for this program size (dynamic shared memory = program + warp stacks)
  size_t sm = smem_bytes(s->prog_bytes);
  if (sm > 200 * 1024) return "schema program too large for shared memory";
  {
    // The attribute is per function, i.e. shared by every engine of the process (several
    // engines / shards may publish programs of different sizes): only ever raise it.
    static std::mutex attr_mu;
    static size_t attr_by_device[64] = {0};
    size_t& attr_bytes = attr_by_device[device & 63];
    std::lock_guard<std::mutex> g(attr_mu);
    if (sm > attr_bytes) {
      cudaFuncSetAttribute(check_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
      cudaFuncSetAttribute(check_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
      cudaFuncSetAttribute(check_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
      attr_bytes = sm;
    }
  }
  int occ = 0, occ_c = 0, occ_s = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, check_kernel<false>, kThreads, sm);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, check_kernel<false, true>, kThreads, sm);
  if (occ_s < occ) occ = occ_s;  // one grid size serves both
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, check_kernel<true>, kThreads, sm);
  if (occ < 1 || occ_c < 1) return "check kernel cannot be resident (occupancy 0)";
  blocks_per_sm_ = occ;
  blocks_per_sm_count_ = occ_c;
  size_t warps = static_cast<size_t>(sm_count_) * std::max(occ, occ_c) * kWarpsPerBlock;
  if (!spill_.ensure(warps * spill_cap_ * sizeof(uint4))) return "out of device memory (spill areas)";
  if (const char* v = std::getenv("ZGPU_MEMO_AFTER")) memo_after_ = static_cast<uint32_t>(std::atoi(v));
  if (!memo_.ensure(warps * memo_entries_ * sizeof(unsigned long long))) memo_entries_ = 0;  // optional
  // readers that already hold the old snapshot keep it alive until they finish
  snap = s;
  return "";
}

int Device::run_pass(const Snapshot& s, const zg_check* queries, uint64_t nq, uint8_t* val, uint8_t* out, bool final_codes,
                     bool raw, zg_check* subq, uint32_t* subq_parent, cudaStream_t st, bool count, std::string* err,
                     const unsigned long long* ready) {
  KParams p{};
  p.ready = ready;
  p.row_ptr = s.row_ptr.as<uint32_t>();
  p.col = s.col.as<uint32_t>();
  p.exp = s.exp.p ? s.exp.as<uint32_t>() : nullptr;
  p.rrow_ptr = s.rrow_ptr.as<uint32_t>();
  p.rcol = s.rcol.as<uint32_t>();
  p.invert = (invert && shard_count <= 1) ? 1 : 0;  // a shard does not hold the subject's reverse rows
  p.prog = s.prog.as<uint8_t>();
  p.prog_bytes = s.prog_bytes;
  p.queries = queries;
  p.nq = nq;
  p.L = s.max_leaves;
  p.val = val;
  p.out = out;
  p.final_codes = final_codes;
  unsigned long long* ctrl = ctrl_.as<unsigned long long>();
  p.next = ctrl;
  p.subq_count = ctrl + 1;
  p.alg_bytes = ctrl + 2;
  p.flags = reinterpret_cast<uint32_t*>(ctrl + 3);
  p.events = ctrl + 8;
  p.spill = spill_.as<uint4>();
  p.spill_cap = spill_cap_;
  p.subq = subq;
  p.subq_parent = subq_parent;
  p.subq_cap = subq ? subq_cap_ : 0;
  p.now = now;
  p.budget = budget_;
  p.raw_items = raw;
  p.memo = memo_.as<unsigned long long>();
  p.memo_entries = memo_.p ? memo_entries_ : 0;
  p.memo_after = memo_after_;
  p.shard_count = shard_count;
  p.shard_rank = shard_rank;
  ZG_CUDA(cudaMemsetAsync(ctrl, 0, 16, st));  // next, subq_count
  const int per_sm = count ? blocks_per_sm_count_ : blocks_per_sm_;
  uint64_t want = (nq + kThreads - 1) / kThreads;
  int grid = static_cast<int>(std::min<uint64_t>(want, static_cast<uint64_t>(sm_count_) * per_sm));
  if (grid < 1) grid = 1;
  size_t sm = smem_bytes(s.prog_bytes);
  if (count) check_kernel<true><<<grid, kThreads, sm, st>>>(p);
  else if (ready) check_kernel<false, true><<<grid, kThreads, sm, st>>>(p);
  else check_kernel<false><<<grid, kThreads, sm, st>>>(p);
  ZG_CUDA(cudaGetLastError());
  ++launches;
  return ZG_OK;
}

void Device::read_events(uint64_t* spills, uint64_t* memo_batches) {
  unsigned long long ev[2] = {0, 0};
  if (ctrl_.p && cudaSetDevice(device) == cudaSuccess &&
      cudaMemcpyAsync(ev, ctrl_.as<unsigned long long>() + 8, sizeof ev, cudaMemcpyDeviceToHost, stream) == cudaSuccess)
    cudaStreamSynchronize(stream);
  *spills = ev[0];
  *memo_batches = ev[1];
}

void Device::finish_timing() {
  if (!timing_pending_) return;
  if (cudaEventSynchronize(ev1_) == cudaSuccess) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev0_, ev1_) == cudaSuccess) last_ms = ms;
  }
  timing_pending_ = false;
}

int Device::check_device(const zg_check* d_items, uint64_t n, uint8_t* d_out, cudaStream_t st, bool raw_items,
                         uint64_t* count_bytes, std::string* err, bool top, const unsigned long long* ready) {
  std::shared_ptr<Snapshot> s = snap;
  if (!s) {
    if (err) *err = "no snapshot published (call zg_publish after loading relationships)";
    return ZG_ENOSNAPSHOT;
  }
  if (n == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  // st == NULL is the caller's legacy default stream, not the engine's own stream.
  // Scratch (ctrl_, spill_, pass buffers) is shared by every call: order this call
  // after the previous one even when they were enqueued on different streams.
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(st, last_done_, 0));
  const bool count = count_bytes != nullptr;
  unsigned long long* ctrl = ctrl_.as<unsigned long long>();
  // alg_bytes and the sticky flags (spill overflow, budget hit) are cleared once per caller-level
  // call: the halves of a split batch accumulate into them
  if (top) {
    ZG_CUDA(cudaMemsetAsync(ctrl + 2, 0, 16, st));
    cudaEventRecord(ev0_, st);
  }
  const uint32_t L = s->max_leaves;
  int rc = ZG_OK;
  if (!s->has_nonpure) {
    // every slot is a relation or a pure union: one leaf per check, answered in one launch
    rc = run_pass(*s, d_items, n, nullptr, d_out, true, raw_items, nullptr, nullptr, st, count, err, ready);
    if (rc) return rc;
  } else {
    if (n * L >= (1ull << 32)) {
      if (err) *err = "batch too large for one call (n * leaves >= 2^32)";
      return ZG_EINVAL;
    }
    std::vector<uint64_t> nq;  // queries per pass level
    auto level = [&](size_t lv) {
      if (q_.size() <= lv) {
        q_.resize(lv + 1);
        parent_.resize(lv + 1);
        val_.resize(lv + 1);
      }
    };
    const zg_check* queries = d_items;
    uint64_t cur = n;
    bool raw = raw_items;
    for (size_t lv = 0; cur > 0; ++lv) {
      if (lv > ZG_MAX_DEPTH + 1) {
        if (err) *err = "sub-query passes exceeded the depth cap";
        return ZG_ECUDA;
      }
      level(lv + 1);
      nq.push_back(cur);
      if (cur * L >= (1ull << 32)) {
        if (err) *err = "pass too large (queries * leaves >= 2^32)";
        return ZG_EINVAL;
      }
      if (!val_[lv].ensure(cur * L + 4) || !q_[lv + 1].ensure(subq_cap_ * sizeof(zg_check)) ||
          !parent_[lv + 1].ensure(subq_cap_ * 4)) {
        if (err) *err = "out of device memory (pass buffers)";
        return ZG_ENOMEM;
      }
      // leaves the kernel skips (short circuit of the boolean tree) must read as "false"
      ZG_CUDA(cudaMemsetAsync(val_[lv].p, 0, cur * L + 4, st));
      // Level 0 writes the v1 codes straight into d_out: when the pass raises no sub-query (the
      // common case: edges into a non-pure permission are rare) they are final and nothing is folded.
      rc = run_pass(*s, queries, cur, val_[lv].as<uint8_t>(), lv == 0 ? d_out : nullptr, true, raw,
                    q_[lv + 1].as<zg_check>(), parent_[lv + 1].as<uint32_t>(), st, count, err, lv == 0 ? ready : nullptr);
      if (rc) return rc;
      unsigned long long host_ctrl[4];
      ZG_CUDA(cudaMemcpyAsync(host_ctrl, ctrl, sizeof host_ctrl, cudaMemcpyDeviceToHost, st));
      ZG_CUDA(cudaStreamSynchronize(st));
      if (host_ctrl[1] > subq_cap_) {
        // more sub-queries than the pass buffer holds: checks are independent, so answer the
        // batch in two halves (each raises about half as many) instead of failing the call
        if (n < 2) {
          if (err) *err = "sub-query buffer overflow on a single check: raise zg_config.subquery_capacity";
          return ZG_ENOMEM;
        }
        const uint64_t half = n / 2;
        ++split_batches;
        int r1 = check_device(d_items, half, d_out, st, raw_items, nullptr, err, false);
        if (r1) return r1;
        return check_device(d_items + half, n - half, d_out + half, st, raw_items, nullptr, err, false);
      }
      cur = host_ctrl[1];
      queries = q_[lv + 1].as<zg_check>();
      raw = false;
      if (cur) ++passes;
    }
    if (nq.size() > 1) {
      for (size_t lv = nq.size(); lv-- > 0;) {
        const unsigned blk = 256;
        const uint64_t m = nq[lv];
        fold_kernel<<<static_cast<unsigned>((m + blk - 1) / blk), blk, 0, st>>>(
            s->prog.as<uint8_t>(), lv == 0 ? d_items : q_[lv].as<zg_check>(), m, L, val_[lv].as<uint8_t>(),
            lv == 0 ? d_out : nullptr, lv == 0 ? nullptr : parent_[lv].as<uint32_t>(),
            lv == 0 ? nullptr : val_[lv - 1].as<uint8_t>(), 0);
        ++launches;
      }
      ZG_CUDA(cudaGetLastError());
    }
  }
  if (top) {
    cudaEventRecord(ev1_, st);
    timing_pending_ = true;
  }
  cudaEventRecord(last_done_, st);
  have_last_ = true;
  checks += n;
  if (count) {
    unsigned long long host_ctrl[4];
    ZG_CUDA(cudaMemcpyAsync(host_ctrl, ctrl, sizeof host_ctrl, cudaMemcpyDeviceToHost, st));
    ZG_CUDA(cudaStreamSynchronize(st));
    *count_bytes = host_ctrl[2];
    last_alg_bytes = host_ctrl[2];
  }
  return ZG_OK;
}

int Device::check_host(const zg_check* items, uint64_t n, uint8_t* out, std::string* err) {
  std::vector<HostReq> one{{items, n, out}};
  return check_host_multi(one, err);
}

// One launch sequence for the requests of one or several callers (see Batcher in capi.cu).
//
// Host buffers -> answers, copies included. Large inputs are STREAMED: the items go up in chunks on a copy
// stream, each chunk followed by an 8-byte copy of "items landed so far" into the word the kernel polls, and
// the check kernel -- launched once, persistent -- starts on the first chunk while the rest is still on the
// PCIe bus (warps wait for their batch: KParams::ready). H2D and the kernel overlap without per-chunk
// launches, whose tails cost more than the copies ("tried and reverted").
int Device::check_host_multi(const std::vector<HostReq>& reqs, std::string* err) {
  uint64_t total = 0;
  for (const auto& r : reqs) total += r.n;
  if (total == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  if (!stage_in_.ensure(total * sizeof(zg_check)) || !stage_out_.ensure(total)) {
    if (err) *err = "out of device memory (staging)";
    return ZG_ENOMEM;
  }
  // Pinned caller buffers (zg_host_alloc) are copied straight from/to; pageable ones and the requests of
  // several callers go through the engine's pinned staging area.
  cudaPointerAttributes a{};
  const bool single = reqs.size() == 1;
  const bool in_pinned = single && cudaPointerGetAttributes(&a, reqs[0].items) == cudaSuccess && a.type == cudaMemoryTypeHost;
  const bool out_pinned = single && cudaPointerGetAttributes(&a, reqs[0].out) == cudaSuccess && a.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (!in_pinned && pin_in_cap_ < total * sizeof(zg_check)) {
    if (pin_in_) cudaFreeHost(pin_in_);
    pin_in_cap_ = 0;
    ZG_CUDA(cudaMallocHost(&pin_in_, total * sizeof(zg_check)));
    pin_in_cap_ = total * sizeof(zg_check);
  }
  if (!out_pinned && pin_out_cap_ < total) {
    if (pin_out_) cudaFreeHost(pin_out_);
    pin_out_cap_ = 0;
    ZG_CUDA(cudaMallocHost(&pin_out_, total));
    pin_out_cap_ = total;
  }
  unsigned long long* d_ready = ctrl_.as<unsigned long long>() + 10;
  const bool streamed = stream_h2d && total >= kStreamMinItems;
  const unsigned long long* ready_arg = nullptr;
  // A large group of many callers (1 000 lists of 10 000 items under BASELINE config 5): one thread gathering 16 B per
  // item into the staging area takes longer than the kernel that answers them. Helpers copy contiguous shares.
  bool gathered = false;
  if (!in_pinned && total >= kParallelGatherItems && reqs.size() >= 2 * kGatherThreads) {
    std::vector<uint64_t> start(reqs.size() + 1, 0);
    for (size_t i = 0; i < reqs.size(); ++i) start[i + 1] = start[i] + reqs[i].n;
    auto share = [&](unsigned t) {
      const size_t b = reqs.size() * t / kGatherThreads, e = reqs.size() * (t + 1) / kGatherThreads;
      for (size_t i = b; i < e; ++i)
        std::memcpy(static_cast<zg_check*>(pin_in_) + start[i], reqs[i].items, reqs[i].n * sizeof(zg_check));
    };
    std::thread helpers[kGatherThreads - 1];
    for (unsigned t = 1; t < kGatherThreads; ++t) helpers[t - 1] = std::thread(share, t);
    share(0);
    for (auto& h : helpers) h.join();
    gathered = true;
  }
  if (!streamed) {
    const void* src = reqs[0].items;
    if (gathered) {
      src = pin_in_;
    } else if (!in_pinned) {
      uint64_t off = 0;
      for (const auto& r : reqs) {
        std::memcpy(static_cast<zg_check*>(pin_in_) + off, r.items, r.n * sizeof(zg_check));
        off += r.n;
      }
      src = pin_in_;
    }
    ZG_CUDA(cudaMemcpyAsync(stage_in_.p, src, total * sizeof(zg_check), cudaMemcpyHostToDevice, stream));
  } else {
    // copy stream: after the previous call (it may still read stage_in_ / the ready word), reset the word,
    // then chunk, count, chunk, count ... ; the compute stream only waits for the reset
    const uint64_t n_chunks = (total + kStreamChunkItems - 1) / kStreamChunkItems;
    if (pin_ready_cap_ < n_chunks) {
      if (pin_ready_) cudaFreeHost(pin_ready_);
      pin_ready_cap_ = 0;
      ZG_CUDA(cudaMallocHost(&pin_ready_, n_chunks * 8));
      pin_ready_cap_ = n_chunks;
    }
    if (have_last_) ZG_CUDA(cudaStreamWaitEvent(cstream_, last_done_, 0));
    ZG_CUDA(cudaMemsetAsync(d_ready, 0, 8, cstream_));
    ZG_CUDA(cudaEventRecord(ev_reset_, cstream_));
    ZG_CUDA(cudaStreamWaitEvent(stream, ev_reset_, 0));
    size_t ri = 0;          // request being gathered
    uint64_t roff = 0;      // items of it already gathered
    for (uint64_t ck = 0; ck < n_chunks; ++ck) {
      const uint64_t b = ck * kStreamChunkItems, e = std::min(total, b + kStreamChunkItems);
      const zg_check* src;
      if (in_pinned) {
        src = reqs[0].items + b;
      } else if (gathered) {
        src = static_cast<const zg_check*>(pin_in_) + b;
      } else {
        zg_check* dst = static_cast<zg_check*>(pin_in_) + b;
        uint64_t need = e - b, w = 0;
        while (need) {
          const uint64_t take = std::min(need, reqs[ri].n - roff);
          std::memcpy(dst + w, reqs[ri].items + roff, take * sizeof(zg_check));
          w += take;
          roff += take;
          need -= take;
          if (roff == reqs[ri].n) {
            ++ri;
            roff = 0;
          }
        }
        src = dst;
      }
      ZG_CUDA(cudaMemcpyAsync(stage_in_.as<zg_check>() + b, src, (e - b) * sizeof(zg_check), cudaMemcpyHostToDevice, cstream_));
      pin_ready_[ck] = e;
      ZG_CUDA(cudaMemcpyAsync(d_ready, pin_ready_ + ck, 8, cudaMemcpyHostToDevice, cstream_));
    }
    ready_arg = d_ready;
    ++streamed_calls;
  }
  int rc = check_device(stage_in_.as<zg_check>(), total, stage_out_.as<uint8_t>(), stream, true, nullptr, err, true, ready_arg);
  if (rc) {
    if (streamed) cudaStreamSynchronize(cstream_);  // nothing of this call may still be in flight
    return rc;
  }
  void* dst = out_pinned ? static_cast<void*>(reqs[0].out) : pin_out_;
  ZG_CUDA(cudaMemcpyAsync(dst, stage_out_.p, total, cudaMemcpyDeviceToHost, stream));
  uint32_t flags = 0;
  ZG_CUDA(cudaMemcpyAsync(&flags, ctrl_.as<unsigned long long>() + 3, 4, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  if (streamed) ZG_CUDA(cudaStreamSynchronize(cstream_));
  if (flags & 32u) {
    if (err) *err = "host-to-device copy of the request stalled (streamed admission timed out)";
    return ZG_ECUDA;
  }
  if (flags & 1u) {
    if (err) *err = "expansion stack overflow (per-warp spill area exhausted)";
    return ZG_ENOMEM;
  }
  if (!out_pinned) {
    uint64_t off = 0;
    for (const auto& r : reqs) {
      std::memcpy(r.out, static_cast<uint8_t*>(pin_out_) + off, r.n);
      off += r.n;
    }
  }
  if (!single) {
    ++coalesced_launches;
    coalesced_requests += reqs.size();
  }
  return ZG_OK;
}

// ---- sharded store: one pass per call; the host exchanges the raised sub-queries -----------
//
// Buffers per level lv: q_[lv] = this level's queries, val_[lv] = their leaf values,
// q_[lv+1] / parent_[lv+1] = the sub-queries the pass raised (the same buffers the single-GPU
// multi-pass loop uses; here their consumers live on other ranks).
int Device::shard_pass(const zg_check* queries, uint64_t n, int level, uint64_t* n_sub, std::string* err) {
  std::shared_ptr<Snapshot> s = snap;
  if (!s) {
    if (err) *err = "no snapshot published";
    return ZG_ENOSNAPSHOT;
  }
  ZG_CUDA(cudaSetDevice(device));
  const size_t lv = static_cast<size_t>(level);
  if (q_.size() <= lv + 1) {
    q_.resize(lv + 2);
    parent_.resize(lv + 2);
    val_.resize(lv + 2);
  }
  if (shard_nq_.size() <= lv) {
    shard_nq_.resize(lv + 1);
    shard_nsub_.resize(lv + 1);
  }
  shard_nq_[lv] = n;
  shard_nsub_[lv] = 0;
  *n_sub = 0;
  if (n == 0) return ZG_OK;
  const uint32_t L = s->max_leaves;
  if (n * L >= (1ull << 32)) {
    if (err) *err = "pass too large (n * leaves >= 2^32)";
    return ZG_EINVAL;
  }
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  // level > 0 queries arrive in a host buffer too; level 0 is staged apart from the raised-sub-query buffer q_[1]
  DevBuf& qbuf = lv == 0 ? stage_in_ : q_[lv];  // q_[lv] (lv > 0) was the producer-side buffer of level lv-1 on
                                              // THIS rank; its content was already read back by shard_subqueries
  if (!qbuf.ensure(std::max<size_t>(n * sizeof(zg_check), subq_cap_ * sizeof(zg_check))) || !val_[lv].ensure(n * L + 4) ||
      !q_[lv + 1].ensure(subq_cap_ * sizeof(zg_check)) || !parent_[lv + 1].ensure(subq_cap_ * 4)) {
    if (err) *err = "out of device memory (shard pass buffers)";
    return ZG_ENOMEM;
  }
  ZG_CUDA(cudaMemcpyAsync(qbuf.p, queries, n * sizeof(zg_check), cudaMemcpyHostToDevice, stream));
  unsigned long long* ctrl = ctrl_.as<unsigned long long>();
  ZG_CUDA(cudaMemsetAsync(ctrl + 2, 0, 16, stream));
  ZG_CUDA(cudaMemsetAsync(val_[lv].p, 0, n * L + 4, stream));
  int rc = run_pass(*s, qbuf.as<zg_check>(), n, val_[lv].as<uint8_t>(), nullptr, false, level == 0, q_[lv + 1].as<zg_check>(),
                    parent_[lv + 1].as<uint32_t>(), stream, false, err);
  if (rc) return rc;
  unsigned long long host_ctrl[4];
  ZG_CUDA(cudaMemcpyAsync(host_ctrl, ctrl, sizeof host_ctrl, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaEventRecord(last_done_, stream));
  have_last_ = true;
  ZG_CUDA(cudaStreamSynchronize(stream));
  const uint32_t flags = static_cast<uint32_t>(host_ctrl[3] & 0xFFFFFFFFu);
  if (host_ctrl[1] > subq_cap_) {
    if (err) *err = "sub-query buffer overflow: raise zg_config.subquery_capacity or shrink the batch";
    return ZG_ENOMEM;
  }
  if (flags & 1u) {
    if (err) *err = "expansion stack overflow";
    return ZG_ENOMEM;
  }
  shard_nsub_[lv] = host_ctrl[1];
  *n_sub = host_ctrl[1];
  checks += level == 0 ? n : 0;
  ++passes;
  return ZG_OK;
}

int Device::shard_subqueries(int level, zg_check* out, uint64_t n, std::string* err) {
  const size_t lv = static_cast<size_t>(level);
  if (lv >= shard_nsub_.size() || n != shard_nsub_[lv]) {
    if (err) *err = "shard_subqueries: level / count mismatch";
    return ZG_EINVAL;
  }
  if (n == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  ZG_CUDA(cudaMemcpyAsync(out, q_[lv + 1].p, n * sizeof(zg_check), cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  return ZG_OK;
}

int Device::shard_fold(int level, const uint8_t* child_vals, uint64_t n_sub, uint8_t* out, std::string* err) {
  std::shared_ptr<Snapshot> s = snap;
  const size_t lv = static_cast<size_t>(level);
  if (!s || lv >= shard_nq_.size() || n_sub != shard_nsub_[lv]) {
    if (err) *err = "shard_fold: level / count mismatch";
    return ZG_EINVAL;
  }
  const uint64_t n = shard_nq_[lv];
  if (n == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  const unsigned blk = 256;
  if (n_sub) {
    if (!shard_tmp_.ensure(n_sub)) return ZG_ENOMEM;
    ZG_CUDA(cudaMemcpyAsync(shard_tmp_.p, child_vals, n_sub, cudaMemcpyHostToDevice, stream));
    or_children_kernel<<<static_cast<unsigned>((n_sub + blk - 1) / blk), blk, 0, stream>>>(
        parent_[lv + 1].as<uint32_t>(), shard_tmp_.as<uint8_t>(), n_sub, val_[lv].as<uint8_t>());
    ++launches;
  }
  if (!stage_out_.ensure(n)) return ZG_ENOMEM;
  const DevBuf& qbuf = lv == 0 ? stage_in_ : q_[lv];
  fold_kernel<<<static_cast<unsigned>((n + blk - 1) / blk), blk, 0, stream>>>(
      s->prog.as<uint8_t>(), qbuf.as<zg_check>(), n, s->max_leaves, val_[lv].as<uint8_t>(), stage_out_.as<uint8_t>(),
      nullptr, nullptr, level == 0 ? 0 : 1);
  ++launches;
  ZG_CUDA(cudaMemcpyAsync(out, stage_out_.p, n, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  return ZG_OK;
}

int Device::route_by_owner(const zg_check* d_items, uint64_t n, int level, uint32_t n_dest, zg_check* d_routed, uint32_t* d_src,
                           uint64_t* counts, std::string* err) {
  if (n_dest == 0 || n_dest > kMaxRouteDest) {
    if (err) *err = "route_by_owner: destination count out of range";
    return ZG_EINVAL;
  }
  for (uint32_t d = 0; d < n_dest; ++d) counts[d] = 0;
  ZG_CUDA(cudaSetDevice(device));
  if (!d_items) {
    const size_t lv = static_cast<size_t>(level);
    if (lv >= shard_nsub_.size() || n != shard_nsub_[lv]) {
      if (err) *err = "route_by_owner: level / count mismatch";
      return ZG_EINVAL;
    }
    d_items = q_[lv + 1].as<zg_check>();
  }
  if (n == 0) return ZG_OK;
  if (!route_ctrl_.ensure(2 * kMaxRouteDest * 8)) return ZG_ENOMEM;
  unsigned long long* d_counts = route_ctrl_.as<unsigned long long>();
  unsigned long long* d_cursor = d_counts + kMaxRouteDest;
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  ZG_CUDA(cudaMemsetAsync(d_counts, 0, kMaxRouteDest * 8, stream));
  const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((n + 255) / 256, static_cast<uint64_t>(sm_count_) * 8));
  route_count_kernel<<<grid, 256, 0, stream>>>(d_items, n, n_dest, d_counts);
  unsigned long long h[kMaxRouteDest];
  ZG_CUDA(cudaMemcpyAsync(h, d_counts, n_dest * 8, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  unsigned long long off[kMaxRouteDest], acc = 0;
  for (uint32_t d = 0; d < n_dest; ++d) {
    counts[d] = h[d];
    off[d] = acc;
    acc += h[d];
  }
  ZG_CUDA(cudaMemcpyAsync(d_cursor, off, n_dest * 8, cudaMemcpyHostToDevice, stream));
  route_scatter_kernel<<<grid, 256, 0, stream>>>(d_items, n, n_dest, d_cursor, d_routed, d_src);
  launches += 2;
  ZG_CUDA(cudaGetLastError());
  ZG_CUDA(cudaEventRecord(last_done_, stream));
  have_last_ = true;
  ZG_CUDA(cudaStreamSynchronize(stream));
  return ZG_OK;
}

int Device::shard_pass_dev(const zg_check* d_queries, uint64_t n, int level, uint64_t* n_sub, std::string* err) {
  std::shared_ptr<Snapshot> s = snap;
  if (!s) {
    if (err) *err = "no snapshot published";
    return ZG_ENOSNAPSHOT;
  }
  ZG_CUDA(cudaSetDevice(device));
  const size_t lv = static_cast<size_t>(level);
  if (q_.size() <= lv + 1) {
    q_.resize(lv + 2);
    parent_.resize(lv + 2);
    val_.resize(lv + 2);
  }
  if (shard_nq_.size() <= lv) {
    shard_nq_.resize(lv + 1);
    shard_nsub_.resize(lv + 1);
  }
  shard_nq_[lv] = n;
  shard_nsub_[lv] = 0;
  *n_sub = 0;
  if (n == 0) return ZG_OK;
  const uint32_t L = s->max_leaves;
  if (n * L >= (1ull << 32)) {
    if (err) *err = "pass too large (n * leaves >= 2^32)";
    return ZG_EINVAL;
  }
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  // the level's queries are kept (the fold re-reads them): own copy, device to device
  DevBuf& qbuf = lv == 0 ? stage_in_ : q_[lv];
  if (!qbuf.ensure(std::max<size_t>(n * sizeof(zg_check), lv == 0 ? 0 : subq_cap_ * sizeof(zg_check))) ||
      !val_[lv].ensure(n * L + 4) || !q_[lv + 1].ensure(subq_cap_ * sizeof(zg_check)) || !parent_[lv + 1].ensure(subq_cap_ * 4)) {
    if (err) *err = "out of device memory (shard pass buffers)";
    return ZG_ENOMEM;
  }
  ZG_CUDA(cudaMemcpyAsync(qbuf.p, d_queries, n * sizeof(zg_check), cudaMemcpyDeviceToDevice, stream));
  unsigned long long* ctrl = ctrl_.as<unsigned long long>();
  ZG_CUDA(cudaMemsetAsync(ctrl + 2, 0, 16, stream));
  ZG_CUDA(cudaMemsetAsync(val_[lv].p, 0, n * L + 4, stream));
  // level 0 items come from callers (flags ignored); deeper levels are raised sub-queries (flags = hop depth)
  int rc = run_pass(*s, qbuf.as<zg_check>(), n, val_[lv].as<uint8_t>(), nullptr, false, level == 0, q_[lv + 1].as<zg_check>(),
                    parent_[lv + 1].as<uint32_t>(), stream, false, err);
  if (rc) return rc;
  unsigned long long host_ctrl[4];
  ZG_CUDA(cudaMemcpyAsync(host_ctrl, ctrl, sizeof host_ctrl, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaEventRecord(last_done_, stream));
  have_last_ = true;
  ZG_CUDA(cudaStreamSynchronize(stream));
  const uint32_t flags = static_cast<uint32_t>(host_ctrl[3] & 0xFFFFFFFFu);
  if (host_ctrl[1] > subq_cap_) {
    if (err) *err = "sub-query buffer overflow: raise zg_config.subquery_capacity or shrink the batch";
    return ZG_ENOMEM;
  }
  if (flags & 1u) {
    if (err) *err = "expansion stack overflow";
    return ZG_ENOMEM;
  }
  shard_nsub_[lv] = host_ctrl[1];
  *n_sub = host_ctrl[1];
  checks += level == 0 ? n : 0;
  ++passes;
  return ZG_OK;
}

int Device::shard_fold_dev(int level, const uint8_t* d_child_vals, const uint32_t* d_src, uint64_t n_sub, uint8_t* d_out,
                           bool final_codes, std::string* err) {
  std::shared_ptr<Snapshot> s = snap;
  const size_t lv = static_cast<size_t>(level);
  if (!s || lv >= shard_nq_.size() || n_sub != shard_nsub_[lv]) {
    if (err) *err = "shard_fold_dev: level / count mismatch";
    return ZG_EINVAL;
  }
  const uint64_t n = shard_nq_[lv];
  if (n == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  const unsigned blk = 256;
  if (n_sub) {
    or_children_src_kernel<<<static_cast<unsigned>((n_sub + blk - 1) / blk), blk, 0, stream>>>(
        parent_[lv + 1].as<uint32_t>(), d_src, d_child_vals, n_sub, val_[lv].as<uint8_t>());
    ++launches;
  }
  const DevBuf& qbuf = lv == 0 ? stage_in_ : q_[lv];
  fold_kernel<<<static_cast<unsigned>((n + blk - 1) / blk), blk, 0, stream>>>(
      s->prog.as<uint8_t>(), qbuf.as<zg_check>(), n, s->max_leaves, val_[lv].as<uint8_t>(), d_out, nullptr, nullptr,
      final_codes ? 0 : 1);
  ++launches;
  ZG_CUDA(cudaGetLastError());
  ZG_CUDA(cudaEventRecord(last_done_, stream));
  have_last_ = true;
  ZG_CUDA(cudaStreamSynchronize(stream));
  return ZG_OK;
}

int Device::unroute(const uint32_t* d_src, const uint8_t* d_val, uint64_t n, uint8_t* d_out, std::string* err) {
  if (n == 0) return ZG_OK;
  ZG_CUDA(cudaSetDevice(device));
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  unroute_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(d_src, d_val, n, d_out);
  ++launches;
  ZG_CUDA(cudaGetLastError());
  ZG_CUDA(cudaEventRecord(last_done_, stream));
  have_last_ = true;
  ZG_CUDA(cudaStreamSynchronize(stream));
  return ZG_OK;
}

int Device::lookup_candidates(const Snapshot& s, uint16_t res_type, const zg_check& proto, uint64_t* n_cand, bool* overflow,
                              std::string* err) {
  *overflow = false;
  *n_cand = 0;
  const size_t vis_bytes = static_cast<size_t>((s.total_bits + 31) / 32) * 4 + 16;
  if (!rb_visited_.ensure(vis_bytes) || !rb_front_[0].ensure(rb_cap_ * 8) || !rb_front_[1].ensure(rb_cap_ * 8) ||
      !rb_cand_.ensure(rb_cap_ * 4)) {
    *overflow = true;  // not enough memory for the BFS buffers: exhaustive scan instead
    return ZG_OK;
  }
  unsigned long long* ctrl = ctrl_.as<unsigned long long>();  // [4] next_count, [5] cand_count; flags at ctrl+3
  ZG_CUDA(cudaMemsetAsync(rb_visited_.p, 0, vis_bytes, stream));
  ZG_CUDA(cudaMemsetAsync(ctrl + 3, 0, 24, stream));
  const unsigned long long seed = (static_cast<unsigned long long>(proto.stype) << 32) | proto.subj;
  ZG_CUDA(cudaMemcpyAsync(rb_front_[0].p, &seed, 8, cudaMemcpyHostToDevice, stream));
  RbfsParams p{};
  p.rrow_ptr = s.rrow_ptr.as<uint32_t>();
  p.rcol = s.rcol.as<uint32_t>();
  p.prog = s.prog.as<uint8_t>();
  p.visited = rb_visited_.as<uint32_t>();
  p.type_bit_base = s.type_bit_base.as<unsigned long long>();
  p.want_type = res_type;
  p.cand = rb_cand_.as<uint32_t>();
  p.cand_count = ctrl + 5;
  p.cand_cap = rb_cap_;
  p.next_count = ctrl + 4;
  p.next_cap = rb_cap_;
  p.flags = reinterpret_cast<uint32_t*>(ctrl + 3);
  unsigned long long n_in = 1;
  int cur = 0;
  for (int level = 0; n_in > 0; ++level) {
    if (level > ZG_MAX_DEPTH + 1) break;  // deeper objects cannot be within the dispatch cap anyway
    p.frontier = rb_front_[cur].as<unsigned long long>();
    p.n_in = n_in;
    p.next = rb_front_[cur ^ 1].as<unsigned long long>();
    p.wildcard_level = (level == 0 && proto.srel == kNone) ? 1 : 0;
    ZG_CUDA(cudaMemsetAsync(ctrl + 4, 0, 8, stream));
    const unsigned blk = 256;
    const unsigned long long threads = n_in * 32;
    rbfs_expand_kernel<<<static_cast<unsigned>((threads + blk - 1) / blk), blk, 0, stream>>>(p);
    ++launches;
    unsigned long long host[3];
    ZG_CUDA(cudaMemcpyAsync(host, ctrl + 3, sizeof host, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaStreamSynchronize(stream));
    if (static_cast<uint32_t>(host[0]) & 16u) {
      *overflow = true;
      return ZG_OK;
    }
    n_in = host[1];
    *n_cand = host[2];
    cur ^= 1;
  }
  return ZG_OK;
}

int Device::lookup(uint16_t res_type, const zg_check& proto, std::vector<uint32_t>* ids, std::string* err) {
  ids->clear();
  std::shared_ptr<Snapshot> s = snap;
  if (!s) {
    if (err) *err = "no snapshot published";
    return ZG_ENOSNAPSHOT;
  }
  if (res_type >= s->resources.size()) {
    if (err) *err = "unknown resource type";
    return ZG_EINVAL;
  }
  ZG_CUDA(cudaSetDevice(device));
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  // flat union of direct relations and a relation-less subject: the answer is the union of
  // the subject's reverse rows (classes of other subject types contribute nothing)
  if (use_rbfs && proto.srel == kNone && proto.perm < s->flat_n.size() && s->flat_n[proto.perm] >= 0) {
    const int nc = s->flat_n[proto.perm];
    const uint64_t cap = rb_cap_;
    if (!rb_cand_.ensure(cap * 4 + 8)) return ZG_ENOMEM;
    // layout: [count u64][ids ...] so the count and the first ids come back in ONE copy
    unsigned long long* d_count = rb_cand_.as<unsigned long long>();
    uint32_t* d_ids = rb_cand_.as<uint32_t>() + 2;
    lookup_flat_kernel<<<1, 32, 0, stream>>>(s->rrow_ptr.as<uint32_t>(), s->rcol.as<uint32_t>(),
                                             s->flat_cls[proto.perm].as<FlatLookupClass>(), nc, proto.stype, proto.subj,
                                             d_ids, cap, d_count);
    ++launches;
    constexpr uint64_t kFirst = 4094;
    static thread_local std::vector<uint32_t> host;
    host.resize(2 + kFirst);
    ZG_CUDA(cudaMemcpyAsync(host.data(), rb_cand_.p, host.size() * 4, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaStreamSynchronize(stream));
    unsigned long long cnt = 0;
    std::memcpy(&cnt, host.data(), 8);
    if (cnt <= cap) {
      ++lookups_flat;
      ids->resize(cnt);
      const uint64_t first = std::min<uint64_t>(cnt, kFirst);
      if (first) std::memcpy(ids->data(), host.data() + 2, first * 4);
      if (cnt > first) {
        ZG_CUDA(cudaMemcpyAsync(ids->data() + first, d_ids + first, (cnt - first) * 4, cudaMemcpyDeviceToHost, stream));
        ZG_CUDA(cudaStreamSynchronize(stream));
      }
      std::sort(ids->begin(), ids->end());
      ids->erase(std::unique(ids->begin(), ids->end()), ids->end());
      return ZG_OK;
    }
    // more results than the buffer: fall through to the general path
  }
  // candidates: reverse BFS from the subject (superset of the answer), else every resource of the type
  if (s->resources_stale) {
    std::string rerr = gpu_resource_lists(s.get(), stream);
    if (!rerr.empty()) {
      if (err) *err = rerr;
      return ZG_ECUDA;
    }
  }
  uint64_t n = s->n_resources[res_type];
  const uint32_t* cand = s->resources[res_type].as<uint32_t>();
  if (n == 0) return ZG_OK;
  if (use_rbfs && proto.stype < s->n_objects.size()) {
    uint64_t nc = 0;
    bool overflow = false;
    int rc = lookup_candidates(*s, res_type, proto, &nc, &overflow, err);
    if (rc) return rc;
    if (!overflow) {
      ++lookups_rbfs;
      n = nc;
      cand = rb_cand_.as<uint32_t>();
      if (n == 0) return ZG_OK;
    } else {
      ++lookups_exhaustive;
    }
  } else {
    ++lookups_exhaustive;
  }
  const size_t count_off = (n * 4 + 7) & ~size_t(7);  // the result count lives after the ids
  if (!lk_jobs_.ensure(n * sizeof(zg_check)) || !lk_codes_.ensure(n) || !lk_ids_.ensure(count_off + 8)) {
    if (err) *err = "out of device memory (lookup)";
    return ZG_ENOMEM;
  }
  const unsigned blk = 256, grid = static_cast<unsigned>((n + blk - 1) / blk);
  lookup_fill_kernel<<<grid, blk, 0, stream>>>(cand, n, proto, lk_jobs_.as<zg_check>());
  ++launches;
  int rc = check_device(lk_jobs_.as<zg_check>(), n, lk_codes_.as<uint8_t>(), stream, true, nullptr, err);
  if (rc) return rc;
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(lk_ids_.as<uint8_t>() + count_off);
  ZG_CUDA(cudaMemsetAsync(d_count, 0, 8, stream));
  lookup_compact_kernel<<<grid, blk, 0, stream>>>(cand, lk_codes_.as<uint8_t>(), n, lk_ids_.as<uint32_t>(), n, d_count,
                                                  reinterpret_cast<uint32_t*>(ctrl_.as<unsigned long long>() + 3));
  ++launches;
  unsigned long long cnt = 0;
  uint32_t lflags = 0;
  ZG_CUDA(cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaMemcpyAsync(&lflags, ctrl_.as<unsigned long long>() + 3, 4, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  if (lflags & 1u) {
    if (err) *err = "expansion stack overflow (per-warp spill area exhausted)";
    return ZG_ENOMEM;
  }
  if (lflags & 8u) {
    // a candidate's check ended in an error (dispatch depth or work budget): the embedded SpiceDB fails
    // the LookupResources call in that case; a silently shortened answer would read as "not allowed"
    if (err) *err = "LookupResources: a candidate could not be decided (max dispatch depth / work budget exceeded)";
    return ZG_EDEPTH;
  }
  ids->resize(cnt);
  if (cnt) {
    // ascending ids: radix-sort on the device when the answer is large (a host sort of 50 k ids
    // costs milliseconds), into the job buffer that is free again by now
    const uint32_t* src = lk_ids_.as<uint32_t>();
    if (cnt > 2048) {
      std::string serr = sort_u32(lk_ids_.as<uint32_t>(), lk_jobs_.as<uint32_t>(), cnt, stream);
      if (!serr.empty()) {
        if (err) *err = serr;
        return ZG_ECUDA;
      }
      src = lk_jobs_.as<uint32_t>();
    }
    ZG_CUDA(cudaMemcpyAsync(ids->data(), src, cnt * 4, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaStreamSynchronize(stream));
    if (cnt <= 2048) std::sort(ids->begin(), ids->end());
  }
  return ZG_OK;
}


int Device::lookup_batch(const std::vector<LookupReq>& reqs, std::vector<std::vector<uint32_t>>* ids, std::vector<int>* rcs,
                         std::string* err) {
  const size_t K = reqs.size();
  ids->assign(K, {});
  rcs->assign(K, ZG_OK);
  std::shared_ptr<Snapshot> s = snap;
  if (!s) {
    if (err) *err = "no snapshot published";
    return ZG_ENOSNAPSHOT;
  }
  auto one_by_one = [&]() -> int {
    for (size_t i = 0; i < K; ++i) {
      std::string e1;
      (*rcs)[i] = lookup(reqs[i].res_type, reqs[i].proto, &(*ids)[i], &e1);
      if ((*rcs)[i] && (*rcs)[i] != ZG_EDEPTH) {
        if (err) *err = e1;
        return (*rcs)[i];
      }
    }
    return ZG_OK;
  };
  // the batch buffers overflowed: halves hold about half as much (down to the single path, which degrades to
  // the exhaustive scan by itself)
  auto halves = [&]() -> int {
    const size_t h = K / 2;
    std::vector<LookupReq> a(reqs.begin(), reqs.begin() + h), b(reqs.begin() + h, reqs.end());
    std::vector<std::vector<uint32_t>> ia, ib;
    std::vector<int> ra, rb;
    int rc = lookup_batch(a, &ia, &ra, err);
    if (rc) return rc;
    rc = lookup_batch(b, &ib, &rb, err);
    if (rc) return rc;
    for (size_t i = 0; i < h; ++i) (*ids)[i] = std::move(ia[i]), (*rcs)[i] = ra[i];
    for (size_t i = h; i < K; ++i) (*ids)[i] = std::move(ib[i - h]), (*rcs)[i] = rb[i - h];
    return ZG_OK;
  };
  if (K == 1 || !use_rbfs) return one_by_one();
  if (K > kMaxLookupBatch) return halves();
  const uint64_t cap = lb_cap_;
  for (const auto& r : reqs)
    if (r.res_type >= s->resources.size() || r.proto.stype >= s->n_objects.size()) return one_by_one();
  ZG_CUDA(cudaSetDevice(device));
  if (have_last_) ZG_CUDA(cudaStreamWaitEvent(stream, last_done_, 0));
  const unsigned long long words = (s->total_bits + 31) / 32 + 4;
  constexpr int kLevelsPerRound = 10;
  const size_t n_ctrl = 8 + ZG_MAX_DEPTH + 4 + kMaxLookupBatch;  // flags, cand count, key count, err mask | level counts | per lookup
  if (!rb_visited_.ensure(words * 4 * K) || !rb_front_[0].ensure(cap * 8) || !rb_front_[1].ensure(cap * 8) ||
      !lk_jobs_.ensure(cap * sizeof(zg_check)) || !lb_owner_.ensure(cap) || !lk_codes_.ensure(cap) ||
      !lb_params_.ensure(K * sizeof(LookupParam)) || !lb_ctrl_.ensure(n_ctrl * 8) || !lb_keys_[0].ensure(cap * 8) ||
      !lb_keys_[1].ensure(cap * 8) || !lk_ids_.ensure(cap * 4))
    return one_by_one();  // not enough memory for the batch buffers: the single path scales down by itself
  std::vector<LookupParam> lp(K);
  std::vector<unsigned long long> seeds(K);
  for (size_t i = 0; i < K; ++i) {
    const zg_check& q = reqs[i].proto;
    lp[i] = LookupParam{q.subj, reqs[i].res_type, q.perm, q.stype, q.srel};
    seeds[i] = (static_cast<unsigned long long>(i) << 48) | (static_cast<unsigned long long>(q.stype) << 32) | q.subj;
  }
  unsigned long long* ctrl = lb_ctrl_.as<unsigned long long>();
  unsigned long long* counts = ctrl + 8;
  unsigned long long* per_lookup = counts + ZG_MAX_DEPTH + 4;
  ZG_CUDA(cudaMemsetAsync(ctrl, 0, n_ctrl * 8, stream));
  ZG_CUDA(cudaMemsetAsync(rb_visited_.p, 0, words * 4 * K, stream));
  ZG_CUDA(cudaMemcpyAsync(lb_params_.p, lp.data(), K * sizeof(LookupParam), cudaMemcpyHostToDevice, stream));
  ZG_CUDA(cudaMemcpyAsync(rb_front_[0].p, seeds.data(), K * 8, cudaMemcpyHostToDevice, stream));
  const unsigned long long kk = K;
  ZG_CUDA(cudaMemcpyAsync(counts, &kk, 8, cudaMemcpyHostToDevice, stream));
  MrbfsParams p{};
  p.rrow_ptr = s->rrow_ptr.as<uint32_t>();
  p.rcol = s->rcol.as<uint32_t>();
  p.prog = s->prog.as<uint8_t>();
  p.lk = lb_params_.as<LookupParam>();
  p.counts = counts;
  p.cap = cap;
  p.visited = rb_visited_.as<uint32_t>();
  p.words = words;
  p.type_bit_base = s->type_bit_base.as<unsigned long long>();
  p.cand = lk_jobs_.as<zg_check>();
  p.cand_owner = lb_owner_.as<uint8_t>();
  p.cand_count = ctrl + 1;
  p.cand_cap = cap;
  p.flags = reinterpret_cast<uint32_t*>(ctrl);
  unsigned long long host[4] = {0, 0, 0, 0};
  int level = 0, cur = 0;
  for (;;) {
    // a round of levels without a host round trip: a level whose input is empty costs one empty launch
    for (int r = 0; r < kLevelsPerRound && level <= ZG_MAX_DEPTH + 1; ++r, ++level) {
      p.in = rb_front_[cur].as<unsigned long long>();
      p.out = rb_front_[cur ^ 1].as<unsigned long long>();
      p.level = level;
      mrbfs_expand_kernel<<<sm_count_ * 8, 256, 0, stream>>>(p);
      ++launches;
      cur ^= 1;
    }
    unsigned long long last = 0;
    ZG_CUDA(cudaMemcpyAsync(host, ctrl, sizeof host, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaMemcpyAsync(&last, counts + level, 8, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaStreamSynchronize(stream));
    if (static_cast<uint32_t>(host[0]) & 16u) return halves();  // frontier / candidate buffers overflowed
    if (last == 0 || level > ZG_MAX_DEPTH + 1) break;  // deeper objects cannot be within the dispatch cap anyway
  }
  const uint64_t n_cand = host[1];
  lookup_batches += 1;
  lookups_batched += K;
  if (n_cand == 0) return ZG_OK;
  int rc = check_device(lk_jobs_.as<zg_check>(), n_cand, lk_codes_.as<uint8_t>(), stream, true, nullptr, err);
  if (rc) return rc;
  const unsigned blk = 256, grid = static_cast<unsigned>((n_cand + blk - 1) / blk);
  lookup_keys_kernel<<<grid, blk, 0, stream>>>(lk_jobs_.as<zg_check>(), lb_owner_.as<uint8_t>(), lk_codes_.as<uint8_t>(), n_cand,
                                               lb_keys_[0].as<unsigned long long>(), ctrl + 2, per_lookup, ctrl + 3);
  ++launches;
  std::vector<unsigned long long> cnt(K);
  uint32_t cflags = 0;
  ZG_CUDA(cudaMemcpyAsync(host, ctrl, sizeof host, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaMemcpyAsync(cnt.data(), per_lookup, K * 8, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaMemcpyAsync(&cflags, ctrl_.as<unsigned long long>() + 3, 4, cudaMemcpyDeviceToHost, stream));
  ZG_CUDA(cudaStreamSynchronize(stream));
  if (cflags & 1u) {
    if (err) *err = "expansion stack overflow (per-warp spill area exhausted)";
    return ZG_ENOMEM;
  }
  const uint64_t n_keys = host[2];
  const unsigned long long err_mask = host[3];
  if (n_keys) {
    // ascending (lookup, id): one radix sort over the 32 id bits and the few lookup bits
    int kbits = 1;
    while ((1u << kbits) < K) ++kbits;
    std::string serr = sort_u64(lb_keys_[0].as<unsigned long long>(), lb_keys_[1].as<unsigned long long>(), n_keys, 32 + kbits, stream);
    if (!serr.empty()) {
      if (err) *err = serr;
      return ZG_ECUDA;
    }
    low_words_kernel<<<static_cast<unsigned>((n_keys + blk - 1) / blk), blk, 0, stream>>>(lb_keys_[1].as<unsigned long long>(),
                                                                                      n_keys, lk_ids_.as<uint32_t>());
    ++launches;
    // one copy back into pinned memory, then split per lookup
    if (pin_lk_cap_ < n_keys * 4) {
      if (pin_lk_) cudaFreeHost(pin_lk_);
      pin_lk_cap_ = 0;
      ZG_CUDA(cudaMallocHost(&pin_lk_, n_keys * 4 + (n_keys * 4) / 4));
      pin_lk_cap_ = n_keys * 4 + (n_keys * 4) / 4;
    }
    ZG_CUDA(cudaMemcpyAsync(pin_lk_, lk_ids_.p, n_keys * 4, cudaMemcpyDeviceToHost, stream));
    ZG_CUDA(cudaStreamSynchronize(stream));
    uint64_t off = 0;
    for (size_t i = 0; i < K; ++i) {
      const uint32_t* src = static_cast<const uint32_t*>(pin_lk_) + off;
      (*ids)[i].assign(src, src + cnt[i]);
      off += cnt[i];
    }
  }
  for (size_t i = 0; i < K; ++i)
    if ((err_mask >> i) & 1ull) {
      (*rcs)[i] = ZG_EDEPTH;
      (*ids)[i].clear();
    }
  return ZG_OK;
}

}  // namespace zg
