// capi.cu -- the C ABI declared in include/zgpu.h. Thin glue: argument checks,
// string <-> id resolution, locking; the work is in store.cc and device.cu.
#include <algorithm>
#include <condition_variable>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <functional>
#include <memory>
#include <atomic>
#include <mutex>
#include <shared_mutex>
#include <new>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/zgpu.h"
#include "device.h"
#include "schema.h"
#include "store.h"

using namespace zg;

// The device lock is FIFO. std::mutex makes no fairness promise, and glibc's lets a thread that unlocks and locks again
// at once overtake every sleeper: a caller issuing writes back to back (each publishes under this lock) starved the
// leaders of the check and lookup queues for seconds in tests/cabi/batcher_stress.c. Few threads ever wait here (one
// leader per queue, writers, direct callers), so one shared condition variable is enough.
class FairMutex {
  std::mutex m_;
  std::condition_variable cv_;
  uint64_t next_ = 0, serving_ = 0;

 public:
  void lock() {
    std::unique_lock<std::mutex> lk(m_);
    const uint64_t ticket = next_++;
    cv_.wait(lk, [&] { return serving_ == ticket; });
  }
  void unlock() {
    {
      std::lock_guard<std::mutex> lk(m_);
      ++serving_;
    }
    cv_.notify_all();
  }
};

// Coalescing batcher ("group commit") for zg_check_bulk. The proxy calls the boundary
// from one goroutine per rule check (pkg/authz/check.go:77-93) and per list request
// (pkg/authz/postfilter.go:127-134): many concurrent, mostly small calls. The first
// caller to arrive becomes the leader; while its launch is in flight later callers
// queue, and the next leader answers ALL of them with one launch sequence. No timer:
// batching is driven purely by arrivals during the previous launch.
struct BatchReq {
  const zg_check* items;
  uint64_t n;
  uint8_t* out;
  int rc = ZG_OK;
  std::string err;
  bool done = false, lead = false;
  std::condition_variable cv;  // every waiter sleeps on its own: a finished group wakes exactly its members and
                               // ONE next leader, not every queued caller (1 000 of them under BASELINE config 5)
};
struct Batcher {
  std::mutex m;
  std::vector<BatchReq*> queue;
  bool leader_active = false;
  uint64_t groups = 0, waited = 0, handed_over = 0;  // protocol counters (ZGPU_BATCHER_STATS=1 prints them at destroy)
  static constexpr uint64_t kMaxItemsPerLaunch = 1ull << 24;
};

// One engine may own several GPUs of the box (zg_config.n_devices): every device holds a REPLICA of the
// snapshot (1e8 relationships = 2 GB, two orders of magnitude under one GPU's HBM) and answers its slice of
// every batch; no data-path collective (SURVEY.md 8e mode 1). Device 0 is driven inline by the calling
// thread, every further device by its own worker thread (CUDA's current device is per-thread state).
struct DeviceWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<void()> task;
  bool has = false, stop = false, busy = false;
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has || stop; });
        if (stop) return;
        std::function<void()> f = std::move(task);
        has = false;
        lk.unlock();
        f();
        lk.lock();
        busy = false;
        cv.notify_all();
      }
    });
  }
  void submit(std::function<void()> f) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return !busy; });
    task = std::move(f);
    has = busy = true;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return !busy; });
  }
  ~DeviceWorker() {
    if (th.joinable()) {
      {
        std::lock_guard<std::mutex> lk(m);
        stop = true;
      }
      cv.notify_all();
      th.join();
    }
  }
};
struct Replica {
  Device dev;
  DeviceWorker worker;
};

struct zg_engine {
  // Two locks, always taken in this order:
  //   mu     the device: one launch sequence (a coalesced group of checks, a batch of lookups) or one publish at
  //          a time. Held for milliseconds by whoever leads a group.
  //   names  the schema and the store (interning tables, relationship set): shared by everything that resolves
  //          or renders names, exclusive for writers. Never held across GPU work except by a publish, so that
  //          callers can resolve their strings and QUEUE behind a running group instead of waiting for it.
  FairMutex mu;  // the device lock: one launch sequence or publish at a time
  mutable std::shared_mutex names;
  std::vector<std::unique_ptr<Replica>> replicas;  // devices 1 .. n-1 (device 0 is `dev`)
  Batcher batcher;
  Schema schema;
  bool has_schema = false;
  Store store;
  Device dev;
  int64_t clock = 0;
  uint64_t revision = 0;
  std::atomic<bool> dirty{false};  // store changed since the last publish
  bool host_only = false;
  // Concurrent LookupResources calls (one goroutine per list request: pkg/authz/responsefilterer.go:165) are
  // coalesced like the checks: the leader answers up to 64 of them with one batched launch sequence.
  struct LookupKey {
    uint64_t revision = ~0ull;
    uint32_t subj = 0;
    uint16_t res_type = 0, perm = 0, stype = 0, srel = 0;
    int64_t clock = 0;
    bool operator==(const LookupKey& o) const {
      return revision == o.revision && subj == o.subj && res_type == o.res_type && perm == o.perm && stype == o.stype &&
             srel == o.srel && clock == o.clock;
    }
  };
  struct LookupJob {
    LookupKey key;       // revision / clock filled in by the leader
    bool want_self = false;
    std::vector<uint32_t> ids;
    bool self_member = false;
    int rc = ZG_OK;
    std::string err;
    bool done = false, lead = false;
    std::condition_variable cv;
  };
  struct LookupBatcher {
    std::mutex m;
    std::vector<LookupJob*> queue;
    bool leader_active = false;
  } lookups;
  // ZG_E2BIG protocol: recent answers are kept so that a caller's retry with a larger buffer (same arguments,
  // same snapshot revision) does not recompute them
  std::deque<std::pair<LookupKey, std::shared_ptr<const std::vector<uint32_t>>>> lookup_cache;
  static constexpr size_t kLookupCache = 128;
  HostSnapshot last_built;  // kept only for zg_debug_row / host-only engines
  bool keep_built = false;
  // Watch feed (pkg/authz/watch.go:29-31): what WriteRelationships / DeleteRelationships changed,
  // stamped with the revision that made it visible. Bounded: the oldest entries are dropped and
  // a reader that asks for them is told so (as SpiceDB does past its GC window).
  struct WatchEntry {
    uint64_t revision;
    zg_tuple t;
    uint32_t expires_at, op;
  };
  std::deque<WatchEntry> watch_log;
  uint64_t watch_floor = 0;  // every change of a revision <= floor may be gone
  static constexpr size_t kWatchCap = 1u << 20;
  void log_changes(const std::vector<zg_update>& u, const std::vector<uint8_t>& changed) {
    for (size_t i = 0; i < u.size(); ++i)
      if (changed[i]) watch_log.push_back(WatchEntry{revision, u[i].t, u[i].expires_at, u[i].op});
    while (watch_log.size() > kWatchCap) {
      watch_floor = watch_log.front().revision;
      watch_log.pop_front();
    }
    // a revision is dropped whole: a reader never sees part of a write
    while (!watch_log.empty() && watch_log.front().revision <= watch_floor) watch_log.pop_front();
  }
};

// fn(device, index) on every device of the engine, concurrently; returns when all are done.
static void on_all_devices(zg_engine* e, const std::function<void(Device&, size_t)>& fn) {
  for (size_t i = 0; i < e->replicas.size(); ++i) {
    Replica* r = e->replicas[i].get();
    r->worker.submit([r, i, &fn] { fn(r->dev, i + 1); });
  }
  fn(e->dev, 0);
  for (auto& r : e->replicas) r->worker.wait();
}

#define LOCK_DEVICE(e) std::lock_guard<FairMutex> g((e)->mu)
#define LOCK_NAMES_SHARED(e) std::shared_lock<std::shared_mutex> ng((e)->names)
#define LOCK_NAMES_UNIQUE(e) std::unique_lock<std::shared_mutex> ng((e)->names)

static thread_local std::string g_err;
static constexpr size_t kMaxLookupGroup = 64;  // = kMaxLookupBatch (kernels.cuh)
static const char* kShardedMsg =
    "sharded engine: use zg_shard_pass / zg_shard_subqueries / zg_shard_fold (dist.ShardedStoreChecker)";

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
extern "C" const char* zg_last_error(void) { return g_err.c_str(); }
extern "C" size_t zg_last_error_copy(char* buf, size_t cap) {
  const size_t n = g_err.size();
  if (buf && cap) {
    const size_t m = n < cap - 1 ? n : cap - 1;
    std::memcpy(buf, g_err.data(), m);
    buf[m] = 0;
  }
  return n;
}

static uint32_t now_of(const zg_engine* e) {
  return static_cast<uint32_t>(e->clock ? e->clock : static_cast<int64_t>(time(nullptr)));
}
static bool none_rel(const char* s) { return !s || !*s || std::strcmp(s, "...") == 0; }

extern "C" int zg_engine_create(const zg_config* cfg, zg_engine** out) {
  if (!out) return fail(ZG_EINVAL, "out is NULL");
  *out = nullptr;
  zg_engine* e = new (std::nothrow) zg_engine();
  if (!e) return fail(ZG_ENOMEM, "out of memory");
  if (cfg && (cfg->flags & ZG_FLAG_HOST_ONLY)) {
    e->host_only = true;
    e->keep_built = true;
    *out = e;
    return ZG_OK;
  }
  std::string err = e->dev.init(cfg ? cfg->device : -1, cfg ? cfg->subquery_capacity : 0, cfg ? cfg->work_budget : 0);
  if (!err.empty()) {
    delete e;
    return fail(ZG_ECUDA, err);
  }
  if (cfg && (cfg->flags & ZG_FLAG_FORWARD_ONLY)) e->dev.invert = false;
  if (cfg && cfg->shard_count > 1) {
    if (cfg->shard_rank >= cfg->shard_count) {
      delete e;
      return fail(ZG_EINVAL, "shard_rank must be < shard_count");
    }
    e->dev.shard_count = cfg->shard_count;
    e->dev.shard_rank = cfg->shard_rank;
  }
  if (cfg && cfg->n_devices > 1) {
    if (cfg->shard_count > 1) {
      delete e;
      return fail(ZG_EINVAL, "n_devices > 1 makes replicas; a sharded engine owns one device");
    }
    int visible = 0;
    cudaGetDeviceCount(&visible);
    const int first = e->dev.device;
    const int want = cfg->n_devices == ZG_ALL_DEVICES ? visible - first : static_cast<int>(cfg->n_devices);
    if (want < 1 || first + want > visible) {
      delete e;
      return fail(ZG_EINVAL, "n_devices exceeds the visible CUDA devices");
    }
    std::vector<std::string> errs(static_cast<size_t>(want));
    for (int d = 1; d < want; ++d) {
      e->replicas.emplace_back(new Replica());
      e->replicas.back()->worker.start();
    }
    const zg_config c = *cfg;
    on_all_devices(e, [&](Device& dev, size_t i) {
      if (i == 0) return;
      errs[i] = dev.init(first + static_cast<int>(i), c.subquery_capacity, c.work_budget);
      if (c.flags & ZG_FLAG_FORWARD_ONLY) dev.invert = false;
    });
    for (const auto& m : errs)
      if (!m.empty()) {
        delete e;
        return fail(ZG_ECUDA, m);
      }
  }
  *out = e;
  return ZG_OK;
}
extern "C" void zg_engine_destroy(zg_engine* e) {
  if (e && std::getenv("ZGPU_BATCHER_STATS"))
    std::fprintf(stderr, "zgpu batcher: %llu groups, %llu callers waited, %llu hand-overs\n",
                 static_cast<unsigned long long>(e->batcher.groups), static_cast<unsigned long long>(e->batcher.waited),
                 static_cast<unsigned long long>(e->batcher.handed_over));
  delete e;
}

extern "C" int zg_load_schema(zg_engine* e, const char* dsl, size_t len) {
  if (!e || !dsl) return fail(ZG_EINVAL, "NULL argument");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  Schema s;
  std::string err = s.parse(std::string(dsl, len));
  if (!err.empty()) return fail(ZG_EINVAL, err);
  e->schema = std::move(s);
  e->has_schema = true;
  e->store.reset(&e->schema);
  e->store.shard_count = e->dev.shard_count;
  e->store.shard_rank = e->dev.shard_rank;
  e->dev.snap.reset();
  for (auto& r : e->replicas) r->dev.snap.reset();
  e->dirty = true;
  return ZG_OK;
}

#define NEED_SCHEMA(e, ret)              \
  if (!(e) || !(e)->has_schema) {        \
    g_err = "no schema loaded";          \
    return ret;                          \
  }

extern "C" int zg_num_types(const zg_engine* e) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  return static_cast<int>(e->schema.types.size());
}
extern "C" int zg_num_slots(const zg_engine* e) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  return static_cast<int>(e->schema.slots.size());
}
extern "C" int zg_type_id(const zg_engine* e, const char* n) {
  NEED_SCHEMA(e, -1);
  return n ? e->schema.type_id(n) : -1;
}
extern "C" int zg_slot_id(const zg_engine* e, int t, const char* n) {
  NEED_SCHEMA(e, -1);
  return n ? e->schema.slot_id(t, n) : -1;
}
extern "C" int zg_slot_type(const zg_engine* e, int s) {
  NEED_SCHEMA(e, -1);
  return s >= 0 && s < static_cast<int>(e->schema.slots.size()) ? e->schema.slots[s].type : -1;
}
extern "C" int zg_slot_is_permission(const zg_engine* e, int s) {
  NEED_SCHEMA(e, -1);
  return s >= 0 && s < static_cast<int>(e->schema.slots.size()) ? e->schema.slots[s].is_perm : -1;
}
extern "C" const char* zg_slot_name(const zg_engine* e, int s) {
  NEED_SCHEMA(e, nullptr);
  return s >= 0 && s < static_cast<int>(e->schema.slots.size()) ? e->schema.slots[s].name.c_str() : nullptr;
}
extern "C" const char* zg_type_name(const zg_engine* e, int t) {
  NEED_SCHEMA(e, nullptr);
  return t >= 0 && t < static_cast<int>(e->schema.types.size()) ? e->schema.types[t].name.c_str() : nullptr;
}

extern "C" uint32_t zg_intern_object(zg_engine* e, int t, const char* id) {
  NEED_SCHEMA(e, ZG_NO_OBJECT);
  if (!id || !*id || t < 0 || t >= static_cast<int>(e->schema.types.size())) return ZG_NO_OBJECT;
  LOCK_NAMES_UNIQUE(e);
  return e->store.intern(t, id);
}
extern "C" uint32_t zg_find_object(const zg_engine* e, int t, const char* id) {
  NEED_SCHEMA(e, ZG_NO_OBJECT);
  if (!id) return ZG_NO_OBJECT;
  LOCK_NAMES_SHARED(e);
  return e->store.find(t, id);
}
extern "C" int zg_object_name(const zg_engine* e, int t, uint32_t id, char* buf, size_t cap) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  LOCK_NAMES_SHARED(e);
  std::string_view n;
  if (!e->store.name(t, id, &n)) return fail(ZG_EINVAL, "object has no name");
  if (n.size() + 1 > cap || !buf) return ZG_E2BIG;
  std::memcpy(buf, n.data(), n.size());
  buf[n.size()] = 0;
  return static_cast<int>(n.size());
}

extern "C" int zg_load_tuples(zg_engine* e, const zg_tuple* t, const uint32_t* expires, uint64_t n) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!t && n) return fail(ZG_EINVAL, "NULL tuples");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  std::string err = e->store.load(t, expires, n);
  if (!err.empty()) return fail(ZG_EINVAL, err);
  e->dirty = true;
  return ZG_OK;
}
extern "C" int zg_apply_updates(zg_engine* e, const zg_update* u, uint64_t n) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!u && n) return fail(ZG_EINVAL, "NULL updates");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  int code = ZG_OK;
  std::string err = e->store.apply(u, n, &code);
  if (!err.empty()) return fail(code, err);
  e->dirty = true;
  return ZG_OK;
}

// Caller holds e->mu AND e->names exclusively.
static int publish_locked(zg_engine* e) {
  (void)e->store.layout();  // settles the object capacities once: the replicas' publishes below only read the store
  if (!e->host_only) {
    // default: build the CSR on the GPU (csrc/build.cu). ZGPU_HOST_BUILD=1 builds on the host
    // and uploads; ZGPU_VERIFY_BUILD=1 does both and compares every array (tests).
    const char* hb = std::getenv("ZGPU_HOST_BUILD");
    const char* vb = std::getenv("ZGPU_VERIFY_BUILD");
    if (!(hb && *hb && *hb != '0')) {
      const bool verify = vb && *vb && *vb != '0';
      // a few updates against a resident snapshot: merge them in (build.cu gpu_apply_delta); anything the
      // journal cannot express (bulk load, new layout, no snapshot yet) rebuilds
      const char* nd = std::getenv("ZGPU_NO_DELTA");
      const bool no_delta = nd && *nd && *nd != '0';
      std::vector<std::string> errs(1 + e->replicas.size());
      on_all_devices(e, [&](Device& dev, size_t i) {  // every replica merges / builds for itself, concurrently
        std::string err = no_delta ? std::string("full") : dev.publish_delta(e->store, e->schema, e->revision + 1);
        if (err == "full") err = dev.publish_gpu(e->store, e->schema, e->revision + 1, verify);
        else if (err.empty() && verify) err = dev.verify_against_host(e->store, e->schema);
        errs[i] = err;
      });
      for (const auto& err : errs)
        if (!err.empty()) return fail(ZG_ECUDA, err);
      ++e->revision;
      e->store.journal_clear();
      e->last_built = HostSnapshot();
      e->dirty = false;
      return ZG_OK;
    }
  }
  HostSnapshot h = e->store.build();
  if (!h.err.empty()) return fail(ZG_EINVAL, h.err);
  if (e->host_only) {
    ++e->revision;
    e->last_built = std::move(h);
    e->dirty = false;
    return ZG_OK;
  }
  std::string err = e->dev.publish(h, e->schema, ++e->revision);
  if (!err.empty()) return fail(ZG_ECUDA, err);
  if (e->keep_built) e->last_built = std::move(h);
  e->dirty = false;
  return ZG_OK;
}
extern "C" int zg_publish(zg_engine* e) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  return publish_locked(e);
}
extern "C" int zg_clear_relationships(zg_engine* e) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  e->store.clear_relationships();
  e->dirty = true;
  // the feed cannot describe "everything went away" change by change: readers restart from here
  e->watch_log.clear();
  int rc = publish_locked(e);
  e->watch_floor = e->revision;
  return rc;
}
extern "C" uint64_t zg_num_tuples(const zg_engine* e) {
  if (!e || !e->has_schema) return 0;
  LOCK_NAMES_UNIQUE(e);  // match() folds the duplicates of bulk loads into the index
  std::vector<uint64_t> idx;
  Store::Filter f;
  e->store.match(f, 0, &idx);  // also folds duplicates of bulk loads
  return e->store.size();
}
extern "C" void zg_set_clock(zg_engine* e, int64_t t) {
  if (e) e->clock = t;
}

// ---- string resolution ---------------------------------------------------------------

static std::string rel_text(const zg_rel_str& r) {
  auto s = [](const char* x) { return std::string(x ? x : ""); };
  return s(r.res_type) + ":" + s(r.res_id) + "#" + s(r.relation) + "@" + s(r.subj_type) + ":" + s(r.subj_id) +
         (none_rel(r.subj_rel) ? "" : "#" + s(r.subj_rel));
}

// SpiceDB object-id validation (v1 API): 1..1024 characters of [a-zA-Z0-9/_|\-=+], or "*".
static bool valid_object_id(const char* id) {
  size_t n = std::strlen(id);
  if (n == 0 || n > 1024) return false;
  if (n == 1 && id[0] == '*') return true;
  for (size_t i = 0; i < n; ++i) {
    const unsigned char c = static_cast<unsigned char>(id[i]);
    if (!(std::isalnum(c) || c == '/' || c == '_' || c == '|' || c == '-' || c == '=' || c == '+')) return false;
  }
  return true;
}

// Resolve a relationship to WRITE (interns objects). Returns "" or an error.
static std::string resolve_write(zg_engine* e, const zg_rel_str& r, zg_tuple* t) {
  const Schema& sc = e->schema;
  if (!r.res_type || !r.res_id || !r.relation || !r.subj_type || !r.subj_id) return "relationship has NULL fields";
  int rt = sc.type_id(r.res_type), st = sc.type_id(r.subj_type);
  if (rt < 0) return std::string("object definition `") + r.res_type + "` not found";
  if (st < 0) return std::string("object definition `") + r.subj_type + "` not found";
  int rel = sc.slot_id(rt, r.relation);
  if (rel < 0 || sc.slots[rel].is_perm)
    return std::string("relation `") + r.relation + "` not found under definition `" + r.res_type + "`";
  if (!valid_object_id(r.res_id) || std::strcmp(r.res_id, "*") == 0)
    return std::string("invalid resource object id `") + r.res_id + "`";
  if (!valid_object_id(r.subj_id)) return std::string("invalid subject object id `") + r.subj_id + "`";
  t->rel = static_cast<uint16_t>(rel);
  t->stype = static_cast<uint16_t>(st);
  t->flags = 0;
  t->srel = kNone;
  if (std::strcmp(r.subj_id, "*") == 0) {
    if (!none_rel(r.subj_rel)) return "wildcard subjects cannot have a relation";
    t->srel = kWildcard;
    t->subj = 0;
  } else {
    if (!none_rel(r.subj_rel)) {
      int sr = sc.slot_id(st, r.subj_rel);
      if (sr < 0) return std::string("relation `") + r.subj_rel + "` not found under definition `" + r.subj_type + "`";
      t->srel = static_cast<uint16_t>(sr);
    }
    t->subj = e->store.intern(st, r.subj_id);
  }
  t->res = e->store.intern(rt, r.res_id);
  return "";
}

static Store::Filter resolve_filter(const zg_engine* e, const zg_filter_str& f, std::string* err) {
  const Schema& sc = e->schema;
  Store::Filter o;
  auto set = [](const char* s) { return s && *s; };
  if (set(f.res_type)) {
    o.res_type = sc.type_id(f.res_type);
    if (o.res_type < 0) {
      *err = std::string("object definition `") + f.res_type + "` not found";
      return o;
    }
  }
  if (set(f.res_id)) {
    if (o.res_type < 0) {
      *err = "resource id filter needs a resource type";
      return o;
    }
    o.has_res = true;
    o.res = e->store.find(o.res_type, f.res_id);
    if (o.res == ZG_NO_OBJECT) o.impossible = true;
  }
  if (set(f.relation)) {
    if (o.res_type < 0) {
      *err = "relation filter needs a resource type";
      return o;
    }
    o.rel = sc.slot_id(o.res_type, f.relation);
    if (o.rel < 0) {
      *err = std::string("relation `") + f.relation + "` not found under definition `" + f.res_type + "`";
      return o;
    }
  }
  if (set(f.subj_type)) {
    o.subj_type = sc.type_id(f.subj_type);
    if (o.subj_type < 0) {
      *err = std::string("object definition `") + f.subj_type + "` not found";
      return o;
    }
  }
  if (set(f.subj_id)) {
    if (o.subj_type < 0) {
      *err = "subject id filter needs a subject type";
      return o;
    }
    if (std::strcmp(f.subj_id, "*") == 0) {
      o.subj_wildcard = true;
    } else {
      o.has_subj = true;
      o.subj = e->store.find(o.subj_type, f.subj_id);
      if (o.subj == ZG_NO_OBJECT) o.impossible = true;
    }
  }
  if (set(f.subj_rel)) {
    if (o.subj_type < 0) {
      *err = "subject relation filter needs a subject type";
      return o;
    }
    o.has_srel = true;
    if (std::strcmp(f.subj_rel, "...") == 0) {
      o.srel = kNone;
    } else {
      int sr = sc.slot_id(o.subj_type, f.subj_rel);
      if (sr < 0) o.impossible = true;
      else o.srel = static_cast<uint16_t>(sr);
    }
  }
  return o;
}

static int check_preconditions(zg_engine* e, const zg_precondition_str* pre, uint64_t n_pre) {
  std::vector<uint64_t> idx;
  for (uint64_t i = 0; i < n_pre; ++i) {
    std::string err;
    Store::Filter f = resolve_filter(e, pre[i].filter, &err);
    if (!err.empty()) return fail(ZG_EINVAL, "precondition " + std::to_string(i) + ": " + err);
    e->store.match(f, now_of(e), &idx);
    const bool any = !idx.empty();
    if (pre[i].op == ZG_PRECOND_MUST_MATCH && !any)
      return fail(ZG_EPRECOND, "precondition " + std::to_string(i) + " (MUST_MATCH) failed");
    if (pre[i].op == ZG_PRECOND_MUST_NOT_MATCH && any)
      return fail(ZG_EPRECOND, "precondition " + std::to_string(i) + " (MUST_NOT_MATCH) failed");
    if (pre[i].op != ZG_PRECOND_MUST_MATCH && pre[i].op != ZG_PRECOND_MUST_NOT_MATCH)
      return fail(ZG_EINVAL, "unknown precondition operation");
  }
  return ZG_OK;
}

extern "C" int zg_write_relationships(zg_engine* e, const zg_update_str* ups, uint64_t n, const zg_precondition_str* pre,
                                      uint64_t n_pre) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!ups && n) || (!pre && n_pre)) return fail(ZG_EINVAL, "NULL argument");
  if (n > 1000)  // pkg/spicedb/spicedb.go:34 WithMaximumUpdatesPerWrite(1000)
    return fail(ZG_EINVAL, "update count of " + std::to_string(n) + " is greater than maximum allowed of 1000");
  if (n_pre > 1000)  // pkg/spicedb/spicedb.go:35
    return fail(ZG_EINVAL, "precondition count is greater than maximum allowed of 1000");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  std::vector<zg_update> u(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (ups[i].op == ZG_OP_DELETE) {
      // deleting a relationship whose objects were never written is a no-op, not an error
      const zg_rel_str& r = ups[i].rel;
      const Schema& sc = e->schema;
      int rt = r.res_type ? sc.type_id(r.res_type) : -1, st = r.subj_type ? sc.type_id(r.subj_type) : -1;
      int rel = rt >= 0 && r.relation ? sc.slot_id(rt, r.relation) : -1;
      if (rt < 0 || st < 0 || rel < 0 || sc.slots[rel].is_perm || !r.res_id || !r.subj_id)
        return fail(ZG_EINVAL, "update " + std::to_string(i) + ": malformed relationship " + rel_text(r));
    }
    std::string err = resolve_write(e, ups[i].rel, &u[i].t);
    if (!err.empty()) return fail(ZG_EINVAL, "update " + std::to_string(i) + " (" + rel_text(ups[i].rel) + "): " + err);
    u[i].expires_at = ups[i].expires_at;
    u[i].op = ups[i].op;
    // two updates of one relationship in a single write are rejected by SpiceDB
    for (uint64_t j = 0; j < i; ++j)
      if (key_of(u[j].t) == key_of(u[i].t))
        return fail(ZG_EINVAL, "found more than one update with relationship " + rel_text(ups[i].rel));
  }
  int rc = check_preconditions(e, pre, n_pre);
  if (rc) return rc;
  int code = ZG_OK;
  std::vector<uint8_t> changed;
  std::string err = e->store.apply(u.data(), n, &code, &changed);
  if (!err.empty()) return fail(code, err);
  rc = publish_locked(e);
  if (rc == ZG_OK) e->log_changes(u, changed);
  return rc;
}

extern "C" int zg_delete_relationships(zg_engine* e, const zg_filter_str* filter, const zg_precondition_str* pre,
                                       uint64_t n_pre, uint64_t* n_deleted) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!filter || (!pre && n_pre)) return fail(ZG_EINVAL, "NULL argument");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  std::string err;
  Store::Filter f = resolve_filter(e, *filter, &err);
  if (!err.empty()) return fail(ZG_EINVAL, err);
  int rc = check_preconditions(e, pre, n_pre);
  if (rc) return rc;
  std::vector<uint64_t> idx;
  e->store.match(f, 0, &idx);  // expired relationships are deleted too
  std::vector<zg_update> u(idx.size());
  for (size_t i = 0; i < idx.size(); ++i) {
    u[i].t = e->store.tuples[idx[i]];
    u[i].expires_at = 0;
    u[i].op = ZG_OP_DELETE;
  }
  int code = ZG_OK;
  std::vector<uint8_t> changed;
  err = e->store.apply(u.data(), u.size(), &code, &changed);
  if (!err.empty()) return fail(code, err);
  if (n_deleted) *n_deleted = idx.size();
  rc = publish_locked(e);
  if (rc == ZG_OK) e->log_changes(u, changed);
  return rc;
}

static std::string tuple_text(const zg_engine* e, const zg_tuple& t) {
  const Schema& sc = e->schema;
  const SlotInfo& rel = sc.slots[t.rel];
  auto obj = [&](int type, uint32_t id) {
    std::string_view n;
    return e->store.name(type, id, &n) ? std::string(n) : std::to_string(id);
  };
  std::string s = sc.types[rel.type].name + ":" + obj(rel.type, t.res) + "#" + rel.name + "@" + sc.types[t.stype].name + ":";
  if (t.srel == kWildcard) return s + "*";
  s += obj(t.stype, t.subj);
  if (t.srel != kNone) s += "#" + sc.slots[t.srel].name;
  return s;
}

extern "C" int zg_read_relationships(zg_engine* e, const zg_filter_str* filter, char* buf, size_t cap, size_t* need,
                                     uint64_t* n_out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!filter) return fail(ZG_EINVAL, "NULL filter");
  LOCK_NAMES_UNIQUE(e);  // match() may (re)build the relationship index
  std::string err;
  Store::Filter f = resolve_filter(e, *filter, &err);
  if (!err.empty()) return fail(ZG_EINVAL, err);
  std::vector<uint64_t> idx;
  e->store.match(f, now_of(e), &idx);
  std::vector<std::string> lines;
  lines.reserve(idx.size());
  for (uint64_t i : idx) lines.push_back(tuple_text(e, e->store.tuples[i]));
  std::sort(lines.begin(), lines.end());
  size_t total = 1;
  for (const auto& l : lines) total += l.size() + 1;
  if (need) *need = total;
  if (n_out) *n_out = lines.size();
  if (total > cap || !buf) return ZG_E2BIG;
  size_t w = 0;
  for (const auto& l : lines) {
    std::memcpy(buf + w, l.data(), l.size());
    w += l.size();
    buf[w++] = '\n';
  }
  buf[w] = 0;
  return ZG_OK;
}

extern "C" int zg_watch_read(zg_engine* e, uint64_t since_revision, const char* res_type, char* buf, size_t cap,
                             size_t* need, uint64_t* n_out, uint64_t* through_revision) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  LOCK_DEVICE(e);  // the change log is appended under the device lock
  LOCK_NAMES_SHARED(e);
  int want = -1;
  if (res_type && *res_type) {
    want = e->schema.type_id(res_type);
    if (want < 0) return fail(ZG_EINVAL, std::string("object definition `") + res_type + "` not found");
  }
  if (since_revision < e->watch_floor)
    return fail(ZG_EPRECOND, "watch: revision " + std::to_string(since_revision) + " is older than the retained feed (" +
                                 std::to_string(e->watch_floor) + ")");
  static const char* kOp[] = {"TOUCH", "CREATE", "DELETE"};
  std::string out;
  uint64_t n = 0;
  // entries are in revision order: binary search for the first one after `since`
  auto it = std::partition_point(e->watch_log.begin(), e->watch_log.end(),
                                 [&](const zg_engine::WatchEntry& w) { return w.revision <= since_revision; });
  for (; it != e->watch_log.end(); ++it) {
    if (want >= 0 && e->schema.slots[it->t.rel].type != want) continue;
    out += std::to_string(it->revision);
    out += ' ';
    out += kOp[it->op];
    out += ' ';
    out += tuple_text(e, it->t);
    if (it->expires_at && it->op != ZG_OP_DELETE) out += " " + std::to_string(it->expires_at);
    out += '\n';
    ++n;
  }
  if (need) *need = out.size() + 1;
  if (n_out) *n_out = n;
  if (through_revision) *through_revision = e->revision;
  if (out.size() + 1 > cap || !buf) return ZG_E2BIG;
  std::memcpy(buf, out.data(), out.size());
  buf[out.size()] = 0;
  return ZG_OK;
}

// ---- hot path ------------------------------------------------------------------------

// Makes pending writes visible before a string entry point resolves against them. Takes both locks only when
// there is something to publish; call it with NO lock held.
static int publish_if_needed(zg_engine* e) {
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  if (e->dev.shard_count > 1) return fail(ZG_EINVAL, kShardedMsg);
  if (!e->dirty.load() && e->dev.snap) return ZG_OK;
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  if (e->dirty.load() || !e->dev.snap) return publish_locked(e);
  return ZG_OK;
}

// Runs one group of queued requests; the caller holds the device lock.
static void run_group(zg_engine* e, std::vector<BatchReq*>& group) {
  int rc = ZG_OK;
  std::string err;
  if (e->dev.shard_count > 1) {
    rc = ZG_EINVAL;
    err = kShardedMsg;
  } else if (e->host_only) {
    rc = ZG_ECUDA;
    err = "host-only engine: no CUDA device, and libzgpu has no CPU fallback";
  } else if (!e->dev.snap) {
    rc = ZG_ENOSNAPSHOT;
    err = "no snapshot published (call zg_publish)";
  } else {
    std::vector<Device::HostReq> reqs;
    reqs.reserve(group.size());
    uint64_t total = 0;
    for (BatchReq* r : group) {
      reqs.push_back({r->items, r->n, r->out});
      total += r->n;
    }
    const size_t nd = 1 + e->replicas.size();
    if (nd == 1 || total < nd * 4096) {
      e->dev.now = now_of(e);
      rc = e->dev.check_host_multi(reqs, &err);
    } else {
      // replicas: every device answers a contiguous slice of the group's items
      std::vector<std::vector<Device::HostReq>> part(nd);
      const uint64_t per = (total + nd - 1) / nd;
      size_t d = 0;
      uint64_t room = per;
      for (const auto& r : reqs) {
        uint64_t off = 0;
        while (off < r.n) {
          const uint64_t take = std::min(room, r.n - off);
          part[d].push_back({r.items + off, take, r.out + off});
          off += take;
          room -= take;
          if (room == 0 && d + 1 < nd) {
            ++d;
            room = per;
          }
        }
      }
      std::vector<int> rcs(nd, ZG_OK);
      std::vector<std::string> errs(nd);
      const uint32_t now = now_of(e);
      on_all_devices(e, [&](Device& dev, size_t i) {
        dev.now = now;
        if (!part[i].empty()) rcs[i] = dev.check_host_multi(part[i], &errs[i]);
      });
      for (size_t i = 0; i < nd && !rc; ++i)
        if (rcs[i]) {
          rc = rcs[i];
          err = errs[i];
        }
    }
  }
  for (BatchReq* r : group) {
    r->rc = rc;
    r->err = err;
  }
}

extern "C" int zg_check_bulk(zg_engine* e, const zg_check* items, uint64_t n, uint8_t* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!items || !out) && n) return fail(ZG_EINVAL, "NULL argument");
  if (n == 0) return ZG_OK;
  Batcher& b = e->batcher;
  BatchReq me{items, n, out};
  std::unique_lock<std::mutex> lk(b.m);
  b.queue.push_back(&me);
  if (b.leader_active) {
    ++b.waited;
    me.cv.wait(lk, [&] { return me.done || me.lead; });
  }
  if (!me.done) {
    // leader (nobody was leading, or the previous leader handed over to the head of the queue -- me): take as many
    // queued requests as fit one launch, mine first, answer them, wake their callers, hand leadership on
    b.leader_active = true;
    std::vector<BatchReq*> group;
    lk.unlock();
    {
      // the device first, the group second: whatever queued while this leader waited for the GPU (behind a batch of
      // lookups, a publish, the previous group's tail) rides along
      std::lock_guard<FairMutex> dev_lock(e->mu);
      lk.lock();
      uint64_t total = 0;
      size_t take = 0;
      while (take < b.queue.size() && (group.empty() || total + b.queue[take]->n <= Batcher::kMaxItemsPerLaunch)) {
        total += b.queue[take]->n;
        group.push_back(b.queue[take++]);
      }
      b.queue.erase(b.queue.begin(), b.queue.begin() + static_cast<long>(take));
      lk.unlock();
      run_group(e, group);
    }
    lk.lock();
    ++b.groups;
    for (BatchReq* r : group) {
      r->done = true;
      if (r != &me) r->cv.notify_one();
    }
    if (!b.queue.empty()) {
      ++b.handed_over;
      b.queue.front()->lead = true;
      b.queue.front()->cv.notify_one();
    } else {
      b.leader_active = false;
    }
  }
  lk.unlock();
  return me.rc ? fail(me.rc, me.err) : ZG_OK;
}

extern "C" int zg_check_bulk_device(zg_engine* e, const zg_check* d_items, uint64_t n, uint8_t* d_out, void* stream) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!d_items || !d_out) && n) return fail(ZG_EINVAL, "NULL argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  if (e->dev.shard_count > 1) return fail(ZG_EINVAL, kShardedMsg);
  e->dev.now = now_of(e);
  std::string err;
  int rc = e->dev.check_device(d_items, n, d_out, static_cast<cudaStream_t>(stream), true, nullptr, &err);
  return rc ? fail(rc, err) : ZG_OK;
}

extern "C" int zg_count_alg_bytes(zg_engine* e, const zg_check* items, uint64_t n, uint64_t* bytes) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!items || !bytes) return fail(ZG_EINVAL, "NULL argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  e->dev.now = now_of(e);
  std::string err;
  void *d_in = nullptr, *d_out = nullptr;
  if (cudaMalloc(&d_in, n * sizeof(zg_check)) != cudaSuccess || cudaMalloc(&d_out, n ? n : 1) != cudaSuccess) {
    if (d_in) cudaFree(d_in);
    return fail(ZG_ENOMEM, "out of device memory");
  }
  cudaMemcpy(d_in, items, n * sizeof(zg_check), cudaMemcpyHostToDevice);
  uint64_t b = 0;
  int rc = e->dev.check_device(static_cast<zg_check*>(d_in), n, static_cast<uint8_t*>(d_out), e->dev.stream, true, &b, &err);
  cudaFree(d_in);
  cudaFree(d_out);
  if (rc) return fail(rc, err);
  *bytes = b;
  return ZG_OK;
}

// Resolve a CHECK item without interning anything (query strings must not grow the store).
// The items of one bulk call mostly repeat their literal fields -- a post-filter sends one template
// per list item, so type, permission and subject are constant and only the resource id varies
// (pkg/authz/postfilter.go:88-110; "low-hanging fruit" in pkg/rules/rules.go:1015-1017) -- so the
// schema lookups and the subject lookup of the previous item are reused when the strings repeat.
struct ResolveMemo {
  struct SchemaPart {  // (res_type, relation, subj_type, subj_rel) -> slots
    std::string res_type, relation, subj_type, subj_rel;
    int rt = -1, st = -1, perm = -1;
    uint16_t srel = kNone;
    bool ok = false, used = false;
  };
  static constexpr int kWays = 4;  // a bulk request interleaves a handful of templates at most
  SchemaPart parts[kWays];
  int last = 0, next = 0;
  std::string subj_id;  // subject of the previous item
  int subj_st = -1;
  uint32_t su = ZG_NO_OBJECT;
};

static inline bool same(const std::string& a, const char* b) { return std::strcmp(a.c_str(), b) == 0; }

static const ResolveMemo::SchemaPart& resolve_schema_part(const Schema& sc, const zg_rel_str& r, const char* srel_s,
                                                          ResolveMemo* m) {
  for (int k = 0; k < ResolveMemo::kWays; ++k) {
    const int w = (m->last + k) % ResolveMemo::kWays;
    const ResolveMemo::SchemaPart& p = m->parts[w];
    if (p.used && same(p.res_type, r.res_type) && same(p.relation, r.relation) && same(p.subj_type, r.subj_type) &&
        same(p.subj_rel, srel_s)) {
      m->last = w;
      return p;
    }
  }
  const int w = m->next;
  m->next = (m->next + 1) % ResolveMemo::kWays;
  ResolveMemo::SchemaPart& p = m->parts[w];
  p.res_type = r.res_type;
  p.relation = r.relation;
  p.subj_type = r.subj_type;
  p.subj_rel = srel_s;
  p.used = true;
  p.ok = false;
  p.srel = kNone;
  p.rt = sc.type_id(r.res_type);
  p.st = sc.type_id(r.subj_type);
  if (p.rt >= 0 && p.st >= 0) {
    p.perm = sc.slot_id(p.rt, r.relation);
    const int sr = *srel_s ? sc.slot_id(p.st, srel_s) : 0;
    if (p.perm >= 0 && sr >= 0) {
      if (*srel_s) p.srel = static_cast<uint16_t>(sr);
      p.ok = true;
    }
  }
  m->last = w;
  return p;
}

static void resolve_check(const zg_engine* e, const zg_rel_str& r, zg_check* c, ResolveMemo* m) {
  c->res = ZG_NO_OBJECT;
  c->subj = ZG_NO_OBJECT - 1;
  c->perm = kNone;  // invalid -> ZG_ITEM_ERROR
  c->stype = 0;
  c->srel = kNone;
  c->flags = 0;
  if (!r.res_type || !r.res_id || !r.relation || !r.subj_type || !r.subj_id) return;
  const ResolveMemo::SchemaPart& p = resolve_schema_part(e->schema, r, none_rel(r.subj_rel) ? "" : r.subj_rel, m);
  if (!p.ok) return;
  const int rt = p.rt, st = p.st;
  c->srel = p.srel;
  c->stype = static_cast<uint16_t>(st);
  c->res = e->store.find(rt, r.res_id, std::strlen(r.res_id));
  if (!(m->subj_st == st && same(m->subj_id, r.subj_id))) {
    m->subj_id = r.subj_id;
    m->su = e->store.find(st, r.subj_id, m->subj_id.size());
    m->subj_st = st;
  }
  uint32_t su = m->su;
  // two never-written names that are the same object must still compare equal
  if (su == ZG_NO_OBJECT) su = (c->res == ZG_NO_OBJECT && rt == st && std::strcmp(r.res_id, r.subj_id) == 0)
                                   ? ZG_NO_OBJECT : ZG_NO_OBJECT - 1;
  c->subj = su;
  c->perm = static_cast<uint16_t>(p.perm);
}

static void resolve_checks_locked(const zg_engine* e, const zg_rel_str* items, uint64_t n, zg_check* out) {
  auto run = [&](uint64_t b, uint64_t en) {
    ResolveMemo memo;
    for (uint64_t i = b; i < en; ++i) resolve_check(e, items[i], &out[i], &memo);
  };
  // Resolution only reads the schema and the interning tables (the engine lock is held), so a large
  // batch is split over host threads; below ~16k items one thread finishes before others would start.
  constexpr uint64_t kPerThread = 8192;
  unsigned hw = std::thread::hardware_concurrency();
  uint64_t nt = std::min<uint64_t>({n / kPerThread, hw ? hw : 1u, 16u});
  if (nt < 2) return run(0, n);
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  const uint64_t per = (n + nt - 1) / nt;
  for (uint64_t t = 1; t < nt; ++t) th.emplace_back(run, std::min(n, t * per), std::min(n, (t + 1) * per));
  run(0, std::min(n, per));
  for (auto& x : th) x.join();
}

extern "C" int zg_resolve_checks(zg_engine* e, const zg_rel_str* items, uint64_t n, zg_check* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!items || !out) && n) return fail(ZG_EINVAL, "NULL argument");
  LOCK_NAMES_SHARED(e);
  resolve_checks_locked(e, items, n, out);
  return ZG_OK;
}

static int resolve_packed_locked(zg_engine* e, const char* res_type, const char* relation, const char* subj_type,
                                 const char* subj_rel, const char* res_ids, const uint32_t* res_off, const char* subj_ids,
                                 const uint32_t* subj_off, uint64_t n, zg_check* out) {
  for (uint64_t i = 0; i < n; ++i)
    if (res_off[i] > res_off[i + 1] || (subj_off && subj_off[i] > subj_off[i + 1]))
      return fail(ZG_EINVAL, "offsets must not decrease");
  // the literal fields once (same rules as resolve_check)
  ResolveMemo memo;
  zg_check base;
  const zg_rel_str proto{res_type, "", relation, subj_type, subj_off ? "" : subj_ids, subj_rel};
  resolve_check(e, proto, &base, &memo);
  const ResolveMemo::SchemaPart& part = memo.parts[memo.last];
  const bool ok = base.perm != kNone;
  const size_t one_len = subj_off ? 0 : std::strlen(subj_ids);
  auto run = [&](uint64_t b, uint64_t en) {
    for (uint64_t i = b; i < en; ++i) {
      out[i] = base;
      if (!ok) continue;  // unknown type / relation: answers ZG_ITEM_ERROR like the string path
      const char* rid = res_ids + res_off[i];
      const size_t rl = res_off[i + 1] - res_off[i];
      const char* sid = subj_off ? subj_ids + subj_off[i] : subj_ids;
      const size_t sl = subj_off ? subj_off[i + 1] - subj_off[i] : one_len;
      const uint32_t res = e->store.find(part.rt, rid, rl);
      uint32_t su = subj_off ? e->store.find(part.st, sid, sl) : memo.su;
      if (su == ZG_NO_OBJECT)
        su = (res == ZG_NO_OBJECT && part.rt == part.st && rl == sl && std::memcmp(rid, sid, rl) == 0) ? ZG_NO_OBJECT
                                                                                                       : ZG_NO_OBJECT - 1;
      out[i].res = res;
      out[i].subj = su;
    }
  };
  constexpr uint64_t kPerThread = 16384;
  const unsigned hw = std::thread::hardware_concurrency();
  const uint64_t nt = std::min<uint64_t>({n / kPerThread, hw ? hw : 1u, 16u});
  if (nt < 2) {
    run(0, n);
    return ZG_OK;
  }
  std::vector<std::thread> th;
  const uint64_t per = (n + nt - 1) / nt;
  for (uint64_t t = 1; t < nt; ++t) th.emplace_back(run, std::min(n, t * per), std::min(n, (t + 1) * per));
  run(0, std::min(n, per));
  for (auto& x : th) x.join();
  return ZG_OK;
}

#define PACKED_ARGS_OK() \
  (res_type && relation && subj_type && (n == 0 || (res_ids && res_off && subj_ids && out)))

extern "C" int zg_resolve_checks_packed(zg_engine* e, const char* res_type, const char* relation, const char* subj_type,
                                        const char* subj_rel, const char* res_ids, const uint32_t* res_off,
                                        const char* subj_ids, const uint32_t* subj_off, uint64_t n, zg_check* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!PACKED_ARGS_OK()) return fail(ZG_EINVAL, "NULL argument");
  if (n == 0) return ZG_OK;
  LOCK_NAMES_SHARED(e);
  return resolve_packed_locked(e, res_type, relation, subj_type, subj_rel, res_ids, res_off, subj_ids, subj_off, n, out);
}

extern "C" int zg_check_bulk_packed(zg_engine* e, const char* res_type, const char* relation, const char* subj_type,
                                    const char* subj_rel, const char* res_ids, const uint32_t* res_off, const char* subj_ids,
                                    const uint32_t* subj_off, uint64_t n, uint8_t* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!PACKED_ARGS_OK()) return fail(ZG_EINVAL, "NULL argument");
  if (n == 0) return ZG_OK;
  std::vector<zg_check> c(n);
  {
    int rc = publish_if_needed(e);
    if (rc) return rc;
    LOCK_NAMES_SHARED(e);
    rc = resolve_packed_locked(e, res_type, relation, subj_type, subj_rel, res_ids, res_off, subj_ids, subj_off, n, c.data());
    if (rc) return rc;
  }
  return zg_check_bulk(e, c.data(), n, out);
}

extern "C" int zg_check_bulk_str(zg_engine* e, const zg_rel_str* items, uint64_t n, uint8_t* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!items || !out) && n) return fail(ZG_EINVAL, "NULL argument");
  std::vector<zg_check> c(n);
  {
    int rc = publish_if_needed(e);
    if (rc) return rc;
    LOCK_NAMES_SHARED(e);
    resolve_checks_locked(e, items, n, c.data());
  }
  // Interned ids stay valid across later writes (interning only appends), so the launch can go
  // through the batcher like any zg_check_bulk call: concurrent string callers -- one goroutine per
  // rule check in the proxy, pkg/authz/check.go:77-93 -- share launches too.
  return zg_check_bulk(e, c.data(), n, out);
}

// ---- list templates: scanned items -> interned checks without per-item strings --------------

static void utf8_append(uint32_t cp, std::string* out) {
  if (cp < 0x80) {
    out->push_back(static_cast<char>(cp));
  } else if (cp < 0x800) {
    out->push_back(static_cast<char>(0xC0 | (cp >> 6)));
    out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  } else if (cp < 0x10000) {
    out->push_back(static_cast<char>(0xE0 | (cp >> 12)));
    out->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
    out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  } else {
    out->push_back(static_cast<char>(0xF0 | (cp >> 18)));
    out->push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
    out->push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
    out->push_back(static_cast<char>(0x80 | (cp & 0x3F)));
  }
}

static int hex4(const char* s) {
  int v = 0;
  for (int i = 0; i < 4; ++i) {
    const char h = s[i];
    int d = h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : -1;
    if (d < 0) return -1;
    v = v * 16 + d;
  }
  return v;
}

// Appends the decoded contents of a JSON string (the range zg_list_scan recorded, already validated).
static void json_unescape_append(const char* s, size_t n, std::string* out) {
  if (!std::memchr(s, '\\', n)) {
    out->append(s, n);
    return;
  }
  for (size_t i = 0; i < n;) {
    if (s[i] != '\\' || i + 1 >= n) {
      out->push_back(s[i++]);
      continue;
    }
    const char x = s[i + 1];
    i += 2;
    switch (x) {
      case 'b': out->push_back('\b'); break;
      case 'f': out->push_back('\f'); break;
      case 'n': out->push_back('\n'); break;
      case 'r': out->push_back('\r'); break;
      case 't': out->push_back('\t'); break;
      case 'u': {
        int v = i + 4 <= n ? hex4(s + i) : -1;
        if (v < 0) {
          utf8_append(0xFFFD, out);
          break;
        }
        i += 4;
        uint32_t cp = static_cast<uint32_t>(v);
        if (cp >= 0xD800 && cp < 0xDC00 && i + 6 <= n && s[i] == '\\' && s[i + 1] == 'u') {
          const int lo = hex4(s + i + 2);
          if (lo >= 0xDC00 && lo < 0xE000) {
            cp = 0x10000 + ((cp - 0xD800) << 10) + (static_cast<uint32_t>(lo) - 0xDC00);
            i += 6;
          }
        }
        if (cp >= 0xD800 && cp < 0xE000) cp = 0xFFFD;  // lone surrogate, as encoding/json decodes it
        utf8_append(cp, out);
        break;
      }
      default: out->push_back(x);  // \" \\ \/
    }
  }
}

extern "C" int zg_list_resolve(zg_engine* e, const char* body, size_t len, const zg_list_item* items, uint64_t n,
                               const zg_list_template* tpl, zg_check* out, uint8_t* checked) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!tpl || ((!body || !items || !out || !checked) && n)) return fail(ZG_EINVAL, "NULL argument");
  if (tpl->id_kind > ZG_ID_NAMESPACED_NAME) return fail(ZG_EINVAL, "unknown id_kind");
  LOCK_NAMES_SHARED(e);
  // the literal fields once; the resource id per item
  ResolveMemo memo;
  zg_check base;
  const zg_rel_str proto{tpl->res_type, "", tpl->permission, tpl->subj_type, tpl->subj_id, tpl->subj_rel};
  resolve_check(e, proto, &base, &memo);
  const ResolveMemo::SchemaPart& part = memo.parts[memo.last];
  const bool ok = base.perm != kNone;
  const uint32_t su = ok ? memo.su : ZG_NO_OBJECT;
  const std::string req_name = tpl->req_name ? tpl->req_name : "", req_ns = tpl->req_namespace ? tpl->req_namespace : "";
  std::string name, ns, id;
  for (uint64_t i = 0; i < n; ++i) {
    const zg_list_item& it = items[i];
    out[i] = base;
    checked[i] = 0;
    if (!(it.flags & ZG_ITEM_IS_OBJECT)) continue;
    if (it.name_off > len || it.name_len > len - it.name_off || it.ns_off > len || it.ns_len > len - it.ns_off)
      return fail(ZG_EINVAL, "item range outside the body");
    name.clear();
    ns.clear();
    if ((it.flags & ZG_ITEM_HAS_METADATA) && (it.flags & ZG_ITEM_RAW_NAMES)) {
      name.assign(body + it.name_off, it.name_len);
      ns.assign(body + it.ns_off, it.ns_len);
    } else if (it.flags & ZG_ITEM_HAS_METADATA) {
      json_unescape_append(body + it.name_off, it.name_len, &name);
      json_unescape_append(body + it.ns_off, it.ns_len, &ns);
    }
    // pkg/rules/rules.go:312-339
    if (name.empty()) name = req_name;
    if (ns.empty()) ns = req_ns;
    if (tpl->flags & ZG_TPL_CLEAR_NAMESPACE) ns.clear();
    if (tpl->id_kind == ZG_ID_NAMESPACED_NAME && !ns.empty()) {
      id = ns;
      id += '/';
      id += name;
    } else {
      id = name;
    }
    if (id.empty()) continue;  // "T:#perm@..." is not a relationship: the check is skipped
    checked[i] = 1;
    if (!ok) continue;  // unknown type / permission: answers ZG_ITEM_ERROR like the string path
    const uint32_t res = e->store.find(part.rt, id.data(), id.size());
    out[i].res = res;
    uint32_t s2 = su;
    if (s2 == ZG_NO_OBJECT)
      s2 = (res == ZG_NO_OBJECT && part.rt == part.st && tpl->subj_id && id == tpl->subj_id) ? ZG_NO_OBJECT : ZG_NO_OBJECT - 1;
    out[i].subj = s2;
  }
  return ZG_OK;
}

extern "C" int zg_list_postfilter(zg_engine* e, const char* body, size_t len, const zg_list_template* tpl, uint32_t n_tpl,
                                  char* out, size_t cap, size_t* out_len) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!body || !out_len || (!tpl && n_tpl)) return fail(ZG_EINVAL, "NULL argument");
  auto passthrough = [&]() {
    *out_len = len;
    if (len > cap || !out) return static_cast<int>(ZG_E2BIG);
    std::memcpy(out, body, len);
    return static_cast<int>(ZG_OK);
  };
  // one scan in the common case: an item is rarely shorter than 128 bytes (the buffer is not
  // initialised, so its untouched pages cost nothing)
  size_t icap = len / 128 + 16;
  std::unique_ptr<zg_list_item[]> items(new (std::nothrow) zg_list_item[icap]);
  if (!items) return fail(ZG_ENOMEM, "out of host memory");
  uint64_t ib = 0, ie = 0;
  int64_t n = zg_list_scan(body, len, ZG_LIST_ITEMS, items.get(), icap, &ib, &ie);
  if (n == ZG_E2BIG) {
    n = zg_list_scan(body, len, ZG_LIST_ITEMS, nullptr, 0, &ib, &ie);  // count, then one exact rescan
    if (n > 0) {
      icap = static_cast<size_t>(n);
      items.reset(new (std::nothrow) zg_list_item[icap]);
      if (!items) return fail(ZG_ENOMEM, "out of host memory");
      n = zg_list_scan(body, len, ZG_LIST_ITEMS, items.get(), icap, &ib, &ie);
    }
  }
  if (n < 0) return fail(ZG_EINVAL, "failed to parse list response");
  if (n == 0) return passthrough();  // no items array, or an empty one (postfilter.go:25-35)
  const uint64_t N = static_cast<uint64_t>(n);
  std::vector<uint8_t> keep(N, 1), checked(N);
  std::vector<zg_check> one(N), all;
  std::vector<uint64_t> owner;
  for (uint32_t t = 0; t < n_tpl; ++t) {
    int rc = zg_list_resolve(e, body, len, items.get(), N, &tpl[t], one.data(), checked.data());
    if (rc) return rc;
    for (uint64_t i = 0; i < N; ++i)
      if (checked[i]) {
        all.push_back(one[i]);
        owner.push_back(i);
      }
  }
  if (!all.empty()) {
    {
      int rc = publish_if_needed(e);
      if (rc) return rc;
    }
    std::vector<uint8_t> codes(all.size());
    int rc = zg_check_bulk(e, all.data(), all.size(), codes.data());  // ONE launch for the whole list
    if (rc) return rc;
    for (size_t k = 0; k < codes.size(); ++k)
      if (codes[k] != ZG_HAS_PERMISSION) keep[owner[k]] = 0;
  }
  int rc = zg_list_filter(body, len, items.get(), N, keep.data(), ib, ie, ZG_LIST_EMPTY_AS_NULL, out, cap, out_len);
  return rc == ZG_EINVAL ? fail(rc, "zg_list_filter: inconsistent item ranges") : rc;
}

// Answers a group of queued lookups under the engine lock: cache hits first, the rest in one batched launch
// sequence, then the self-membership of userset subjects (Check is the arbiter of LookupResources).
static void run_lookup_group(zg_engine* e, std::vector<zg_engine::LookupJob*>& group) {  // device lock held
  auto fail_all = [&](int rc, const std::string& msg) {
    for (auto* j : group) {
      j->rc = rc;
      j->err = msg;
    }
  };
  if (e->host_only) return fail_all(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  if (e->dev.shard_count > 1) return fail_all(ZG_EINVAL, kShardedMsg);
  if (e->dirty.load() || !e->dev.snap) {
    LOCK_NAMES_UNIQUE(e);
    int rc = publish_locked(e);
    if (rc) return fail_all(rc, g_err);
  }
  e->dev.now = now_of(e);
  std::vector<Device::LookupReq> reqs;
  std::vector<size_t> owner;  // job index of each request
  std::vector<int> dup_of(group.size(), -1);
  for (size_t i = 0; i < group.size(); ++i) {
    zg_engine::LookupJob* j = group[i];
    j->key.revision = e->revision;
    j->key.clock = e->clock ? e->clock : -static_cast<int64_t>(e->dev.now);  // wall clock: valid within the same second
    bool hit = false;
    for (const auto& c : e->lookup_cache)
      if (c.first == j->key) {
        j->ids = *c.second;
        hit = true;
        break;
      }
    if (hit) continue;
    for (size_t k = 0; k < i && dup_of[i] < 0; ++k)
      if (group[k]->key == j->key && dup_of[k] < 0) dup_of[i] = static_cast<int>(k);
    if (dup_of[i] >= 0) continue;
    zg_check proto{};
    proto.subj = j->key.subj;
    proto.perm = j->key.perm;
    proto.stype = j->key.stype;
    proto.srel = j->key.srel;
    reqs.push_back(Device::LookupReq{j->key.res_type, proto});
    owner.push_back(i);
  }
  if (!reqs.empty()) {
    std::vector<std::vector<uint32_t>> ids;
    std::vector<int> rcs;
    std::string err;
    int rc = ZG_OK;
    const size_t nd = 1 + e->replicas.size();
    if (nd == 1 || reqs.size() < 2) {
      rc = e->dev.lookup_batch(reqs, &ids, &rcs, &err);
    } else {
      // replicas: the lookups of the group are dealt out to the devices
      std::vector<std::vector<Device::LookupReq>> part(nd);
      std::vector<std::vector<size_t>> which(nd);
      for (size_t r = 0; r < reqs.size(); ++r) {
        part[r % nd].push_back(reqs[r]);
        which[r % nd].push_back(r);
      }
      std::vector<std::vector<std::vector<uint32_t>>> pids(nd);
      std::vector<std::vector<int>> prcs(nd);
      std::vector<int> drc(nd, ZG_OK);
      std::vector<std::string> derr(nd);
      const uint32_t now = e->dev.now;
      on_all_devices(e, [&](Device& dev, size_t i) {
        dev.now = now;
        if (!part[i].empty()) drc[i] = dev.lookup_batch(part[i], &pids[i], &prcs[i], &derr[i]);
      });
      ids.resize(reqs.size());
      rcs.assign(reqs.size(), ZG_OK);
      for (size_t i = 0; i < nd; ++i) {
        if (drc[i] && !rc) {
          rc = drc[i];
          err = derr[i];
        }
        for (size_t k = 0; k < which[i].size() && !drc[i]; ++k) {
          ids[which[i][k]] = std::move(pids[i][k]);
          rcs[which[i][k]] = prcs[i][k];
        }
      }
    }
    if (rc) return fail_all(rc, err);
    for (size_t r = 0; r < reqs.size(); ++r) {
      zg_engine::LookupJob* j = group[owner[r]];
      if (rcs[r]) {
        j->rc = rcs[r];
        j->err = "LookupResources: a candidate could not be decided (max dispatch depth / work budget exceeded)";
        continue;
      }
      j->ids = std::move(ids[r]);
      // A userset subject T:x#r is a member of T:x#P for every relation r inlined into P's union, with or
      // without relationships, and the reverse walk only reaches x through stored edges: ask the check
      // kernel about the one candidate res = subj.
      const zg_engine::LookupKey& k = j->key;
      if (k.srel != kNone && k.stype == k.res_type && k.subj < ZG_NO_OBJECT - 1 &&
          !std::binary_search(j->ids.begin(), j->ids.end(), k.subj)) {
        zg_check self = reqs[r].proto;
        self.res = k.subj;
        uint8_t code = 0;
        rc = e->dev.check_host(&self, 1, &code, &err);
        if (rc) {
          j->rc = rc;
          j->err = err;
          continue;
        }
        if (code == ZG_HAS_PERMISSION) j->ids.insert(std::upper_bound(j->ids.begin(), j->ids.end(), k.subj), k.subj);
      }
      e->lookup_cache.emplace_back(k, std::make_shared<const std::vector<uint32_t>>(j->ids));
      while (e->lookup_cache.size() > zg_engine::kLookupCache) e->lookup_cache.pop_front();
    }
  }
  for (size_t i = 0; i < group.size(); ++i) {
    zg_engine::LookupJob* j = group[i];
    if (dup_of[i] >= 0) {
      const zg_engine::LookupJob* src = group[dup_of[i]];
      j->ids = src->ids;
      j->rc = src->rc;
      j->err = src->err;
    }
    // never-written userset subject T:x#r (subj is the sentinel): still a member of T:x#perm when r is perm
    // itself or a relation inlined into its union
    if (!j->rc && j->want_self && j->key.srel != kNone && j->key.stype == j->key.res_type && j->key.subj >= ZG_NO_OBJECT - 1) {
      zg_check self{};
      self.res = self.subj = ZG_NO_OBJECT;  // the same never-written object on both sides
      self.perm = j->key.perm;
      self.stype = j->key.stype;
      self.srel = j->key.srel;
      uint8_t code = 0;
      std::string cerr;
      int crc = e->dev.check_host(&self, 1, &code, &cerr);
      if (crc) {
        j->rc = crc;
        j->err = cerr;
      } else {
        j->self_member = code == ZG_HAS_PERMISSION;
      }
    }
  }
}

// Queue one lookup; the first caller to arrive while nobody leads answers everything queued (<= 64 per
// launch sequence). Must be called WITHOUT e->mu held.
static int lookup_queued(zg_engine* e, uint16_t res_type, uint16_t perm, uint16_t stype, uint32_t subj, uint16_t srel,
                         std::vector<uint32_t>* ids, bool* self_member = nullptr) {
  if (self_member) *self_member = false;
  const Schema& sc = e->schema;
  if (res_type >= sc.types.size() || perm >= sc.slots.size() || sc.slots[perm].type != res_type)
    return fail(ZG_EINVAL, "unknown resource type or permission");
  if (stype >= sc.types.size() || (srel != kNone && (srel >= sc.slots.size() || sc.slots[srel].type != stype)))
    return fail(ZG_EINVAL, "unknown subject type or relation");
  zg_engine::LookupJob me;
  me.key.subj = subj;
  me.key.res_type = res_type;
  me.key.perm = perm;
  me.key.stype = stype;
  me.key.srel = srel;
  me.want_self = self_member != nullptr;
  auto& b = e->lookups;
  std::unique_lock<std::mutex> lk(b.m);
  b.queue.push_back(&me);
  if (b.leader_active) me.cv.wait(lk, [&] { return me.done || me.lead; });
  if (!me.done) {  // leader: see zg_check_bulk
    b.leader_active = true;
    std::vector<zg_engine::LookupJob*> group;
    lk.unlock();
    {
      std::lock_guard<FairMutex> dev_lock(e->mu);  // the device first, the group second (as for checks)
      lk.lock();
      const size_t take = std::min<size_t>(b.queue.size(), kMaxLookupGroup);
      group.assign(b.queue.begin(), b.queue.begin() + static_cast<long>(take));
      b.queue.erase(b.queue.begin(), b.queue.begin() + static_cast<long>(take));
      lk.unlock();
      run_lookup_group(e, group);
    }
    lk.lock();
    for (auto* j : group) {
      j->done = true;
      if (j != &me) j->cv.notify_one();
    }
    if (!b.queue.empty()) {
      b.queue.front()->lead = true;
      b.queue.front()->cv.notify_one();
    } else {
      b.leader_active = false;
    }
  }
  lk.unlock();
  if (me.rc) return fail(me.rc, me.err);
  *ids = std::move(me.ids);
  if (self_member) *self_member = me.self_member;
  return ZG_OK;
}

extern "C" int zg_lookup_resources(zg_engine* e, uint16_t res_type, uint16_t perm, uint16_t stype, uint32_t subj,
                                   uint16_t srel, uint32_t* out_ids, uint64_t cap, uint64_t* n_out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!n_out) return fail(ZG_EINVAL, "NULL n_out");
  // no lock here: the call QUEUES behind a running group (the group's leader validates the engine state)
  std::vector<uint32_t> ids;
  int rc = lookup_queued(e, res_type, perm, stype, subj, srel, &ids);
  if (rc) return rc;
  *n_out = ids.size();
  if (ids.size() > cap || (!out_ids && !ids.empty())) return ZG_E2BIG;
  if (!ids.empty()) std::memcpy(out_ids, ids.data(), ids.size() * 4);
  return ZG_OK;
}

// (resource type, permission, subject) strings -> ids, under the engine lock. su == ZG_NO_OBJECT: never written.
static int resolve_lookup_strs(zg_engine* e, const char* res_type, const char* perm, const char* subj_type, const char* subj_id,
                               const char* subj_rel, int* rt, int* p, int* st, uint16_t* sr, uint32_t* su) {
  int rc = publish_if_needed(e);
  if (rc) return rc;
  LOCK_NAMES_SHARED(e);
  const Schema& sc = e->schema;
  *rt = sc.type_id(res_type);
  *st = sc.type_id(subj_type);
  if (*rt < 0) return fail(ZG_EINVAL, std::string("object definition `") + res_type + "` not found");
  if (*st < 0) return fail(ZG_EINVAL, std::string("object definition `") + subj_type + "` not found");
  *p = sc.slot_id(*rt, perm);
  if (*p < 0) return fail(ZG_EINVAL, std::string("relation/permission `") + perm + "` not found under definition `" + res_type + "`");
  *sr = kNone;
  if (!none_rel(subj_rel)) {
    int s = sc.slot_id(*st, subj_rel);
    if (s < 0) return fail(ZG_EINVAL, std::string("relation `") + subj_rel + "` not found under definition `" + subj_type + "`");
    *sr = static_cast<uint16_t>(s);
  }
  *su = e->store.find(*st, subj_id);
  return ZG_OK;
}

extern "C" int zg_lookup_resources_str(zg_engine* e, const char* res_type, const char* perm, const char* subj_type,
                                       const char* subj_id, const char* subj_rel, char* buf, size_t cap, size_t* need,
                                       uint64_t* n_out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!res_type || !perm || !subj_type || !subj_id) return fail(ZG_EINVAL, "NULL argument");
  int rt, p, st;
  uint16_t sr;
  uint32_t su;
  int rc = resolve_lookup_strs(e, res_type, perm, subj_type, subj_id, subj_rel, &rt, &p, &st, &sr, &su);
  if (rc) return rc;
  std::vector<uint32_t> ids;
  bool self_member = false;
  rc = lookup_queued(e, static_cast<uint16_t>(rt), static_cast<uint16_t>(p), static_cast<uint16_t>(st),
                     su == ZG_NO_OBJECT ? ZG_NO_OBJECT - 1 : su, sr, &ids, &self_member);
  if (rc) return rc;
  std::vector<std::string> names;
  {
    LOCK_NAMES_SHARED(e);
    for (uint32_t id : ids) {
      std::string_view n;
      names.push_back(e->store.name(rt, id, &n) ? std::string(n) : std::to_string(id));
    }
  }
  // never-written userset subject that names itself
  if (su == ZG_NO_OBJECT && self_member) names.push_back(subj_id);
  size_t total = 1;
  for (const auto& n : names) total += n.size() + 1;
  if (need) *need = total;
  if (n_out) *n_out = names.size();
  if (total > cap || !buf) return ZG_E2BIG;
  size_t w = 0;
  for (const auto& n : names) {
    std::memcpy(buf + w, n.data(), n.size());
    w += n.size();
    buf[w++] = '\n';
  }
  buf[w] = 0;
  return ZG_OK;
}

// ---- pre-filter: LookupResources ids -> kept list items --------------------------------------

extern "C" int zg_list_keep_allowed(zg_engine* e, const char* body, size_t len, const zg_list_item* items, uint64_t n,
                                    uint32_t mode, const char* res_type, const char* req_namespace, const uint32_t* allowed,
                                    uint64_t n_allowed, const char* self_name, uint8_t* keep) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!res_type || ((!body || !items || !keep) && n) || (!allowed && n_allowed)) return fail(ZG_EINVAL, "NULL argument");
  if (mode > ZG_LIST_PROTOBUF) return fail(ZG_EINVAL, "unknown mode");
  LOCK_NAMES_SHARED(e);
  const int rt = e->schema.type_id(res_type);
  if (rt < 0) return fail(ZG_EINVAL, std::string("object definition `") + res_type + "` not found");
  const std::string req_ns = req_namespace ? req_namespace : "", self = self_name ? self_name : "";
  auto has = [&](const std::string& id) {
    if (!self.empty() && id == self) return true;
    const uint32_t oid = e->store.find(rt, id.data(), id.size());
    return oid != ZG_NO_OBJECT && std::binary_search(allowed, allowed + n_allowed, oid);
  };
  std::string name, ns, id;
  for (uint64_t i = 0; i < n; ++i) {
    const zg_list_item& it = items[i];
    keep[i] = 0;
    if (!(it.flags & ZG_ITEM_IS_OBJECT)) return fail(ZG_EINVAL, "failed to decode response body: element is not an object");
    if (mode == ZG_LIST_TABLE_ROWS && !(it.flags & ZG_ITEM_HAS_OBJECT))
      return fail(ZG_EINVAL, "error decoding partial object metadata from table row");
    if (it.name_off > len || it.name_len > len - it.name_off || it.ns_off > len || it.ns_len > len - it.ns_off)
      return fail(ZG_EINVAL, "item range outside the body");
    name.clear();
    ns.clear();
    if ((it.flags & ZG_ITEM_HAS_METADATA) && (it.flags & ZG_ITEM_RAW_NAMES)) {
      name.assign(body + it.name_off, it.name_len);  // protobuf strings: plain bytes
      ns.assign(body + it.ns_off, it.ns_len);
    } else if (it.flags & ZG_ITEM_HAS_METADATA) {
      json_unescape_append(body + it.name_off, it.name_len, &name);
      json_unescape_append(body + it.ns_off, it.ns_len, &ns);
    }
    // an allowed id maps to (text before the last '/', text after it): a name holding a '/' maps from none
    if (name.empty() || name.find('/') != std::string::npos) continue;
    id = ns;
    id += '/';
    id += name;
    bool k = has(id);
    if (!k && ns == req_ns) k = has(name);  // cluster-scoped ids take the request's namespace (lookups.go:117-127)
    keep[i] = k ? 1 : 0;
  }
  return ZG_OK;
}

extern "C" int zg_list_prefilter(zg_engine* e, const char* body, size_t len, uint32_t mode, const zg_list_template* tpl,
                                 char* out, size_t cap, size_t* out_len) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!body || !out_len || !tpl) return fail(ZG_EINVAL, "NULL argument");
  if (!tpl->res_type || !tpl->permission || !tpl->subj_type || !tpl->subj_id) return fail(ZG_EINVAL, "NULL template field");
  if (mode > ZG_LIST_PROTOBUF) return fail(ZG_EINVAL, "unknown mode");
  // 1. the allowed ids (one LookupResources on the GPU; concurrent list requests share launches)
  std::vector<uint32_t> ids;
  std::string self;
  int rt = -1;
  {
    int p, st;
    uint16_t sr;
    uint32_t su;
    int rc = resolve_lookup_strs(e, tpl->res_type, tpl->permission, tpl->subj_type, tpl->subj_id, tpl->subj_rel, &rt, &p, &st,
                                 &sr, &su);
    if (rc) return rc;
    bool self_member = false;
    rc = lookup_queued(e, static_cast<uint16_t>(rt), static_cast<uint16_t>(p), static_cast<uint16_t>(st),
                       su == ZG_NO_OBJECT ? ZG_NO_OBJECT - 1 : su, sr, &ids, &self_member);
    if (rc) return rc;
    if (su == ZG_NO_OBJECT && self_member) self = tpl->subj_id;  // never-written userset subject naming itself
    // lookups.go:106-109: an id that yields no name fails the whole pre-filter
    LOCK_NAMES_SHARED(e);
    for (uint32_t id : ids) {
      std::string_view nm;
      if (e->store.name(rt, id, &nm) && nm.back() == '/') return fail(ZG_EINVAL, "unable to determine name for resource");
    }
  }
  // 2. scan
  size_t icap = len / 128 + 16;
  std::unique_ptr<zg_list_item[]> items(new (std::nothrow) zg_list_item[icap]);
  if (!items) return fail(ZG_ENOMEM, "out of host memory");
  uint64_t ib = 0, ie = 0;
  int64_t n = zg_list_scan(body, len, mode, items.get(), icap, &ib, &ie);
  if (n == ZG_E2BIG) {
    n = zg_list_scan(body, len, mode, nullptr, 0, &ib, &ie);
    if (n > 0) {
      icap = static_cast<size_t>(n);
      items.reset(new (std::nothrow) zg_list_item[icap]);
      if (!items) return fail(ZG_ENOMEM, "out of host memory");
      n = zg_list_scan(body, len, mode, items.get(), icap, &ib, &ie);
    }
  }
  if (n < 0) return fail(ZG_EINVAL, "failed to decode response body");
  if (ie == 0) {  // no such array: the body passes through
    *out_len = len;
    if (len > cap || !out) return ZG_E2BIG;
    std::memcpy(out, body, len);
    return ZG_OK;
  }
  // 3. keep + splice
  const uint64_t N = static_cast<uint64_t>(n);
  std::vector<uint8_t> keep(N);
  int rc = zg_list_keep_allowed(e, body, len, items.get(), N, mode, tpl->res_type, tpl->req_namespace, ids.data(), ids.size(),
                                self.empty() ? nullptr : self.c_str(), keep.data());
  if (rc) return rc;
  rc = zg_list_filter(body, len, items.get(), N, keep.data(), ib, ie, 0, out, cap, out_len);
  return rc == ZG_EINVAL ? fail(rc, "zg_list_filter: inconsistent item ranges") : rc;
}

extern "C" int zg_stats_get(zg_engine* e, zg_stats* out) {
  if (!e || !out) return fail(ZG_EINVAL, "NULL argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (!e->host_only) e->dev.finish_timing();
  std::memset(out, 0, sizeof *out);
  out->checks = e->dev.checks;
  out->launches = e->dev.launches;
  out->passes = e->dev.passes;
  out->revision = e->revision;
  out->last_alg_bytes = e->dev.last_alg_bytes;
  out->last_kernel_ms = e->dev.last_ms;
  out->coalesced_launches = e->dev.coalesced_launches;
  out->coalesced_requests = e->dev.coalesced_requests;
  out->split_batches = e->dev.split_batches;
  out->delta_publishes = e->dev.delta_publishes;
  out->full_publishes = e->dev.full_publishes;
  out->last_publish_ms = e->dev.last_build_ms;
  out->streamed_calls = e->dev.streamed_calls;
  out->lookup_batches = e->dev.lookup_batches;
  out->lookups_batched = e->dev.lookups_batched;
  if (!e->host_only) e->dev.read_events(&out->stack_spills, &out->memo_batches);
  out->devices = 1 + e->replicas.size();
  for (auto& r : e->replicas) {
    const Device& d = r->dev;
    out->checks += d.checks;
    out->launches += d.launches;
    out->passes += d.passes;
    out->split_batches += d.split_batches;
    out->delta_publishes += d.delta_publishes;
    out->full_publishes += d.full_publishes;
    out->streamed_calls += d.streamed_calls;
    out->lookup_batches += d.lookup_batches;
    out->lookups_batched += d.lookups_batched;
  }
  if (e->dev.snap) {
    out->tuples = e->dev.snap->n_tuples;
    out->snapshot_bytes = e->dev.snap->bytes;
  }
  return ZG_OK;
}

extern "C" int zg_shard_pass(zg_engine* e, const zg_check* queries, uint64_t n, int level, uint64_t* n_sub) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!queries && n) || !n_sub || level < 0 || level > ZG_MAX_DEPTH + 2) return fail(ZG_EINVAL, "bad argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  if (!e->dev.snap) return fail(ZG_ENOSNAPSHOT, "no snapshot published (call zg_publish)");
  e->dev.now = now_of(e);
  std::string err;
  int rc = e->dev.shard_pass(queries, n, level, n_sub, &err);
  return rc ? fail(rc, err) : ZG_OK;
}
extern "C" int zg_shard_subqueries(zg_engine* e, int level, zg_check* out, uint64_t n) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!out && n) || level < 0) return fail(ZG_EINVAL, "bad argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  std::string err;
  int rc = e->dev.shard_subqueries(level, out, n, &err);
  return rc ? fail(rc, err) : ZG_OK;
}
extern "C" int zg_shard_fold(zg_engine* e, int level, const uint8_t* child_vals, uint64_t n_sub, uint8_t* out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if ((!child_vals && n_sub) || level < 0) return fail(ZG_EINVAL, "bad argument");
  std::lock_guard<FairMutex> g(e->mu);
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");
  std::string err;
  int rc = e->dev.shard_fold(level, child_vals, n_sub, out, &err);
  return rc ? fail(rc, err) : ZG_OK;
}

// ---- sharded store, device-resident protocol (no host staging between passes) ----
#define SHARD_PROLOGUE()                                                                                              \
  NEED_SCHEMA(e, ZG_ENOSCHEMA);                                                                                       \
  std::lock_guard<FairMutex> g(e->mu);                                                                               \
  if (e->host_only) return fail(ZG_ECUDA, "host-only engine: no CUDA device, and libzgpu has no CPU fallback");      \
  if (!e->dev.snap) return fail(ZG_ENOSNAPSHOT, "no snapshot published (call zg_publish)");

extern "C" int zg_shard_route_dev(zg_engine* e, const zg_check* d_items, uint64_t n, int level, uint32_t n_dest,
                                  zg_check* d_routed, uint32_t* d_src, uint64_t* counts) {
  if (!counts || ((!d_routed || !d_src) && n)) return fail(ZG_EINVAL, "NULL argument");
  SHARD_PROLOGUE();
  std::string err;
  int rc = e->dev.route_by_owner(d_items, n, level, n_dest, d_routed, d_src, counts, &err);
  return rc ? fail(rc, err) : ZG_OK;
}
extern "C" int zg_shard_pass_dev(zg_engine* e, const zg_check* d_queries, uint64_t n, int level, uint64_t* n_sub) {
  if ((!d_queries && n) || !n_sub || level < 0 || level > ZG_MAX_DEPTH + 2) return fail(ZG_EINVAL, "bad argument");
  SHARD_PROLOGUE();
  e->dev.now = now_of(e);
  std::string err;
  int rc = e->dev.shard_pass_dev(d_queries, n, level, n_sub, &err);
  return rc ? fail(rc, err) : ZG_OK;
}
extern "C" int zg_shard_fold_dev(zg_engine* e, int level, const uint8_t* d_child_vals, const uint32_t* d_src, uint64_t n_sub,
                                 uint8_t* d_out, int final_codes) {
  if (((!d_child_vals || !d_src) && n_sub) || level < 0) return fail(ZG_EINVAL, "bad argument");
  SHARD_PROLOGUE();
  std::string err;
  int rc = e->dev.shard_fold_dev(level, d_child_vals, d_src, n_sub, d_out, final_codes != 0, &err);
  return rc ? fail(rc, err) : ZG_OK;
}
extern "C" int zg_shard_unroute_dev(zg_engine* e, const uint32_t* d_src, const uint8_t* d_val, uint64_t n, uint8_t* d_out) {
  if ((!d_src || !d_val || !d_out) && n) return fail(ZG_EINVAL, "NULL argument");
  SHARD_PROLOGUE();
  std::string err;
  int rc = e->dev.unroute(d_src, d_val, n, d_out, &err);
  return rc ? fail(rc, err) : ZG_OK;
}

extern "C" int zg_debug_row(zg_engine* e, uint16_t rel_slot, uint32_t res, uint32_t cls, uint32_t* out, uint64_t cap,
                            uint64_t* n_out) {
  NEED_SCHEMA(e, ZG_ENOSCHEMA);
  if (!n_out) return fail(ZG_EINVAL, "NULL n_out");
  LOCK_DEVICE(e);
  LOCK_NAMES_UNIQUE(e);
  e->keep_built = true;
  if (e->last_built.rels.empty() && !e->schema.rel_slots.empty()) {
    e->last_built = e->store.build();
    if (!e->last_built.err.empty()) return fail(ZG_EINVAL, e->last_built.err);
  }
  const Schema& sc = e->schema;
  if (rel_slot >= sc.slots.size() || sc.slots[rel_slot].is_perm) return fail(ZG_EINVAL, "not a relation");
  const DRel& r = e->last_built.rels[sc.slots[rel_slot].rel_index];
  *n_out = 0;
  if (cls & 0x80000000u) {  // reverse row: `res` is the SUBJECT id, result = resources listing it
    cls &= 0x7FFFFFFFu;
    if (cls >= r.ncls) return fail(ZG_EINVAL, "class out of range");
    const DCls& c = e->last_built.cls[r.cls_begin + cls];
    const uint32_t row = c.sslot == kWildcard ? 0u : res;
    if (row >= c.nsubj) return ZG_OK;
    const uint64_t ri = c.rrow_base + uint64_t(row) * c.rstride;
    uint32_t b = e->last_built.rrow_ptr[ri], en = e->last_built.rrow_ptr[ri + 1];
    *n_out = en - b;
    if (en - b > cap) return ZG_E2BIG;
    for (uint32_t i = b; i < en; ++i) out[i - b] = e->last_built.rcol[i];
    return ZG_OK;
  }
  if (cls >= r.ncls) return fail(ZG_EINVAL, "class out of range");
  if (res >= r.nres) return ZG_OK;
  uint64_t idx = r.row_base + uint64_t(res) * r.stride + cls;
  uint32_t b = e->last_built.row_ptr[idx], en = e->last_built.row_ptr[idx + 1];
  *n_out = en - b;
  if (en - b > cap) return ZG_E2BIG;
  for (uint32_t i = b; i < en; ++i) out[i - b] = e->last_built.col[i];
  return ZG_OK;
}

extern "C" const char* zg_build_info(void) { return zg::build_info(); }

extern "C" void* zg_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    g_err = "cudaMallocHost failed";
    return nullptr;
  }
  return p;
}
extern "C" void zg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
