// emu_main.cc -- drives spicedb-kubeapi-proxy_b200/csrc/kernels.cuh under the CPU SIMT emulator (cuda_emu.h).
// TEST INFRASTRUCTURE ONLY (built by tests/test_kernel_emulation.py into its own .so; never part of libzgpu.so).
// The host side is the product's own schema compiler and store (schema.cc, store.cc: plain C++) and a restatement of
// the launch sequences of device.cu (pass loop, sub-query split, fold; the batched LookupResources rounds) over
// host memory.
#define ZG_EMULATE 1
#include "cuda_emu.h"

#include <algorithm>
#include <cstdio>
#include <string>

#include "../../spicedb-kubeapi-proxy_b200/csrc/delta.cuh"
#include "../../spicedb-kubeapi-proxy_b200/csrc/kernels.cuh"
#include "../../spicedb-kubeapi-proxy_b200/csrc/store.h"

using namespace zg;

struct EmuOpts {
  int32_t invert;        // direction-optimised probes
  uint32_t memo_after;   // iterations before the path memo switches on
  uint32_t memo_entries; // per warp, power of two, 0 = off
  uint64_t subq_cap;
  uint32_t budget;
  uint32_t now;
  uint32_t grid;         // blocks
  uint32_t spill_cap;
  uint32_t streamed;     // run the STREAMED variant with everything already "landed"
};

struct ShardLevel {
  std::vector<zg_check> q, raised;
  std::vector<uint32_t> parent;
  std::vector<uint8_t> val;
  uint64_t nq = 0, nsub = 0;
};

struct Emu {
  uint32_t shard_count = 1, shard_rank = 0;
  std::vector<ShardLevel> levels;
  Schema sc;
  Store st;
  HostSnapshot h;
  std::vector<uint8_t> blob;
  bool published = false;
  uint64_t alg_bytes = 0, spills = 0, memo_batches = 0, passes = 0, splits = 0;
  uint32_t flags = 0;
};

extern "C" {

void* emu_create(const char* schema, char* err, size_t cap) {
  Emu* e = new Emu();
  std::string m = e->sc.parse(schema);
  if (!m.empty()) {
    std::snprintf(err, cap, "%s", m.c_str());
    delete e;
    return nullptr;
  }
  e->st.reset(&e->sc);
  return e;
}
void emu_destroy(void* h) { delete static_cast<Emu*>(h); }
int emu_type_id(void* h, const char* n) { return static_cast<Emu*>(h)->sc.type_id(n); }
int emu_slot_id(void* h, int t, const char* n) { return static_cast<Emu*>(h)->sc.slot_id(t, n); }
int emu_load(void* h, const zg_tuple* t, const uint32_t* ex, uint64_t n, char* err, size_t cap) {
  std::string m = static_cast<Emu*>(h)->st.load(t, ex, n);
  if (!m.empty()) {
    std::snprintf(err, cap, "%s", m.c_str());
    return -1;
  }
  return 0;
}
int emu_publish(void* h) {
  Emu* e = static_cast<Emu*>(h);
  e->h = e->st.build();
  if (!e->h.err.empty()) return -1;
  e->blob = e->sc.blob(e->h.rels, e->h.cls);
  // rows are streamed with aligned 128-bit loads that may read (and ignore) a few words past the last row
  e->h.col.reserve(e->h.col.size() + 16);
  e->h.rcol.reserve(e->h.rcol.size() + 16);
  e->published = true;
  return 0;
}
uint64_t emu_stat(void* h, int which) {
  Emu* e = static_cast<Emu*>(h);
  switch (which) {
    case 0: return e->alg_bytes;
    case 1: return e->spills;
    case 2: return e->memo_batches;
    case 3: return e->passes;
    case 4: return e->splits;
    case 5: return e->flags;
  }
  return 0;
}

}  // extern "C"

namespace {

struct Ctrl {
  unsigned long long next = 0, subq_count = 0, alg_bytes = 0;
  uint32_t flags = 0;
  unsigned long long events[2] = {0, 0};
  unsigned long long ready = ~0ull;
};

int run_pass(Emu* e, const EmuOpts& o, const zg_check* queries, uint64_t nq, uint8_t* val, uint8_t* out, bool raw,
             zg_check* subq, uint32_t* subq_parent, uint64_t* n_sub, bool count) {
  KParams p{};
  static const uint32_t zero = 0;
  auto ptr = [&](const std::vector<uint32_t>& v) { return v.empty() ? &zero : v.data(); };
  p.row_ptr = ptr(e->h.row_ptr);
  p.col = ptr(e->h.col);
  p.exp = e->sc.has_expiry ? ptr(e->h.exp) : nullptr;
  p.rrow_ptr = ptr(e->h.rrow_ptr);
  p.rcol = ptr(e->h.rcol);
  p.invert = o.invert;
  p.prog = e->blob.data();
  p.prog_bytes = static_cast<uint32_t>(e->blob.size());
  p.queries = queries;
  p.nq = nq;
  p.L = e->sc.max_leaves;
  p.val = val;
  p.out = out;
  p.final_codes = 1;
  Ctrl c;
  p.next = &c.next;
  p.subq_count = &c.subq_count;
  p.alg_bytes = &c.alg_bytes;
  p.flags = &c.flags;
  p.events = c.events;
  p.ready = &c.ready;
  const uint32_t grid = o.grid ? o.grid : 2;
  const size_t warps = size_t(grid) * kWarpsPerBlock;
  std::vector<uint4> spill(warps * o.spill_cap);
  std::vector<unsigned long long> memo(warps * std::max<uint32_t>(o.memo_entries, 1));
  p.spill = spill.data();
  p.spill_cap = o.spill_cap;
  p.subq = subq;
  p.subq_parent = subq_parent;
  p.subq_cap = subq ? o.subq_cap : 0;
  p.now = o.now;
  p.budget = o.budget;
  p.raw_items = raw;
  p.memo = memo.data();
  p.memo_entries = o.memo_entries;
  p.memo_after = o.memo_after;
  p.shard_count = e->shard_count;
  p.shard_rank = e->shard_rank;
  if (e->shard_count > 1) p.invert = 0;  // a shard does not hold the subject's reverse rows
  const size_t sm = p.prog_bytes + size_t(kWarpsPerBlock) * kWarpSmem;
  if (count) zg_emu::launch(check_kernel<true, false>, p, grid, kThreads, sm);
  else if (o.streamed) zg_emu::launch(check_kernel<false, true>, p, grid, kThreads, sm);
  else zg_emu::launch(check_kernel<false, false>, p, grid, kThreads, sm);
  e->alg_bytes += c.alg_bytes;
  e->spills += c.events[0];
  e->memo_batches += c.events[1];
  e->flags |= c.flags;
  if (n_sub) *n_sub = c.subq_count;
  return 0;
}

// device.cu check_device over host memory
int check(Emu* e, const EmuOpts& o, const zg_check* items, uint64_t n, uint8_t* out, bool count) {
  if (n == 0) return 0;
  const uint32_t L = e->sc.max_leaves;
  if (!e->sc.has_nonpure) return run_pass(e, o, items, n, nullptr, out, true, nullptr, nullptr, nullptr, count);
  std::vector<std::vector<zg_check>> q(1);
  std::vector<std::vector<uint32_t>> parent(1);
  std::vector<std::vector<uint8_t>> val;
  std::vector<uint64_t> nq;
  const zg_check* queries = items;
  uint64_t cur = n;
  bool raw = true;
  for (size_t lv = 0; cur > 0; ++lv) {
    if (lv > ZG_MAX_DEPTH + 1) return -2;
    nq.push_back(cur);
    val.emplace_back(cur * L + 4, 0);
    q.emplace_back(o.subq_cap);
    parent.emplace_back(o.subq_cap);
    uint64_t n_sub = 0;
    run_pass(e, o, queries, cur, val[lv].data(), lv == 0 ? out : nullptr, raw, q[lv + 1].data(), parent[lv + 1].data(), &n_sub, count);
    if (n_sub > o.subq_cap) {
      if (n < 2) return -3;
      ++e->splits;
      const uint64_t half = n / 2;
      int r = check(e, o, items, half, out, count);
      if (r) return r;
      return check(e, o, items + half, n - half, out + half, count);
    }
    cur = n_sub;
    queries = q[lv + 1].data();
    raw = false;
    if (cur) ++e->passes;
  }
  if (nq.size() > 1) {
    struct FoldArgs {
      const uint8_t* prog;
      const zg_check* queries;
      unsigned long long n;
      uint32_t L;
      const uint8_t* val;
      uint8_t* out;
      const uint32_t* parent;
      uint8_t* parent_val;
    };
    for (size_t lv = nq.size(); lv-- > 0;) {
      FoldArgs a{e->blob.data(), lv == 0 ? items : q[lv].data(), nq[lv], L, val[lv].data(), lv == 0 ? out : nullptr,
                 lv == 0 ? nullptr : parent[lv].data(), lv == 0 ? nullptr : val[lv - 1].data()};
      zg_emu::launch(+[](const FoldArgs& f) { fold_kernel(f.prog, f.queries, f.n, f.L, f.val, f.out, f.parent, f.parent_val, 0); },
                     a, static_cast<unsigned>((nq[lv] + 255) / 256), 256, 0);
    }
  }
  return 0;
}

}  // namespace

extern "C" int emu_check(void* h, const zg_check* items, uint64_t n, uint8_t* out, const EmuOpts* o, int count) {
  Emu* e = static_cast<Emu*>(h);
  if (!e->published) return -1;
  e->flags = 0;
  return check(e, *o, items, n, out, count != 0);
}

// device.cu lookup_batch over host memory: K lookups, multi-source reverse walk, one verification, sort.
// out_ids: concatenated ascending ids per lookup; out_counts[K]; rcs[K]. Returns 0, or -7 when cap is too small.
extern "C" int emu_lookup_batch(void* h, const uint16_t* res_types, const zg_check* protos, uint32_t K, const EmuOpts* o,
                                uint64_t batch_cap, uint32_t* out_ids, uint64_t cap, uint64_t* out_counts, int* rcs) {
  Emu* e = static_cast<Emu*>(h);
  if (!e->published || K == 0 || K > kMaxLookupBatch) return -1;
  std::vector<unsigned long long> base(e->h.n_objects.size() + 1, 0);
  for (size_t t = 0; t < e->h.n_objects.size(); ++t) base[t + 1] = base[t] + ((uint64_t(e->h.n_objects[t]) + 31) & ~31ull);
  const unsigned long long words = (base.back() + 31) / 32 + 4;
  std::vector<uint32_t> visited(words * K, 0);
  std::vector<unsigned long long> front[2] = {std::vector<unsigned long long>(batch_cap), std::vector<unsigned long long>(batch_cap)};
  std::vector<zg_check> cand(batch_cap);
  std::vector<uint8_t> owner(batch_cap), codes(batch_cap);
  std::vector<LookupParam> lp(K);
  std::vector<unsigned long long> counts(ZG_MAX_DEPTH + 4, 0);
  for (uint32_t i = 0; i < K; ++i) {
    lp[i] = LookupParam{protos[i].subj, res_types[i], protos[i].perm, protos[i].stype, protos[i].srel};
    front[0][i] = (static_cast<unsigned long long>(i) << 48) | (static_cast<unsigned long long>(protos[i].stype) << 32) | protos[i].subj;
  }
  counts[0] = K;
  static const uint32_t zero = 0;
  unsigned long long cand_count = 0;
  uint32_t flags = 0;
  MrbfsParams p{};
  p.rrow_ptr = e->h.rrow_ptr.empty() ? &zero : e->h.rrow_ptr.data();
  p.rcol = e->h.rcol.empty() ? &zero : e->h.rcol.data();
  p.prog = e->blob.data();
  p.lk = lp.data();
  p.counts = counts.data();
  p.cap = batch_cap;
  p.visited = visited.data();
  p.words = words;
  p.type_bit_base = base.data();
  p.cand = cand.data();
  p.cand_owner = owner.data();
  p.cand_count = &cand_count;
  p.cand_cap = batch_cap;
  p.flags = &flags;
  int cur = 0;
  for (int level = 0; level <= ZG_MAX_DEPTH + 1; ++level) {
    p.in = front[cur].data();
    p.out = front[cur ^ 1].data();
    p.level = level;
    zg_emu::launch(mrbfs_expand_kernel, p, 2, 256, 0);
    cur ^= 1;
    if (flags & 16u) return -8;  // overflow: the product halves the batch
    if (counts[level + 1] == 0) break;
  }
  for (uint32_t i = 0; i < K; ++i) {
    out_counts[i] = 0;
    rcs[i] = 0;
  }
  if (cand_count == 0) return 0;
  e->flags = 0;
  int rc = check(e, *o, cand.data(), cand_count, codes.data(), false);
  if (rc) return rc;
  std::vector<unsigned long long> keys(cand_count), per(kMaxLookupBatch, 0);
  unsigned long long n_keys = 0, err_mask = 0;
  struct KeyArgs {
    const zg_check* cand;
    const uint8_t* owner;
    const uint8_t* codes;
    unsigned long long n;
    unsigned long long *keys, *n_keys, *per, *err;
  } ka{cand.data(), owner.data(), codes.data(), cand_count, keys.data(), &n_keys, per.data(), &err_mask};
  zg_emu::launch(+[](const KeyArgs& a) { lookup_keys_kernel(a.cand, a.owner, a.codes, a.n, a.keys, a.n_keys, a.per, a.err); }, ka,
                 static_cast<unsigned>((cand_count + 255) / 256), 256, 0);
  std::sort(keys.begin(), keys.begin() + n_keys);
  if (n_keys > cap) return -7;
  for (unsigned long long i = 0; i < n_keys; ++i) out_ids[i] = static_cast<uint32_t>(keys[i]);
  for (uint32_t i = 0; i < K; ++i) {
    out_counts[i] = per[i];
    if ((err_mask >> i) & 1ull) rcs[i] = ZG_EDEPTH;
  }
  return 0;
}

// ---- incremental publish: build.cu gpu_apply_delta / merge_direction over host memory -----------------------------

namespace {

struct MergeArgs {
  const uint32_t *old_a, *old_b;
  uint32_t *new_a, *new_b;
  uint32_t n_old;
  const uint32_t *ins_pos, *del_pos;
  uint32_t n_ins, n_del;
};
struct RowsArgs {
  uint32_t* row_ptr;
  unsigned long long first, pool;
  const unsigned long long *ins_key, *del_key;
  uint32_t n_ins, n_del;
};

// returns the number of delta entries that contradict the old arrays
uint32_t merge_direction_emu(const HostDelta& h, bool with_exp, std::vector<uint32_t>& row_ptr, uint64_t pool,
                             std::vector<uint32_t>& edges, std::vector<uint32_t>& exp) {
  const uint32_t ni = static_cast<uint32_t>(h.ins_key.size()), nd = static_cast<uint32_t>(h.del_key.size()),
                 nt = with_exp ? static_cast<uint32_t>(h.tch_key.size()) : 0u;
  const uint64_t n_old = edges.size(), n_new = n_old + ni - nd;
  std::vector<uint32_t> ins_pos(ni + 1), del_pos(nd + 1), tch_pos(nt + 1);
  uint32_t bad = 0;
  static const uint32_t zero = 0;
  DeltaDev d{h.ins_key.data(), h.ins_val.data(), with_exp ? h.ins_exp.data() : nullptr, ins_pos.data(), ni,
             h.del_key.data(), h.del_val.data(), del_pos.data(), nd,
             h.tch_key.data(), h.tch_val.data(), h.tch_exp.data(), tch_pos.data(), nt};
  const size_t keys = size_t(ni) + nd + nt;
  struct LocArgs {
    const uint32_t *row_ptr, *col;
    DeltaDev d;
    uint32_t* bad;
  } la{row_ptr.data(), edges.empty() ? &zero : edges.data(), d, &bad};
  if (keys) zg_emu::launch(+[](const LocArgs& a) { delta_locate_kernel(a.row_ptr, a.col, a.d, a.bad); }, la,
                           static_cast<unsigned>((keys + 255) / 256), 256, 0);
  if (ni || nd) {
    std::vector<uint32_t> ne(std::max<uint64_t>(n_new, 1), 0xDEADBEEFu), nx(with_exp ? std::max<uint64_t>(n_new, 1) : 0);
    MergeArgs ma{edges.data(), with_exp ? exp.data() : nullptr, ne.data(), with_exp ? nx.data() : nullptr,
                 static_cast<uint32_t>(n_old), ins_pos.data(), del_pos.data(), ni, nd};
    if (n_old)
      zg_emu::launch(+[](const MergeArgs& a) { delta_merge_kernel(a.old_a, a.new_a, a.old_b, a.new_b, a.n_old, a.ins_pos, a.n_ins, a.del_pos, a.n_del); },
                     ma, static_cast<unsigned>((n_old + kMergeTile - 1) / kMergeTile), 256, 0);
    struct InsArgs {
      uint32_t *new_a, *new_b;
      const uint32_t *ins_pos, *ins_val, *ins_exp;
      uint32_t n_ins;
      const uint32_t* del_pos;
      uint32_t n_del;
    } ia{ne.data(), with_exp ? nx.data() : nullptr, ins_pos.data(), h.ins_val.data(), with_exp ? h.ins_exp.data() : nullptr, ni,
         del_pos.data(), nd};
    if (ni) zg_emu::launch(+[](const InsArgs& a) { delta_insert_kernel(a.new_a, a.new_b, a.ins_pos, a.ins_val, a.ins_exp, a.n_ins, a.del_pos, a.n_del); },
                           ia, (ni + 255) / 256, 256, 0);
    struct TchArgs {
      uint32_t* new_exp;
      const uint32_t *tch_pos, *tch_exp;
      uint32_t n_tch;
      const uint32_t* ins_pos;
      uint32_t n_ins;
      const uint32_t* del_pos;
      uint32_t n_del;
    } ta{nx.data(), tch_pos.data(), h.tch_exp.data(), nt, ins_pos.data(), ni, del_pos.data(), nd};
    if (nt) zg_emu::launch(+[](const TchArgs& a) { delta_touch_kernel(a.new_exp, a.tch_pos, a.tch_exp, a.n_tch, a.ins_pos, a.n_ins, a.del_pos, a.n_del); },
                           ta, (nt + 255) / 256, 256, 0);
    unsigned long long first = pool;
    if (ni) first = std::min<unsigned long long>(first, h.ins_key.front());
    if (nd) first = std::min<unsigned long long>(first, h.del_key.front());
    ++first;
    RowsArgs ra{row_ptr.data(), first, pool, h.ins_key.data(), h.del_key.data(), ni, nd};
    if (first <= pool)
      zg_emu::launch(+[](const RowsArgs& a) { delta_rows_kernel(a.row_ptr, a.first, a.pool, a.ins_key, a.n_ins, a.del_key, a.n_del); }, ra,
                     static_cast<unsigned>((pool + 1 - first + kMergeTile - 1) / kMergeTile), 256, 0);
    ne.resize(n_new);
    edges.swap(ne);
    if (with_exp) {
      nx.resize(n_new);
      exp.swap(nx);
    }
  } else if (nt) {
    struct TchArgs {
      uint32_t* new_exp;
      const uint32_t *tch_pos, *tch_exp;
      uint32_t n_tch;
    } ta{exp.data(), tch_pos.data(), h.tch_exp.data(), nt};
    zg_emu::launch(+[](const TchArgs& a) { delta_touch_kernel(a.new_exp, a.tch_pos, a.tch_exp, a.n_tch, nullptr, 0, nullptr, 0); }, ta,
                   (nt + 255) / 256, 256, 0);
  }
  return bad;
}

}  // namespace

extern "C" {

// Transactional interned updates against the emulator's store (journal on once a snapshot is published).
int emu_apply(void* h, const zg_update* u, uint64_t n, char* err, size_t cap) {
  Emu* e = static_cast<Emu*>(h);
  int code = 0;
  std::string m = e->st.apply(u, n, &code);
  if (!m.empty()) {
    std::snprintf(err, cap, "%s", m.c_str());
    return code ? code : -1;
  }
  return 0;
}
void emu_journal_start(void* h) { static_cast<Emu*>(h)->st.journal_clear(); }

// Merges the journal into the emulator's resident arrays, then compares every array with a fresh host build.
// 0 = identical; 1 = the layout changed (the product rebuilds); negative = mismatch / inconsistency.
int emu_merge_and_verify(void* h, char* err, size_t cap) {
  Emu* e = static_cast<Emu*>(h);
  if (!e->published) return -1;
  if (!e->st.journal_ok || !e->st.layout_stable()) return 1;
  HostSnapshot lay = e->st.layout();
  if (lay.n_objects != e->h.n_objects) return 1;
  HostDelta f, r;
  std::vector<uint32_t> cls_delta;
  std::string perr = prepare_delta(e->st, e->sc, lay, &f, &r, &cls_delta);
  if (perr == "relayout") return 1;
  if (!perr.empty()) {
    std::snprintf(err, cap, "%s", perr.c_str());
    return -2;
  }
  const bool with_exp = e->sc.has_expiry;
  std::vector<uint32_t> none;
  uint32_t bad = merge_direction_emu(f, with_exp, e->h.row_ptr, lay.pool, e->h.col, e->h.exp);
  bad += merge_direction_emu(r, false, e->h.rrow_ptr, lay.rpool, e->h.rcol, none);
  if (bad) {
    std::snprintf(err, cap, "%u delta entries contradict the resident arrays", bad);
    return -3;
  }
  e->st.journal_clear();
  HostSnapshot want = e->st.build();
  auto same = [&](const std::vector<uint32_t>& a, const std::vector<uint32_t>& b, const char* what) {
    if (a == b) return true;
    size_t i = 0;
    while (i < a.size() && i < b.size() && a[i] == b[i]) ++i;
    std::snprintf(err, cap, "%s differs at %zu (sizes %zu / %zu)", what, i, a.size(), b.size());
    return false;
  };
  if (!same(e->h.row_ptr, want.row_ptr, "row_ptr") || !same(e->h.col, want.col, "col") ||
      !same(e->h.rrow_ptr, want.rrow_ptr, "rrow_ptr") || !same(e->h.rcol, want.rcol, "rcol") ||
      (with_exp && !same(e->h.exp, want.exp, "exp")))
    return -4;
  // class emptiness (drives the program's steps) from the per-class deltas
  for (size_t c = 0; c < want.cls.size(); ++c) {
    e->h.cls[c].flags = want.cls[c].flags;
  }
  e->h.n_tuples = want.n_tuples;
  e->blob = e->sc.blob(e->h.rels, e->h.cls);
  return 0;
}

}  // extern "C"

// ---- object-hash sharded store: device.cu route_by_owner / shard_pass_dev / shard_fold_dev / unroute over host memory

extern "C" {

void emu_set_shard(void* h, uint32_t rank, uint32_t count) {
  Emu* e = static_cast<Emu*>(h);
  e->shard_rank = rank;
  e->shard_count = count;
  e->st.shard_rank = rank;
  e->st.shard_count = count;
}

int emu_shard_route(void* h, const zg_check* items, uint64_t n, int level, uint32_t n_dest, zg_check* routed, uint32_t* src,
                    uint64_t* counts) {
  Emu* e = static_cast<Emu*>(h);
  for (uint32_t d = 0; d < n_dest; ++d) counts[d] = 0;
  if (!items) {
    if (level < 0 || static_cast<size_t>(level) >= e->levels.size() || n != e->levels[level].nsub) return -1;
    items = e->levels[level].raised.data();
  }
  if (n == 0) return 0;
  std::vector<unsigned long long> cnt(kMaxRouteDest, 0), cursor(kMaxRouteDest, 0);
  struct CntArgs {
    const zg_check* items;
    unsigned long long n;
    uint32_t n_dest;
    unsigned long long* counts;
  } ca{items, n, n_dest, cnt.data()};
  zg_emu::launch(+[](const CntArgs& a) { route_count_kernel(a.items, a.n, a.n_dest, a.counts); }, ca, 2, 256, 0);
  unsigned long long acc = 0;
  for (uint32_t d = 0; d < n_dest; ++d) {
    counts[d] = cnt[d];
    cursor[d] = acc;
    acc += cnt[d];
  }
  struct ScArgs {
    const zg_check* items;
    unsigned long long n;
    uint32_t n_dest;
    unsigned long long* cursor;
    zg_check* routed;
    uint32_t* src;
  } sa{items, n, n_dest, cursor.data(), routed, src};
  zg_emu::launch(+[](const ScArgs& a) { route_scatter_kernel(a.items, a.n, a.n_dest, a.cursor, a.routed, a.src); }, sa, 2, 256, 0);
  return 0;
}

int emu_shard_pass(void* h, const zg_check* queries, uint64_t n, int level, const EmuOpts* o, uint64_t* n_sub) {
  Emu* e = static_cast<Emu*>(h);
  if (!e->published) return -1;
  if (e->levels.size() <= static_cast<size_t>(level)) e->levels.resize(level + 1);
  ShardLevel& L = e->levels[level];
  L.nq = n;
  L.nsub = 0;
  *n_sub = 0;
  if (n == 0) return 0;
  L.q.assign(queries, queries + n);
  L.val.assign(n * e->sc.max_leaves + 4, 0);
  L.raised.assign(o->subq_cap, zg_check{});
  L.parent.assign(o->subq_cap, 0);
  uint64_t ns = 0;
  // final_codes does not matter here: out is null, the fold writes the level's outputs
  run_pass(e, *o, L.q.data(), n, L.val.data(), nullptr, level == 0, L.raised.data(), L.parent.data(), &ns, false);
  if (ns > o->subq_cap) return -3;
  L.nsub = ns;
  *n_sub = ns;
  return 0;
}

int emu_shard_fold(void* h, int level, const uint8_t* child_vals, const uint32_t* src, uint64_t n_sub, uint8_t* out, int final_codes) {
  Emu* e = static_cast<Emu*>(h);
  if (level < 0 || static_cast<size_t>(level) >= e->levels.size() || n_sub != e->levels[level].nsub) return -1;
  ShardLevel& L = e->levels[level];
  if (L.nq == 0) return 0;
  if (n_sub) {
    struct OrArgs {
      const uint32_t *parent, *src;
      const uint8_t* child;
      unsigned long long n;
      uint8_t* val;
    } oa{L.parent.data(), src, child_vals, n_sub, L.val.data()};
    zg_emu::launch(+[](const OrArgs& a) { or_children_src_kernel(a.parent, a.src, a.child, a.n, a.val); }, oa,
                   static_cast<unsigned>((n_sub + 255) / 256), 256, 0);
  }
  struct FoldArgs {
    const uint8_t* prog;
    const zg_check* q;
    unsigned long long n;
    uint32_t L;
    const uint8_t* val;
    uint8_t* out;
    int raw;
  } fa{e->blob.data(), L.q.data(), L.nq, e->sc.max_leaves, L.val.data(), out, final_codes ? 0 : 1};
  zg_emu::launch(+[](const FoldArgs& a) { fold_kernel(a.prog, a.q, a.n, a.L, a.val, a.out, nullptr, nullptr, a.raw); }, fa,
                 static_cast<unsigned>((L.nq + 255) / 256), 256, 0);
  return 0;
}

int emu_unroute(void* h, const uint32_t* src, const uint8_t* val, uint64_t n, uint8_t* out) {
  (void)h;
  if (n == 0) return 0;
  struct UArgs {
    const uint32_t* src;
    const uint8_t* val;
    unsigned long long n;
    uint8_t* out;
  } ua{src, val, n, out};
  zg_emu::launch(+[](const UArgs& a) { unroute_kernel(a.src, a.val, a.n, a.out); }, ua, static_cast<unsigned>((n + 255) / 256), 256, 0);
  return 0;
}

}  // extern "C"
