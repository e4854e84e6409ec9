// Host-side sanitizer run (ASan + UBSan) over the schema compiler and the relationship store + CSR builder
// (csrc/schema.cc, csrc/store.cc: plain C++, the same translation units libzgpu.so links). Built and run by
// tests/test_host_logic.py::test_host_code_under_sanitizers.
//   1. schema texts, valid and byte-mutated: parse() returns "" or an error, never crashes; a schema that
//      compiles yields a program blob.
//   2. random TOUCH / CREATE / DELETE streams against a model (std::map): after every batch the store's live
//      set and every forward and reverse CSR row equal the model's.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../spicedb-kubeapi-proxy_b200/csrc/schema.h"
#include "../../spicedb-kubeapi-proxy_b200/csrc/store.h"

using namespace zg;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}

static const char* kSchemas[] = {
    "definition user {}\ndefinition group { relation member: user | group#member }\n"
    "definition doc { relation viewer: user | group#member | user:*  relation banned: user  relation parent: doc\n"
    "  permission view = (viewer + parent->view) - banned  permission both = viewer & banned  permission none = nil }\n",
    "use expiration\ndefinition user {}\ndefinition ns { relation creator: user  relation viewer: user with expiration\n"
    "  permission view = viewer + creator  permission admin = creator }\n"
    "definition pod { relation namespace: ns  relation viewer: user  permission view = viewer + namespace->view\n"
    "  permission any = namespace.any(admin) }\n",
    "/* comment */ definition a {} // trailing\ndefinition b { relation r: a | b#r | a:*  permission p = r  permission q = p + r->p }\n",
};

static int fuzz_schemas(long iters) {
  const char alphabet[] = "{}()|#:*+-&>. \n/abdefinitonrlpmsuwhx_0";
  long compiled = 0;
  for (long it = 0; it < iters; ++it) {
    std::string s = kSchemas[rnd() % 3];
    const int muts = static_cast<int>(rnd() % 4);
    for (int m = 0; m < muts && !s.empty(); ++m) {
      const size_t p = rnd() % s.size();
      switch (rnd() % 5) {
        case 0: s[p] = alphabet[rnd() % (sizeof(alphabet) - 1)]; break;
        case 1: s.erase(p, 1 + rnd() % 4); break;
        case 2: s.insert(p, 1, alphabet[rnd() % (sizeof(alphabet) - 1)]); break;
        case 3: s.resize(p); break;
        default: s.insert(p, s.substr(p / 2, rnd() % 40)); break;
      }
    }
    Schema sc;
    if (!sc.parse(s).empty()) continue;
    ++compiled;
    Store st;
    st.reset(&sc);
    HostSnapshot h = st.build();
    if (!h.err.empty()) return std::printf("empty store failed to build: %s\n", h.err.c_str()), 1;
    if (sc.blob(h.rels, h.cls).empty()) return std::printf("compiled schema without a program blob\n"), 1;
  }
  if (compiled < iters / 20) return std::printf("only %ld of %ld schema texts compiled\n", compiled, iters), 1;
  std::printf("schemas: %ld texts, %ld compiled\n", iters, compiled);
  return 0;
}

struct ModelKey {
  uint16_t rel, stype, srel;
  uint32_t res, subj;
  bool operator<(const ModelKey& o) const {
    return std::tie(rel, stype, srel, res, subj) < std::tie(o.rel, o.stype, o.srel, o.res, o.subj);
  }
};

static int verify(const Schema& sc, const Store& st, const std::map<ModelKey, uint32_t>& model) {
  HostSnapshot h = st.build();
  if (!h.err.empty()) return std::printf("build failed: %s\n", h.err.c_str()), 1;
  if (h.n_tuples != model.size() || st.size() != model.size())
    return std::printf("live count: snapshot %llu store %llu model %zu\n", (unsigned long long)h.n_tuples,
                       (unsigned long long)st.size(), model.size()), 1;
  // expected rows from the model
  std::map<std::tuple<int, uint32_t, int>, std::vector<uint32_t>> fwd, rev;  // (rel index, object, class) -> ids
  for (const auto& kv : model) {
    const ModelKey& k = kv.first;
    const int cls = sc.class_of(k.rel, k.stype, k.srel);
    if (cls < 0) return std::printf("model holds a relationship the schema forbids\n"), 1;
    const int ri = sc.slots[k.rel].rel_index;
    fwd[{ri, k.res, cls}].push_back(k.srel == kWildcard ? 0u : k.subj);
    rev[{ri, k.srel == kWildcard ? 0u : k.subj, cls}].push_back(k.res);
  }
  uint64_t seen_f = 0, seen_r = 0;
  for (size_t ri = 0; ri < h.rels.size(); ++ri) {
    const DRel& r = h.rels[ri];
    for (uint32_t res = 0; res < r.nres; ++res)
      for (uint32_t c = 0; c < r.ncls; ++c) {
        const uint64_t idx = r.row_base + uint64_t(res) * r.stride + c;
        if (idx + 1 >= h.row_ptr.size()) return std::printf("row_ptr index out of range\n"), 1;
        const uint32_t b = h.row_ptr[idx], e = h.row_ptr[idx + 1];
        if (b > e || e > h.col.size()) return std::printf("forward row bounds\n"), 1;
        std::vector<uint32_t> got(h.col.begin() + b, h.col.begin() + e);
        if (!std::is_sorted(got.begin(), got.end())) return std::printf("forward row not sorted\n"), 1;
        auto it = fwd.find({static_cast<int>(ri), res, static_cast<int>(c)});
        std::vector<uint32_t> want = it == fwd.end() ? std::vector<uint32_t>() : it->second;
        std::sort(want.begin(), want.end());
        if (got != want) return std::printf("forward row (%zu,%u,%u) differs: %zu vs %zu\n", ri, res, c, got.size(), want.size()), 1;
        seen_f += got.size();
      }
    for (uint32_t c = 0; c < r.ncls; ++c) {
      const DCls& dc = h.cls[r.cls_begin + c];
      for (uint32_t s = 0; s < dc.nsubj; ++s) {
        const uint64_t rri = dc.rrow_base + uint64_t(s) * dc.rstride;
        if (rri + 1 >= h.rrow_ptr.size()) return std::printf("rrow_ptr index out of range\n"), 1;
        const uint32_t b = h.rrow_ptr[rri], e = h.rrow_ptr[rri + 1];
        if (b > e || e > h.rcol.size()) return std::printf("reverse row bounds\n"), 1;
        std::vector<uint32_t> got(h.rcol.begin() + b, h.rcol.begin() + e);
        if (!std::is_sorted(got.begin(), got.end())) return std::printf("reverse row not sorted\n"), 1;
        auto it = rev.find({static_cast<int>(ri), s, static_cast<int>(c)});
        std::vector<uint32_t> want = it == rev.end() ? std::vector<uint32_t>() : it->second;
        std::sort(want.begin(), want.end());
        if (got != want) return std::printf("reverse row (%zu,%u,%u) differs\n", ri, s, c), 1;
        seen_r += got.size();
      }
    }
  }
  if (seen_f != model.size() || seen_r != model.size())
    return std::printf("rows hold %llu forward / %llu reverse entries, model %zu\n", (unsigned long long)seen_f,
                       (unsigned long long)seen_r, model.size()), 1;
  // expirations ride with the entries when the schema uses them
  if (!h.exp.empty() && h.exp.size() != h.col.size()) return std::printf("exp not parallel to col\n"), 1;
  return 0;
}

static int fuzz_store(int schema_i, long batches) {
  Schema sc;
  const std::string err = sc.parse(kSchemas[schema_i]);
  if (!err.empty()) return std::printf("seed schema %d: %s\n", schema_i, err.c_str()), 1;
  Store st;
  st.reset(&sc);
  // the (relation, subject kind) combinations the schema allows
  struct Kind { uint16_t rel, stype, srel; bool expiry; };
  std::vector<Kind> kinds;
  for (size_t s = 0; s < sc.slots.size(); ++s)
    if (!sc.slots[s].is_perm)
      for (const auto& c : sc.slots[s].classes) kinds.push_back({static_cast<uint16_t>(s), c.stype, c.sslot, c.expiry});
  if (kinds.empty()) return std::printf("no relation classes\n"), 1;
  std::vector<uint32_t> nobj(sc.types.size(), 0);
  auto obj = [&](int type) {  // mostly existing objects, sometimes a new one (interned by name)
    if (nobj[type] == 0 || rnd() % 8 == 0) {
      const uint32_t id = st.intern(type, "o" + std::to_string(nobj[type]) + std::string(rnd() % 20, 'x'));
      if (id != nobj[type]) std::abort();
      return nobj[type]++;
    }
    return static_cast<uint32_t>(rnd() % nobj[type]);
  };
  std::map<ModelKey, uint32_t> model;
  long applied = 0, rejected = 0;
  for (long b = 0; b < batches; ++b) {
    std::vector<zg_update> ups(1 + rnd() % 24);
    std::map<ModelKey, uint32_t> next = model;
    bool expect_fail = false;
    std::vector<ModelKey> in_batch;
    for (auto& u : ups) {
      const Kind& k = kinds[rnd() % kinds.size()];
      u.t.rel = k.rel;
      u.t.stype = k.stype;
      u.t.srel = k.srel;
      u.t.flags = 0;
      u.t.res = obj(sc.slots[k.rel].type);
      u.t.subj = k.srel == kWildcard ? static_cast<uint32_t>(rnd()) : obj(k.stype);
      u.op = static_cast<uint32_t>(rnd() % 3);
      u.expires_at = k.expiry && u.op != ZG_OP_DELETE && rnd() % 2 ? 1000u + static_cast<uint32_t>(rnd() % 1000) : 0u;
      const ModelKey mk{k.rel, k.stype, k.srel, u.t.res, k.srel == kWildcard ? 0u : u.t.subj};
      // a CREATE of something that exists BEFORE the batch fails the whole batch (nothing applied)
      if (u.op == ZG_OP_CREATE && model.count(mk)) expect_fail = true;
      in_batch.push_back(mk);
    }
    // the C ABI rejects two updates of one relationship in a write; keep batches free of them here too
    std::vector<ModelKey> sorted = in_batch;
    std::sort(sorted.begin(), sorted.end());
    bool dup = false;
    for (size_t i = 1; i < sorted.size(); ++i)
      if (!(sorted[i - 1] < sorted[i])) dup = true;
    if (dup) continue;
    for (size_t i = 0; i < ups.size(); ++i) {
      if (ups[i].op == ZG_OP_DELETE) next.erase(in_batch[i]);
      else next[in_batch[i]] = ups[i].expires_at;
    }
    int code = 0;
    std::vector<uint8_t> changed;
    const std::string e2 = st.apply(ups.data(), ups.size(), &code, &changed);
    if (expect_fail) {
      if (e2.empty() || code != ZG_EEXIST) return std::printf("CREATE of an existing relationship was accepted\n"), 1;
      ++rejected;
    } else {
      if (!e2.empty()) return std::printf("valid batch rejected: %s\n", e2.c_str()), 1;
      for (size_t i = 0; i < ups.size(); ++i) {
        // kinds: touched (the relationship existed) / inserted / deleted / unchanged (DELETE of nothing)
        const bool was = model.count(in_batch[i]) != 0;
        const uint8_t want = ups[i].op == ZG_OP_DELETE ? (was ? Store::kDeleted : Store::kUnchanged)
                                                       : (was ? Store::kTouched : Store::kInserted);
        if (changed[i] != want) return std::printf("changed[] mask wrong\n"), 1;
      }
      model.swap(next);
      ++applied;
    }
    if (b % 16 == 0 || b + 1 == batches)
      if (verify(sc, st, model)) return 1;
  }
  // filters against the model
  for (int q = 0; q < 200; ++q) {
    Store::Filter f;
    const Kind& k = kinds[rnd() % kinds.size()];
    f.res_type = sc.slots[k.rel].type;
    if (rnd() % 2) f.rel = k.rel;
    if (rnd() % 2 && nobj[f.res_type]) { f.has_res = true; f.res = static_cast<uint32_t>(rnd() % nobj[f.res_type]); }
    if (rnd() % 2) f.subj_type = k.stype;
    std::vector<uint64_t> idx;
    st.match(f, 0, &idx);
    size_t want = 0;
    for (const auto& kv : model) {
      const ModelKey& m = kv.first;
      if (sc.slots[m.rel].type != f.res_type) continue;
      if (f.rel >= 0 && m.rel != f.rel) continue;
      if (f.has_res && m.res != f.res) continue;
      if (f.subj_type >= 0 && m.stype != f.subj_type) continue;
      ++want;
    }
    if (idx.size() != want) return std::printf("filter matched %zu, model %zu\n", idx.size(), want), 1;
  }
  std::printf("store[%d]: %ld batches applied, %ld rejected, %zu live relationships\n", schema_i, applied, rejected, model.size());
  return 0;
}

int main(int argc, char** argv) {
  const long scale = argc > 1 ? std::atol(argv[1]) : 1;
  if (fuzz_schemas(3000 * scale)) return 1;
  for (int s = 0; s < 3; ++s)
    if (fuzz_store(s, 400 * scale)) return 1;
  std::printf("ok\n");
  return 0;
}
