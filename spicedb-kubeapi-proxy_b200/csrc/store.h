// store.h -- host side of the relationship store: object interning, the mutable
// relationship set (what WriteRelationships / DeleteRelationships act on,
// pkg/authz/distributedtx/activity.go:54-76) and the CSR snapshot builder whose
// output the engine uploads to HBM.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/zgpu.h"
#include "schema.h"

namespace zg {

struct Key {
  uint64_t hi, lo;  // hi = rel<<32 | stype<<16 | srel ; lo = res<<32 | subj
  bool operator==(const Key& o) const { return hi == o.hi && lo == o.lo; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t x = k.hi * 0x9E3779B97F4A7C15ull ^ (k.lo + 0xD6E8FEB86659FD93ull);
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 29;
    return static_cast<size_t>(x);
  }
};
inline Key key_of(const zg_tuple& t) {
  return Key{(uint64_t(t.rel) << 32) | (uint64_t(t.stype) << 16) | t.srel,
             (uint64_t(t.res) << 32) | (t.srel == kWildcard ? 0u : t.subj)};
}

// CSR snapshot in host memory, ready to upload. Layout: DESIGN.md "Data layout in HBM".
struct HostSnapshot {
  std::vector<uint32_t> row_ptr;  // pool: per data relation, nres x ncls offsets, +1 sentinel
  std::vector<uint32_t> col;      // subject object ids, sorted within (object, class)
  std::vector<uint32_t> exp;      // parallel to col when the schema uses expiration, else empty
  std::vector<DRel> rels;
  std::vector<DCls> cls;          // schema classes + rrow_base / CF_EMPTY for this snapshot
  std::vector<uint32_t> rrow_ptr; // reverse CSR: per class, one row per SUBJECT object (+1 sentinel)
  std::vector<uint32_t> rcol;     // resource object ids, ascending within a row
  std::vector<std::vector<uint32_t>> resources;  // per type: ids that are the resource of >= 1 relationship
  std::vector<uint32_t> n_objects;               // per type
  uint64_t n_tuples = 0;
  uint64_t pool = 0, rpool = 0;            // entries of row_ptr / rrow_ptr (each has +1 sentinel)
  std::vector<uint64_t> type_base;         // per type: first row_ptr index of its objects
  std::vector<uint32_t> type_ncls;         // per type: row stride (classes of all its relations)
  std::string err;  // non-empty: build failed
};

class Store {
 public:
  void reset(const Schema* s);
  // Drops every relationship; interned object names (and therefore ids) stay.
  void clear_relationships();

  uint32_t intern(int type, const std::string& id);
  uint32_t find(int type, const std::string& id) const;  // ZG_NO_OBJECT
  uint32_t find(int type, const char* id, size_t len) const;  // no temporary string: the check-ingress path
  // false for a numeric-only object (bulk loads) or an id out of range
  bool name(int type, uint32_t id, std::string_view* out) const;

  // Returns "" or an error message; `code` receives the ZG_* code.
  std::string validate(const zg_tuple& t, bool has_expiry) const;
  std::string load(const zg_tuple* t, const uint32_t* expires, uint64_t n);
  // `changed` (optional, n entries): nonzero where the update took effect (a DELETE of a relationship that
  // does not exist changes nothing; TOUCH and CREATE always count): kTouched = the relationship existed
  // (only its expiration may differ), kInserted, kDeleted.
  enum : uint8_t { kUnchanged = 0, kTouched = 1, kInserted = 2, kDeleted = 3 };
  std::string apply(const zg_update* u, uint64_t n, int* code, std::vector<uint8_t>* changed = nullptr);

  // Journal of what apply() changed since journal_clear(): what an incremental publish merges into the
  // device snapshot instead of rebuilding it (build.cu gpu_apply_delta). journal_ok turns false when
  // something the journal cannot express happened (bulk load, clear, re-layout).
  struct JournalEntry {
    zg_tuple t;
    uint32_t expires;
    uint8_t kind;  // kTouched / kInserted / kDeleted
  };
  std::vector<JournalEntry> journal;
  bool journal_ok = false;
  void journal_clear() {
    journal.clear();
    journal_ok = true;
  }

  // Live relationships matching a filter (unset field = -1 / ZG_NO_OBJECT-1 sentinel via has_*).
  struct Filter {
    int res_type = -1;
    bool has_res = false;
    uint32_t res = 0;
    int rel = -1;  // slot
    int subj_type = -1;
    bool has_subj = false;
    uint32_t subj = 0;
    bool subj_wildcard = false;
    bool has_srel = false;
    uint16_t srel = kNone;
    bool impossible = false;  // names an object/relation that does not exist
  };
  void match(const Filter& f, uint32_t now, std::vector<uint64_t>* idx) const;

  // Sizes and bases only (n_objects, rels, cls with rrow_base / nsubj and CF_EMPTY set on every
  // class): what both the host builder and the GPU builder (build.cu) start from.
  // Object counts are CAPACITIES: they grow in steps (1/8 + 4096 beyond what is needed) and never shrink, so
  // that relationships on newly created objects -- the proxy's common write -- do not move any row table.
  HostSnapshot layout() const;
  // true when layout() would still return the capacities it returned last time
  bool layout_stable() const;
  HostSnapshot build() const;
  uint64_t size() const { return live_; }

  const Schema* schema = nullptr;
  // object-hash sharding: relationships whose resource id % shard_count != shard_rank are dropped
  // (object counts still follow every relationship seen, so ids mean the same on all shards)
  uint32_t shard_count = 1, shard_rank = 0;
  std::vector<zg_tuple> tuples;  // flags bit0 = dead (tombstone)
  std::vector<uint32_t> expires; // parallel to tuples (always sized like tuples)

 private:
  void ensure_index();
  // Interned names of one type. Each name lives once, in `arena`, as a record [u32 id][u32 len][bytes];
  // `table` is an open-addressing index over the records (slot = record offset + 1, 0 = empty, power-of-two
  // size, load <= 1/2), so a lookup hashes the caller's bytes in place and touches two cache lines on a hit:
  // the slot and the record. `rec_of` maps an id back to its record (0 = no name) for the output paths.
  struct TypeObjs {
    std::vector<char> arena;
    std::vector<uint64_t> table;
    std::vector<uint64_t> rec_of;
    uint32_t n_interned = 0;
    uint32_t n_numeric = 0;  // max numeric id seen in bulk loads + 1
    uint32_t lookup(const char* s, size_t n) const;
    void insert(uint64_t rec);  // record offset
    uint32_t n_ids() const { return static_cast<uint32_t>(rec_of.size()); }
  };
  std::vector<TypeObjs> objs_;
  mutable std::vector<uint32_t> obj_cap_;  // per type: capacity handed out by the last layout()
  std::unordered_map<Key, uint64_t, KeyHash> index_;
  bool indexed_ = true;  // index_ covers every live tuple
  uint64_t live_ = 0;
};

}  // namespace zg
