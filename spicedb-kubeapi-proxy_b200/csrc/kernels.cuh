// kernels.cuh -- sm_100a CUDA kernels of the permission-check hot path.
//
// What this replaces in the reference: the per-item graph evaluation that
// CheckBulkPermissions / CheckPermission / LookupResources trigger inside the
// embedded SpiceDB (call sites pkg/authz/check.go:48, postfilter.go:134,
// watch.go:50, lookups.go:65; engine config pkg/spicedb/spicedb.go:25-56).
//
// Execution model (DESIGN.md "Kernels"):
//   * persistent grid, 148 SMs x 4 CTAs x 8 warps; every WARP owns a private LIFO of edge
//     RANGES in shared memory (spilling to a per-warp HBM area) and pulls batches of 32
//     checks from a global counter;
//   * admission loads each subject's reverse rows (its direct memberships) into a small
//     per-check set in shared memory: direct probes become shared-memory compares, and a
//     range whose children could only match directly is resolved by searching the
//     memberships' short reverse rows instead of visiting the children (meet in the middle);
//   * one iteration pops ranges worth <= 32 edges, assigns one edge per lane (coalesced reads
//     of each CSR segment), and every lane visits its child node: two row offsets per
//     non-empty edge class, ballot/popc-compacted push of the userset and arrow ranges;
//   * found / error state is two warp-uniform 32-bit masks; a found check's remaining ranges
//     are dropped when popped (short circuit); long-running batches switch on a lossy memo of
//     (check, slot, object, depth) visits so DAG-shaped data stays polynomial;
//   * no tensor cores, no floating point: memory-latency / issue-bound integer traversal.
#pragma once
#ifndef ZG_EMULATE  // tests/emu/: the same source under a CPU SIMT emulator (test infrastructure, never in libzgpu.so)
#include <cuda_runtime.h>
#define ZG_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#ifndef ZG_BLOCK_SHARED_U32
#define ZG_BLOCK_SHARED_U32(name, n) __shared__ uint32_t name[n]
#endif
#endif

#include <cstdint>

#include "../../include/zgpu.h"
#include "schema.h"

namespace zg {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
#ifndef ZG_MIN_BLOCKS
#define ZG_MIN_BLOCKS 4
#endif
#ifndef ZG_STACK_CAP
#define ZG_STACK_CAP 64
#endif
#ifndef ZG_RSET_CAP
#define ZG_RSET_CAP 16
#endif
#ifndef ZG_PRESENCE
#define ZG_PRESENCE 0  // presence word in front of the per-check set walk of a direct inverted probe (measured: -2 .. -3 %,
                       // profiles/r2f: the walk is short and the extra shared-memory word costs more than it saves)
#endif
#ifndef ZG_POP_FAST
#define ZG_POP_FAST 0  // pop: when every popped range holds <= 1 edge, skip the prefix scan / owner search (measured: no gain, r2g)
#endif
#ifndef ZG_L2_MATCH
#define ZG_L2_MATCH 0  // cooperative two-level meet: match.any instead of a search for ranges of <= 16 children.
                       // Measured 2x SLOWER (cfg3: 532 vs ~1 090 Mchecks/s, profiles/r2h): MATCH.ANY is a slow path on sm_100a
#endif
#ifndef ZG_L2_FILTER
#define ZG_L2_FILTER 1  // cooperative two-level meet: a 2 048-bit blocked Bloom filter of the range's children in shared memory
                        // in front of the binary search (the search then runs only when the filter lets an element through).
                        // Measured (profiles/r2j_ab_*): cfg3 1 056 -> 1 164 Mchecks/s, with ZG_L2_SPLIT 1 185; cfg4 unchanged
#endif
#ifndef ZG_L2_SPLIT
#define ZG_L2_SPLIT 1  // cooperative two-level meet: 32 / ng lanes per membership row (any ng) instead of a power of two
                       // (measured alone: cfg3 1 056 -> 1 079)
#endif
#ifndef ZG_L2_MODE
#define ZG_L2_MODE 2  // two-level meet: 2 = warp-cooperative (default), 1 = per-lane Bloom word + 128-bit streaming
                      // (measured slower: 919 vs 1 023 Mchecks/s on cfg3), 0 = per-lane sorted-segment intersection
#endif
constexpr int kMinBlocks = ZG_MIN_BLOCKS;  // resident CTAs per SM the register budget is tuned for
constexpr int kStackCap = ZG_STACK_CAP;    // range items per warp in shared memory
constexpr int kRsetCap = ZG_RSET_CAP;      // reverse-row entries kept per check (subject's direct memberships), <= 31
constexpr int kStateWords = 6;             // per-lane query state parked in shared memory between leaf passes
constexpr int kFCap = 64;                  // children of a range staged in shared memory by the two-level meet
constexpr int kFBloom = ZG_L2_FILTER ? 64 : 0;  // words of the filter over those children
constexpr size_t kWarpSmem = kStackCap * sizeof(uint4) + ((kRsetCap + kStateWords) * 32 + kFCap + kFBloom) * sizeof(uint32_t);
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint16_t kJobDepthMask = 0x00FF;  // zg_check.flags of a raised sub-query: hop depth

// value bits of a (query, leaf) pair and of a sub-query
constexpr uint8_t kValT = 1, kValE = 2;

struct KParams {
  const uint32_t* row_ptr;
  const uint32_t* col;
  const uint32_t* exp;  // nullptr unless the schema uses expiration
  const uint32_t* rrow_ptr;  // reverse CSR (subject -> resources), rows interleaved per subject
  const uint32_t* rcol;
  int invert;  // answer direct probes from the subject's reverse rows when they are small
  const uint8_t* prog;
  uint32_t prog_bytes;
  const zg_check* queries;
  unsigned long long nq;
  uint32_t L;    // leaf stride of val (Snapshot::max_leaves)
  uint8_t* val;  // [nq * L] value bits per (query, leaf): what the fold of a multi-pass run reads (may be null)
  uint8_t* out;  // [nq] v1 codes (final_codes) or value bits (may be null)
  int final_codes;
  unsigned long long* next;  // batch counter (zeroed before launch)
  uint4* spill;              // per-warp spill areas
  uint32_t spill_cap;        // items per warp
  // sub-queries raised when an edge leads into a non-pure permission (or to another shard)
  zg_check* subq;
  uint32_t* subq_parent;  // q * L + leaf of the raising (query, leaf)
  unsigned long long* subq_count;
  unsigned long long subq_cap;
  uint32_t* flags;  // bit0: spill overflow, bit1: sub-query overflow, bit2: budget hit
  uint32_t now;
  uint32_t budget;
  int raw_items;  // queries are caller-supplied zg_check items: flags are ignored
  // Per-warp lossy memo of (job, slot, object, depth) child visits, switched on only for
  // batches that run long (DAG-shaped data multiplies paths; see DESIGN.md "Path memo").
  unsigned long long* memo;
  uint32_t memo_entries;  // per warp, power of two (0 = disabled)
  uint32_t memo_after;    // iterations before the memo is switched on for a batch
  // Object-hash sharded store: this device only holds the relationships whose
  // RESOURCE it owns (owner = object id % shard_count); an edge to an object of another shard
  // becomes a sub-query that is routed to its owner between passes.
  uint32_t shard_count;   // 0 / 1 = not sharded
  uint32_t shard_rank;
  unsigned long long* alg_bytes;  // COUNT variant only
  unsigned long long* events;     // [0] stack spills, [1] batches that switched the path memo on (cumulative)
  // Streamed admission: the queries are still being copied from the host while the kernel runs. *ready =
  // number of leading queries that have landed (a 8-byte copy enqueued after each chunk on the copy stream);
  // a warp waits until its batch is there. nullptr: everything is resident.
  const unsigned long long* ready;
};

// View of the program blob. Only the base pointer is kept in registers; table addresses are
// formed from the header offsets on use (sixteen live 64-bit pointers cost 32 registers).
struct Prog {
  const uint8_t* b;
  __device__ __forceinline__ const DHeader* hdr() const { return reinterpret_cast<const DHeader*>(b); }
#define ZG_TABLE(name, type, field) \
  __device__ __forceinline__ const type* name() const { return reinterpret_cast<const type*>(b + hdr()->field); }
  ZG_TABLE(slots, DSlot, off_slots)
  ZG_TABLE(units, DUnit, off_units)
  ZG_TABLE(cls, DCls, off_cls)
  ZG_TABLE(members, uint16_t, off_members)
  ZG_TABLE(trees, DTree, off_trees)
  ZG_TABLE(tree_ops, DTreeOp, off_tree_ops)
  ZG_TABLE(leaf_units, uint16_t, off_leaf_units)
  ZG_TABLE(type_inv, DTypeInv, off_type_inv)
  ZG_TABLE(inv_cls, uint16_t, off_inv_cls)
  ZG_TABLE(steps, DStep, off_steps)
  ZG_TABLE(type_rcls, DTypeInv, off_type_rcls)
  ZG_TABLE(rcls, uint16_t, off_rcls)
#undef ZG_TABLE
};

__device__ __forceinline__ Prog make_prog(const uint8_t* b) { return Prog{b}; }

// range item: x = begin, y = end (edge indices into col), z = meta, w unused
//   meta: bits 0-4 job slot, 5-10 depth of the children, 11 class has expiry,
//         16-31 slot the children are visited at
__device__ __forceinline__ uint32_t make_meta(uint32_t jslot, uint32_t depth, bool expiry, uint32_t tslot) {
  return jslot | (depth << 5) | (expiry ? (1u << 11) : 0u) | (tslot << 16);
}

template <bool COUNT>
struct WarpCtx {
  uint4* stack;
  uint4* spill;
  int top, spill_top;
  uint32_t spill_cap;
  unsigned found, err;  // warp-uniform, indexed by job slot
  unsigned raised;      // job slots that raised a sub-query in this leaf pass (their value is not final)
  uint32_t my_subj, my_ss;  // this lane's own query: subject id, stype<<16 | srel
  // Direction-optimised probes: rset[i * 32 + job] = resource id of every direct relationship
  // of the job's subject (its reverse rows), grouped by class (boundaries in my_cst), when
  // there are <= kRsetCap.
  uint32_t* rset;
  unsigned long long my_cst;  // class boundaries of this lane's own job (see cst_at)
  unsigned inv_mask;  // jobs whose reverse rows fit
  unsigned long long bytes;
  unsigned long long* events;
  bool fatal;
  unsigned lane;
};

template <bool COUNT>
__device__ __forceinline__ void spill_half(WarpCtx<COUNT>& c) {
  constexpr int H = kStackCap / 2;
  if (c.spill_top + H > static_cast<int>(c.spill_cap)) {
    c.fatal = true;
    c.top = 0;
    return;
  }
  for (int i = c.lane; i < H; i += 32) c.spill[c.spill_top + i] = c.stack[i];
  if (c.lane == 0) atomicAdd(c.events, 1ull);
  __syncwarp();
  int rest = c.top - H;  // <= H: read [H, top) and write [0, rest) never overlap
  uint4 t0, t1;
  int i0 = c.lane, i1 = c.lane + 32;
  if (i0 < rest) t0 = c.stack[H + i0];
  if (i1 < rest) t1 = c.stack[H + i1];
  __syncwarp();
  if (i0 < rest) c.stack[i0] = t0;
  if (i1 < rest) c.stack[i1] = t1;
  __syncwarp();
  c.spill_top += H;
  c.top = rest;
}

template <bool COUNT>
__device__ __forceinline__ void refill(WarpCtx<COUNT>& c) {
  constexpr int H = kStackCap / 2;
  int n = c.spill_top < H ? c.spill_top : H;
  for (int i = c.lane; i < n; i += 32) c.stack[i] = c.spill[c.spill_top - n + i];
  __syncwarp();
  c.spill_top -= n;
  c.top = n;
}

// Warp-collective push: lanes with want=true append their item in lane order.
template <bool COUNT>
__device__ __forceinline__ void push(WarpCtx<COUNT>& c, bool want, const uint4& item) {
  unsigned m = __ballot_sync(kFull, want);
  if (!m || c.fatal) return;
  int cnt = __popc(m);
  if (c.top + cnt > kStackCap) spill_half(c);
  if (c.fatal) return;
  if (want) c.stack[c.top + __popc(m & ((1u << c.lane) - 1u))] = item;
  c.top += cnt;
  __syncwarp();
}

// Binary search of the sorted segment arr[lo, hi) for key.
template <bool COUNT>
__device__ __forceinline__ bool find_in(const uint32_t* __restrict__ arr, WarpCtx<COUNT>& c, uint32_t lo, uint32_t hi,
                                        uint32_t key, uint32_t* at = nullptr) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const uint32_t v = __ldg(arr + mid);
    if (COUNT) c.bytes += 4;
    if (v == key) {
      if (at) *at = mid;
      return true;
    }
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

// Binary search of the sorted direct-subject segment col[lo, hi) for sid.
template <bool COUNT>
__device__ __forceinline__ bool probe(const KParams& p, WarpCtx<COUNT>& c, uint32_t lo, uint32_t hi, uint32_t sid,
                                      bool expiry) {
  uint32_t at = 0;
  if (!find_in(p.col, c, lo, hi, sid, &at)) return false;
  if (expiry) {
    const uint32_t e = __ldg(p.exp + at);
    if (COUNT) c.bytes += 4;
    return e == 0 || e > p.now;
  }
  return true;
}

// Two independent binary searches advanced in lock step (two loads in flight per lane).
template <bool COUNT>
__device__ __forceinline__ void probe2(const uint32_t* __restrict__ arr, WarpCtx<COUNT>& c, uint32_t lo0, uint32_t hi0,
                                       uint32_t k0, uint32_t lo1, uint32_t hi1, uint32_t k1, bool& hit0, bool& hit1) {
  while (lo0 < hi0 || lo1 < hi1) {
    const bool a0 = lo0 < hi0, a1 = lo1 < hi1;
    const uint32_t m0 = lo0 + ((hi0 - lo0) >> 1), m1 = lo1 + ((hi1 - lo1) >> 1);
    uint32_t v0 = 0, v1 = 0;
    if (a0) v0 = __ldg(arr + m0);
    if (a1) v1 = __ldg(arr + m1);
    if (COUNT) c.bytes += 4u * (a0 + a1);
    if (a0) {
      if (v0 == k0) { hit0 = true; lo0 = hi0; }
      else if (v0 < k0) lo0 = m0 + 1; else hi0 = m0;
    }
    if (a1) {
      if (v1 == k1) { hit1 = true; lo1 = hi1; }
      else if (v1 < k1) lo1 = m1 + 1; else hi1 = m1;
    }
  }
}

// Do the sorted segments a[a0, a1) and b[b0, b1) share an element? Sizes within 4x of each other:
// two-pointer merge (one load per step); otherwise every element of the short one is searched in
// the long one.
template <bool COUNT>
__device__ __forceinline__ bool intersects(const uint32_t* __restrict__ a, uint32_t a0, uint32_t a1,
                                           const uint32_t* __restrict__ b, uint32_t b0, uint32_t b1, WarpCtx<COUNT>& c) {
  const uint32_t na = a1 - a0, nb = b1 - b0;
  if (na == 0 || nb == 0) return false;
  if (na > 4u * nb) {
    for (uint32_t x = b0; x < b1; ++x) {
      const uint32_t k = __ldg(b + x);
      if (COUNT) c.bytes += 4;
      if (find_in(a, c, a0, a1, k)) return true;
    }
    return false;
  }
  if (nb > 4u * na) {
    for (uint32_t x = a0; x < a1; ++x) {
      const uint32_t k = __ldg(a + x);
      if (COUNT) c.bytes += 4;
      if (find_in(b, c, b0, b1, k)) return true;
    }
    return false;
  }
  uint32_t va = __ldg(a + a0), vb = __ldg(b + b0);
  if (COUNT) c.bytes += 8;
  for (;;) {
    if (va == vb) return true;
    if (va < vb) {
      if (++a0 == a1) return false;
      va = __ldg(a + a0);
    } else {
      if (++b0 == b1) return false;
      vb = __ldg(b + b0);
    }
    if (COUNT) c.bytes += 4;
  }
}

// Boundaries of a job's reverse-row set per invertible class of its subject type:
// 5 bits each (0..31), boundary i at bit 5*i; class i owns entries [b_i, b_{i+1}).
__device__ __forceinline__ uint32_t cst_at(unsigned long long cst, uint32_t i) {
  return static_cast<uint32_t>(cst >> (5u * i)) & 31u;
}
constexpr int kMaxInvClasses = 11;
// bit of (invertible class index, object) in a check's 32-bit presence word
__device__ __forceinline__ uint32_t rset_bit(uint32_t tinv, uint32_t obj) {
  return ((obj + tinv * 0x632BE5ABu) * 0x9E3779B1u) >> 27;
}
static_assert(kRsetCap <= 31, "class boundaries are 5-bit fields");

// Two-level meet, warp-cooperative: ONE check's range against the reverse rows of its subject's memberships, all 32
// lanes working on it. The per-lane form of this (every lane walking its own check's rows) ran with 9 of 32 lanes
// active on average (profiles/r2d): row counts and early exits differ per check. Here the range's children F
// (<= kFCap, ascending) go to shared memory with one coalesced load, every membership row gets a fixed group of lanes
// (neighbouring lanes read neighbouring words of the row), and each lane tests its element against a Bloom filter of F;
// F itself is binary-searched (warp-uniform trip count) only for what the filter lets through. Returns (warp-uniform) whether some element of some row is a child of the range.
// Deliberately NOT inlined: inlined, its registers take part in the allocation of the whole traversal loop and cost
// schemas that never use it 10 % (cfg4: 1 115 -> 944 Mchecks/s, profiles/r2e-r2g); a hash table of the range instead
// of the binary search was slower still (864 vs 1 094 on cfg3, r2g).
// (all arguments by value: a reference to the warp context would force it into local memory)
template <bool COUNT>
__device__ __noinline__ bool coop_l2(const uint32_t* __restrict__ col, const uint32_t* __restrict__ rrow_ptr,
                                     const uint32_t* __restrict__ rcol, uint32_t* rset, uint32_t lane, uint32_t lo, uint32_t hi,
                                     uint32_t kb, uint32_t ke, uint32_t jslot, unsigned long long rrow_base, uint32_t rstride,
                                     uint32_t nsubj, unsigned long long* bytes) {
  uint32_t* const fs = rset + (kRsetCap + kStateWords) * 32;
  const uint32_t nf = hi - lo, ng = ke - kb;  // nf <= kFCap (64), 1 <= ng <= kRsetCap (16)
#if ZG_L2_MATCH
  if (nf <= 16u) {
    // Short range (the common case): its children sit in lanes 0-15, sixteen row elements at a time in lanes 16-31,
    // and ONE match.any tells every element lane whether an F lane holds the same value: no search at all.
    // Unused lanes hold values no object id can take (ids stay below 0xFFFFFFF0), all different.
    const uint32_t pad = 0xFFFFFFE0u + lane;
    const uint32_t fval = lane < nf ? __ldg(col + lo + lane) : pad;
    const uint32_t el = lane - 16u;  // element lane index (lanes 16-31)
    const uint32_t sh = ng <= 1 ? 4u : (ng <= 2 ? 3u : (ng <= 4 ? 2u : (ng <= 8 ? 1u : 0u)));  // log2(lanes per row)
    const uint32_t w = 1u << sh, row = el >> sh, k = el & (w - 1u);
    uint32_t x = 0, h = 0;
    if (lane >= 16u && row < ng) {
      const uint32_t g = rset[(kb + row) * 32 + jslot];
      if (g < nsubj) {
        const unsigned long long ri = rrow_base + static_cast<unsigned long long>(g) * rstride;
        x = __ldg(rrow_ptr + ri) + k;
        h = __ldg(rrow_ptr + ri + 1);
        if (COUNT && k == 0) *bytes += 8ull + 4ull * (h - (x - k));
      }
    }
    if (COUNT && lane == 0) *bytes += 4ull * nf;
    while (__any_sync(kFull, x < h)) {
      uint32_t v = fval;
      if (lane >= 16u) {
        v = pad;
        if (x < h) {
          v = __ldg(rcol + x);
          x += w;
        }
      }
      const unsigned same = __match_any_sync(kFull, v);
      if (__any_sync(kFull, lane >= 16u && (same & 0xFFFFu))) return true;
    }
    return false;
  }
#endif
#if ZG_L2_FILTER
  // filter word = top 6 bits of the hash, two bits inside it from the next 2 x 5 bits
  uint32_t* const fb = fs + kFCap;
  fb[lane] = 0;
  fb[lane + 32] = 0;
  __syncwarp();
  if (lane < nf) {
    const uint32_t v = __ldg(col + lo + lane), hv = v * 0x9E3779B1u;
    fs[lane] = v;
    atomicOr(&fb[hv >> 26], (1u << ((hv >> 21) & 31u)) | (1u << ((hv >> 16) & 31u)));
  }
  if (lane + 32 < nf) {
    const uint32_t v = __ldg(col + lo + lane + 32), hv = v * 0x9E3779B1u;
    fs[lane + 32] = v;
    atomicOr(&fb[hv >> 26], (1u << ((hv >> 21) & 31u)) | (1u << ((hv >> 16) & 31u)));
  }
#else
  if (lane < nf) fs[lane] = __ldg(col + lo + lane);
  if (lane + 32 < nf) fs[lane + 32] = __ldg(col + lo + lane + 32);
#endif
  // W lanes per membership row: lane = row * W + k walks elements k, k + W, ... of its row. No prefix sums, no owner
  // search: the lanes of a row read the same two offsets (one broadcast load).
#if ZG_L2_SPLIT
  uint32_t w, magic;  // W = 32 / ng; row = lane / W as (lane * ceil(256 / W)) >> 8, exact for lane < 32
  if (ng <= 2) { w = 16; magic = 16; }
  else if (ng == 3) { w = 10; magic = 26; }
  else if (ng == 4) { w = 8; magic = 32; }
  else if (ng == 5) { w = 6; magic = 43; }
  else if (ng == 6) { w = 5; magic = 52; }
  else if (ng <= 8) { w = 4; magic = 64; }
  else if (ng <= 10) { w = 3; magic = 86; }
  else { w = 2; magic = 128; }
  const uint32_t row = (lane * magic) >> 8, k = lane - row * w;
#else
  const uint32_t sh = ng <= 2 ? 4u : (ng <= 4 ? 3u : (ng <= 8 ? 2u : 1u));  // log2(W): 32 / next power of two of ng
  const uint32_t w = 1u << sh, row = lane >> sh, k = lane & (w - 1u);
#endif
  uint32_t x = 0, h = 0;
  if (row < ng) {
    const uint32_t g = rset[(kb + row) * 32 + jslot];
    if (g < nsubj) {
      const unsigned long long ri = rrow_base + static_cast<unsigned long long>(g) * rstride;
      x = __ldg(rrow_ptr + ri) + k;
      h = __ldg(rrow_ptr + ri + 1);
      if (COUNT && k == 0) *bytes += 8ull + 4ull * (h - (x - k));
    }
  }
  if (COUNT && lane == 0) *bytes += 4ull * nf;
  __syncwarp();
  const uint32_t span = nf > 1 ? 1u << (32 - __clz(nf - 1)) : 1u;  // smallest power of two >= nf
  while (__any_sync(kFull, x < h)) {
    bool found = false;
#if ZG_L2_FILTER
    uint32_t t = 0;
    if (x < h) {
      t = __ldg(rcol + x);
      const uint32_t ht = t * 0x9E3779B1u, m = (1u << ((ht >> 21) & 31u)) | (1u << ((ht >> 16) & 31u));
      found = (fb[ht >> 26] & m) == m;  // "maybe": verified below
      x += w;
    }
    if (__any_sync(kFull, found)) {
      uint32_t pos = 0;  // lower bound of t in fs[0, nf): warp-uniform trip count
      for (uint32_t s = span; s >= 1; s >>= 1)
        if (pos + s <= nf && fs[pos + s - 1] < t) pos += s;
      found = found && pos < nf && fs[pos] == t;
      if (__any_sync(kFull, found)) return true;
    }
#else
    if (x < h) {
      const uint32_t t = __ldg(rcol + x);
      uint32_t pos = 0;  // lower bound of t in fs[0, nf): warp-uniform trip count
      for (uint32_t s = span; s >= 1; s >>= 1)
        if (pos + s <= nf && fs[pos + s - 1] < t) pos += s;
      found = pos < nf && fs[pos] == t;
      x += w;
    }
    if (__any_sync(kFull, found)) return true;
#endif
  }
  return false;
}

// Warp-collective node visit: every lane may carry one (job slot, object, unit).
// Evaluates the unit's steps at the object: membership-of-itself test, direct and
// wildcard probes (from the subject's reverse-row set when the job is inverted, else a
// binary search of the row), and pushes the userset / arrow edge ranges.
template <bool COUNT>
__device__ __forceinline__ void visit(const KParams& p, const Prog& pr, WarpCtx<COUNT>& c, bool active,
                                      uint32_t jslot, uint32_t obj, uint32_t unit, uint32_t depth) {
  const uint32_t sid = __shfl_sync(kFull, c.my_subj, jslot & 31);
  const uint32_t ss = __shfl_sync(kFull, c.my_ss, jslot & 31);
  const unsigned long long cst = __shfl_sync(kFull, c.my_cst, jslot & 31);
  const uint32_t stype = ss >> 16, srel = ss & 0xFFFFu;
  bool hit = false;
  int sb = 0, nprobe = 0, npush = 0;
  if (active) {
    const DUnit u = pr.units()[unit];
    if (srel != kNone && sid == obj)
      for (int m = u.mem_begin; m < u.mem_end; ++m) hit = hit || pr.members()[m] == srel;
    if (!hit) {
      sb = u.step_begin;
      nprobe = u.push_begin - u.step_begin;
      npush = u.step_end - u.push_begin;
    }
  }
  const bool inverted = (c.inv_mask >> (jslot & 31)) & 1u;
  // ---- probes (DIRECT / WILD classes): every lane on its own, no warp collective involved. The steps of a unit
  // are ordered probes first (schema.cc); in round 1 every step, probe or not, went through the collective push.
  for (int i = 0; i < nprobe && !hit; ++i) {
    const DStep st = pr.steps()[sb + i];
    const bool expiry = (st.flags & CF_EXPIRY) != 0;
    const bool subject_fits = srel == kNone && stype == st.stype;
    if (!subject_fits) continue;
    if (st.kind == ST_DIRECT && inverted && (st.flags & CF_INVERT)) {
      // direction-optimised probe: is obj among the subject's memberships of this class? Most probes miss: a
      // 32-bit presence word over (class, object) of the check's set (built at admission) rejects them
      // without walking the set (the walk was 15 % of cfg4's instructions, profiles/r2d).
      const uint32_t present = c.rset[(kRsetCap + 5) * 32 + (jslot & 31)];
      if (!ZG_PRESENCE || ((present >> rset_bit(st.tinv, obj)) & 1u)) {
        const uint32_t kb = cst_at(cst, st.tinv), ke = cst_at(cst, st.tinv + 1u);
        for (uint32_t r = kb; r < ke; ++r)
          hit = hit || c.rset[r * 32 + (jslot & 31)] == obj;
      }
    } else if (obj < st.nres) {
      const unsigned long long ridx = st.row_base + static_cast<unsigned long long>(obj) * st.ncls;
      const uint32_t lo = __ldg(p.row_ptr + ridx), hi = __ldg(p.row_ptr + ridx + 1);
      if (COUNT) c.bytes += 8;
      if (hi > lo) {
        if (st.kind == ST_DIRECT) {
          hit = probe(p, c, lo, hi, sid, expiry);
        } else if (expiry) {  // ST_WILD
          const uint32_t e = __ldg(p.exp + lo);
          if (COUNT) c.bytes += 4;
          hit = e == 0 || e > p.now;
        } else {
          hit = true;
        }
      }
    }
  }
  if (hit) npush = 0;
  sb += nprobe;
  // ---- edge classes that may push a range (userset subjects, arrows): warp-collective
  const int maxsteps = __reduce_max_sync(kFull, npush);
  for (int i = 0; i < maxsteps; ++i) {
    bool want = false, l2_want = false;
    uint32_t l2_lo = 0, l2_hi = 0, l2_k = 0;
    uint4 item = make_uint4(0, 0, 0, 0);
    if (i < npush && !hit) {
      const DStep st = pr.steps()[sb + i];
      const bool expiry = (st.flags & CF_EXPIRY) != 0;
      if (obj < st.nres) {
        const unsigned long long ridx = st.row_base + static_cast<unsigned long long>(obj) * st.ncls;
        const uint32_t lo = __ldg(p.row_ptr + ridx), hi = __ldg(p.row_ptr + ridx + 1);
        if (COUNT) c.bytes += 8;
        if (hi > lo) {
          if ((st.flags & kStepTargetLeaf) && inverted && !expiry && depth + 1 <= ZG_MAX_DEPTH) {
            // Children of this range can only be answered by "is the child one of the
            // subject's memberships of class tinv": meet in the middle. No membership (or a
            // child class for another subject type): nothing can match.
            const uint32_t kb = cst_at(cst, st.tinv), ke = stype == st.tstype ? cst_at(cst, st.tinv + 1u) : kb;
            if (ke - kb > hi - lo) {  // fewer children than memberships: visit the children
              want = true;
              item = make_uint4(lo, hi, make_meta(jslot, depth + 1, false, st.tslot), 0);
            } else {
              // For each membership key ask ITS reverse row (short, and shared by every range
              // of this check, so it stays in L1): "does this object list the key as a subject
              // of class gc". Two searches are advanced together for memory-level parallelism.
              const DCls cl = pr.cls()[st.gc];
              for (uint32_t r = kb; r < ke && !hit; r += 2) {
                uint32_t l0 = 0, h0 = 0, l1 = 0, h1 = 0;
                const uint32_t m0 = c.rset[r * 32 + (jslot & 31)];
                const uint32_t m1 = r + 1 < ke ? c.rset[(r + 1) * 32 + (jslot & 31)] : 0xFFFFFFFFu;
                if (m0 < cl.nsubj) {
                  const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(m0) * cl.rstride;
                  l0 = __ldg(p.rrow_ptr + ri);
                  h0 = __ldg(p.rrow_ptr + ri + 1);
                  if (COUNT) c.bytes += 8;
                }
                if (m1 < cl.nsubj) {
                  const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(m1) * cl.rstride;
                  l1 = __ldg(p.rrow_ptr + ri);
                  h1 = __ldg(p.rrow_ptr + ri + 1);
                  if (COUNT) c.bytes += 8;
                }
                bool hit0 = false, hit1 = false;
                probe2(p.rcol, c, l0, h0, obj, l1, h1, obj, hit0, hit1);
                hit = hit0 || hit1;
              }
            }
          } else if ((st.flags & kStepTargetL2) && inverted && depth + 2 <= ZG_MAX_DEPTH) {
            // Two levels at once (namespace -> team#member -> group#member -> user): a child t of this
            // range matches iff some membership g of the subject (class tinv) is a subject of t in the
            // children's single class tgc. The reverse row of g in that class lists exactly those t, in
            // ascending order like this range: the answer is "do the two sorted segments intersect",
            // once per membership. The children are never visited, nothing is pushed.
            const uint32_t kb = cst_at(cst, st.tinv), ke = stype == st.tstype ? cst_at(cst, st.tinv + 1u) : kb;
            const DCls cl = pr.cls()[st.tgc];
            if (ZG_L2_MODE == 2 && ke > kb && hi - lo <= static_cast<uint32_t>(kFCap)) {
              // handed to the whole warp below (coop_l2), one requesting lane at a time
              l2_want = true;
              l2_lo = lo;
              l2_hi = hi;
              l2_k = kb | (ke << 8) | (static_cast<uint32_t>(st.tgc) << 16);
            } else
            if (ZG_L2_MODE == 1 && ke > kb && hi - lo <= 256u) {
              // Short range (the common case): a 64-bit Bloom word of its children stays in a register and the
              // reverse rows are STREAMED against it with 128-bit loads -- independent loads, no dependent
              // chain per element; only a Bloom hit is verified by a binary search of the range.
              unsigned long long bloom = 0;
              for (uint32_t x = lo & ~3u; x < hi; x += 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.col + x));
                if (x >= lo) bloom |= 1ull << ((v.x * 0x9E3779B1u) >> 26);
                if (x + 1 >= lo && x + 1 < hi) bloom |= 1ull << ((v.y * 0x9E3779B1u) >> 26);
                if (x + 2 >= lo && x + 2 < hi) bloom |= 1ull << ((v.z * 0x9E3779B1u) >> 26);
                if (x + 3 >= lo && x + 3 < hi) bloom |= 1ull << ((v.w * 0x9E3779B1u) >> 26);
              }
              if (COUNT) c.bytes += 4ull * (hi - lo);  // the words of the range (the padding of a 128-bit load is not work)
              for (uint32_t r = kb; r < ke && !hit; ++r) {
                const uint32_t g = c.rset[r * 32 + (jslot & 31)];
                if (g >= cl.nsubj) continue;
                const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(g) * cl.rstride;
                const uint32_t l = __ldg(p.rrow_ptr + ri), h = __ldg(p.rrow_ptr + ri + 1);
                if (COUNT) c.bytes += 8 + 4ull * (h - l);
                for (uint32_t x = l & ~3u; x < h && !hit; x += 4) {
                  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.rcol + x));
                  const uint32_t e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (x + k >= l && x + k < h && ((bloom >> ((e4[k] * 0x9E3779B1u) >> 26)) & 1ull))
                      hit = hit || find_in(p.col, c, lo, hi, e4[k]);
                }
              }
            } else {
              for (uint32_t r = kb; r < ke && !hit; ++r) {
                const uint32_t g = c.rset[r * 32 + (jslot & 31)];
                if (g >= cl.nsubj) continue;
                const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(g) * cl.rstride;
                const uint32_t l = __ldg(p.rrow_ptr + ri), h = __ldg(p.rrow_ptr + ri + 1);
                if (COUNT) c.bytes += 8;
                hit = intersects(p.col, lo, hi, p.rcol, l, h, c);
              }
            }
          } else {
            want = true;
            item = make_uint4(lo, hi, make_meta(jslot, depth + 1, expiry, st.tslot), 0);
          }
        }
      }
    }
    if (ZG_L2_MODE == 2) {
      unsigned l2m = __ballot_sync(kFull, l2_want);
      while (l2m) {  // warp-uniform: one requesting lane after the other, the whole warp on its range
        const int r = __ffs(l2m) - 1;
        l2m &= l2m - 1;
        const uint32_t rlo = __shfl_sync(kFull, l2_lo, r), rhi = __shfl_sync(kFull, l2_hi, r), rk = __shfl_sync(kFull, l2_k, r);
        const uint32_t rj = __shfl_sync(kFull, jslot, r) & 31u;
        if ((c.found >> rj) & 1u) continue;  // another node of the same check already answered it
        const DCls rcl = pr.cls()[rk >> 16];
        unsigned long long cb = 0;
        const bool h = coop_l2<COUNT>(p.col, p.rrow_ptr, p.rcol, c.rset, c.lane, rlo, rhi, rk & 0xFFu, (rk >> 8) & 0xFFu, rj,
                                      rcl.rrow_base, rcl.rstride, rcl.nsubj, &cb);
        if (COUNT) c.bytes += cb;
        if (h) c.found |= 1u << rj;
        __syncwarp();
      }
    }
    push(c, want, item);
  }
  c.found |= __reduce_or_sync(kFull, hit ? (1u << jslot) : 0u);
}

__device__ __forceinline__ uint32_t kleene_or(uint32_t a, uint32_t b) {  // 0 F, 1 T, 2 E
  return (a == 1 || b == 1) ? 1u : ((a == 2 || b == 2) ? 2u : 0u);
}
__device__ __forceinline__ uint32_t kleene_and(uint32_t a, uint32_t b) {
  return (a == 0 || b == 0) ? 0u : ((a == 2 || b == 2) ? 2u : 1u);
}

// Postfix boolean program of a non-pure permission over its leaf values (2 bits per leaf in
// `vals`: 0 F, 1 T, 2 E). Kleene's E doubles as "not evaluated yet": the connectives are monotone,
// so a result that is not E cannot change whatever the missing leaves turn out to be.
__device__ __forceinline__ uint32_t eval_tree(const Prog& pr, const DTree& t, unsigned long long vals, bool trivial_self,
                                              uint32_t srel) {
  unsigned long long st = 0;  // 2-bit entries
  int sp = 0;
  for (int i = t.op_begin; i < t.op_end; ++i) {
    const DTreeOp o = pr.tree_ops()[i];
    if (o.kind == T_LEAF) {
      st |= ((vals >> (2u * o.arg)) & 3ull) << (2 * sp);
      ++sp;
    } else if (o.kind == T_TRIVIAL) {
      const unsigned long long v = (trivial_self && srel == o.arg) ? 1ull : 0ull;
      st |= v << (2 * sp);
      ++sp;
    } else {
      const uint32_t b = (st >> (2 * (sp - 1))) & 3u, a = (st >> (2 * (sp - 2))) & 3u;
      uint32_t v;
      if (o.kind == T_OR) v = kleene_or(a, b);
      else if (o.kind == T_AND) v = kleene_and(a, b);
      else v = kleene_and(a, b == 1 ? 0u : (b == 0 ? 1u : 2u));
      sp -= 2;
      st &= ~(0xFull << (2 * sp));
      st |= static_cast<unsigned long long>(v) << (2 * sp);
      ++sp;
    }
  }
  return static_cast<uint32_t>(st & 3u);
}

// Streamed admission: spin until `need` leading queries have landed. If the feed stopped (a failed copy) give
// up after seconds instead of hanging the device: the flag fails the host call, whatever is answered from
// there on is discarded.
__device__ __noinline__ void wait_ready(const unsigned long long* ready, unsigned long long need, uint32_t* flags) {
  const volatile unsigned long long* rd = ready;
  uint32_t spins = 0;
  while (*rd < need && spins < (1u << 23)) {
    __nanosleep(256);
    ++spins;
  }
  if (spins >= (1u << 23)) atomicOr(flags, 32u);
  // CTA-scope fence: orders the item loads after the poll without the L1 invalidate a device-scope fence
  // costs (the items are read with ld.global.cg, which does not use L1)
  __threadfence_block();
}

// Per-lane query state parked in shared memory while the warp traverses (kStateWords words):
//   0 resource object, 1 depth | leaves << 8 | is-tree << 15 | leaf_begin-or-unit << 16,
//   2 tree id, 3 / 4 leaf values (2 bits each), 5 presence word of the reverse-row set
template <bool COUNT, bool STREAMED = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks) check_kernel(const KParams p) {
  ZG_DYNAMIC_SMEM(smem);
  for (uint32_t i = threadIdx.x; i < p.prog_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(p.prog)[i];
  __syncthreads();
  const Prog pr = make_prog(smem);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpCtx<COUNT> c;
  c.lane = lane;
  uint8_t* wsm = smem + p.prog_bytes + warp * kWarpSmem;
  c.stack = reinterpret_cast<uint4*>(wsm);
  c.rset = reinterpret_cast<uint32_t*>(wsm + kStackCap * sizeof(uint4));
  uint32_t* const state = c.rset + kRsetCap * 32 + lane;  // word k at state[k * 32]
  c.spill = p.spill + (static_cast<size_t>(blockIdx.x) * kWarpsPerBlock + warp) * p.spill_cap;
  c.spill_cap = p.spill_cap;
  c.bytes = 0;
  c.events = p.events;
  const uint32_t n_slots = pr.hdr()->n_slots, n_types = pr.hdr()->n_types;
  unsigned long long* const memo = p.memo + (static_cast<size_t>(blockIdx.x) * kWarpsPerBlock + warp) * p.memo_entries;

  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) {
      base = atomicAdd(p.next, 32ull);
      // the copy engine is still feeding the queries: wait until this batch has landed
      if (STREAMED && base < p.nq) wait_ready(p.ready, base + 32 < p.nq ? base + 32 : p.nq, p.flags);
    }
    base = __shfl_sync(kFull, base, 0);
    if (base >= p.nq) break;
    const unsigned long long q = base + lane;
    const bool valid = q < p.nq;
    bool bad = false, expansive = false;
    uint32_t nl = 0;
    c.my_subj = 0;
    c.my_ss = (0u << 16) | kNone;
    if (valid) {
      // ld.global.cg when the items are streamed in: nothing about them may come from a non-coherent cache
      const uint4 raw = STREAMED ? __ldcg(reinterpret_cast<const uint4*>(p.queries) + q)
                                 : __ldg(reinterpret_cast<const uint4*>(p.queries) + q);
      if (COUNT) c.bytes += 17;
      c.my_subj = raw.y;
      const uint32_t perm = raw.z & 0xFFFFu, stype = raw.z >> 16;
      const uint32_t srel = raw.w & 0xFFFFu, fl = p.raw_items ? 0u : raw.w >> 16;
      c.my_ss = (stype << 16) | srel;
      uint32_t w1 = fl & kJobDepthMask, tree = 0;
      if (stype >= n_types || (srel != kNone && (srel >= n_slots || pr.slots()[srel].type != stype))) bad = true;
      if (perm >= n_slots) {
        bad = true;
      } else {
        const DSlot s = pr.slots()[perm];
        if (s.kind == SK_NONPURE) {
          const DTree t = pr.trees()[s.unit];
          tree = s.unit;
          nl = t.n_leaves;
          expansive = (t.flags & UF_EXPANSIVE) != 0;
          w1 |= (nl << 8) | (1u << 15) | (static_cast<uint32_t>(t.leaf_begin) << 16);
        } else {
          nl = 1;
          expansive = (pr.units()[s.unit].flags & UF_EXPANSIVE) != 0;
          w1 |= (1u << 8) | (static_cast<uint32_t>(s.unit) << 16);
        }
      }
      state[0] = raw.x;
      state[32] = w1;
      state[64] = tree;
      state[96] = 0xAAAAAAAAu;  // every leaf: E = not evaluated yet
      state[128] = 0xAAAAAAAAu;
    }
    // ---- direction-optimised probes: load the subject's reverse rows (its direct
    // memberships, class by class) when the check can fan out and they are few. Loaded once per
    // query: every leaf of a non-pure permission shares them.
    {
      bool inv = p.invert && valid && !bad && (c.my_ss & 0xFFFFu) == kNone && expansive;
      uint32_t rcnt = 0, present = 0;
      unsigned long long cst = 0;
      int ib = 0, ncl = 0;
      if (inv) {
        const DTypeInv ti = pr.type_inv()[c.my_ss >> 16];
        ib = ti.begin;
        ncl = ti.end - ti.begin;
        if (ncl > kMaxInvClasses) {
          inv = false;
          ncl = 0;
        }
      }
      const int maxcl = __reduce_max_sync(kFull, ncl);
      for (int i = 0; i < maxcl; ++i) {
        if (inv && i < ncl) {
          const uint32_t gc = pr.inv_cls()[ib + i];
          const DCls cl = pr.cls()[gc];
          if (!(cl.flags & CF_EMPTY) && c.my_subj < cl.nsubj) {
            const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(c.my_subj) * cl.rstride;
            const uint32_t b = __ldg(p.rrow_ptr + ri);
            const uint32_t e = __ldg(p.rrow_ptr + ri + 1);
            if (COUNT) c.bytes += 8;
            if (rcnt + (e - b) > static_cast<uint32_t>(kRsetCap)) {
              inv = false;  // too many memberships: this check probes forward
            } else {
              for (uint32_t x = b; x < e; ++x) {
                const uint32_t m = __ldg(p.rcol + x);
                c.rset[rcnt * 32 + lane] = m;
                present |= 1u << rset_bit(static_cast<uint32_t>(i), m);
                ++rcnt;
              }
              if (COUNT) c.bytes += 4ull * (e - b);
            }
          }
          cst |= static_cast<unsigned long long>(rcnt) << (5u * (i + 1));
        }
      }
      c.inv_mask = __ballot_sync(kFull, inv);
      c.my_cst = cst;
      state[160] = present;
      __syncwarp();
    }
    c.fatal = false;
    bool undet = valid && !bad;
    uint32_t res = 2;  // invalid query -> error
    const int maxnl = __reduce_max_sync(kFull, undet ? static_cast<int>(nl) : 0);
    for (int l = 0; l < maxnl; ++l) {
      const bool act = undet && l < static_cast<int>(nl);
      const unsigned live_jobs = __ballot_sync(kFull, act);
      if (!live_jobs) continue;
      uint32_t obj = 0, unit = 0, depth = 0;
      if (act) {
        obj = state[0];
        const uint32_t w1 = state[32];
        depth = w1 & kJobDepthMask;
        unit = (w1 >> 15) & 1u ? pr.leaf_units()[(w1 >> 16) + l] : (w1 >> 16);
      }
      c.found = 0;
      c.err = 0;
      c.raised = 0;
      c.top = 0;
      c.spill_top = 0;

      visit(p, pr, c, act, lane, obj, unit, depth);

      uint32_t iters = 0;
      bool use_memo = false;
      for (;;) {
        if (!c.fatal && c.top == 0) {
          if (c.spill_top == 0) break;
          refill(c);
        }
        if ((c.found & live_jobs) == live_jobs) break;  // every check already answered
        if (++iters > p.budget || c.fatal) {
          if (lane == 0) atomicOr(p.flags, c.fatal ? 1u : 4u);
          c.err |= live_jobs & ~c.found;
          break;
        }
        if (!use_memo && p.memo_entries && iters > p.memo_after) {
          // this batch is expanding far more than usual: path multiplicity. Remember child
          // visits from here on; an identical (job, slot, object, depth) visit is skipped.
          for (uint32_t i = lane; i < p.memo_entries; i += 32) memo[i] = 0ull;
          if (lane == 0) atomicAdd(p.events + 1, 1ull);
          __syncwarp();
          use_memo = true;
        }
        // ---- pop ranges worth <= 32 edges from the top of the stack
        const int n = c.top < 32 ? c.top : 32;
        uint4 it = make_uint4(0, 0, 0, 0);
        if (static_cast<int>(lane) < n) it = c.stack[c.top - 1 - lane];
        const bool dead = (c.found >> (it.z & 31)) & 1u;
        const uint32_t len = (static_cast<int>(lane) < n && !dead) ? it.y - it.x : 0u;
        uint32_t total, jb, jmeta, jexcl;
        if (ZG_POP_FAST && !__ballot_sync(kFull, len > 1u)) {
          // Every popped range holds at most one edge (arrows to a single parent, singleton usersets: the common
          // case on document / folder hierarchies): no prefix scan and no owner search -- lane k takes the k-th
          // live item, found with one ballot and one find-nth-set.
          const unsigned ones = __ballot_sync(kFull, len == 1u);
          total = static_cast<uint32_t>(__popc(ones));
          c.top -= n;
          const int j = lane < total ? static_cast<int>(__fns(ones, 0, static_cast<int>(lane) + 1)) : 0;
          jb = __shfl_sync(kFull, it.x, j & 31);
          jmeta = __shfl_sync(kFull, it.z, j & 31);
          jexcl = lane;
          __syncwarp();
        } else {
          uint32_t incl = len;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(kFull, incl, d);
            if (static_cast<int>(lane) >= d) incl += v;
          }
          const uint32_t excl = incl - len;
          const unsigned fullm = __ballot_sync(kFull, static_cast<int>(lane) < n && incl <= 32u);
          const int nfull = __popc(fullm);  // prefix of fully consumed items
          total = __shfl_sync(kFull, incl, 31);
          if (total > 32u) total = 32u;
          __syncwarp();
          if (static_cast<int>(lane) == nfull && static_cast<int>(lane) < n)  // partially consumed item stays on top
            c.stack[c.top - 1 - lane].x = it.x + (32u - excl);
          c.top -= nfull;
          __syncwarp();
          // ---- lane k takes edge k: owner item j = #items with incl <= k
          int j = 0;
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int probe_lane = j + step - 1;
            const uint32_t v = __shfl_sync(kFull, incl, probe_lane & 31);
            if (probe_lane < 32 && v <= lane) j += step;
          }
          jb = __shfl_sync(kFull, it.x, j & 31);
          jmeta = __shfl_sync(kFull, it.z, j & 31);
          jexcl = __shfl_sync(kFull, excl, j & 31);
        }
        bool active = lane < total;
        const uint32_t jslot = jmeta & 31u, cdepth = (jmeta >> 5) & 63u, tslot = jmeta >> 16;
        uint32_t child = 0;
        const uint32_t sq_subj = __shfl_sync(kFull, c.my_subj, jslot);
        const uint32_t sq_ss = __shfl_sync(kFull, c.my_ss, jslot);
        if (active) {
          const uint32_t e = jb + (lane - jexcl);
          child = __ldg(p.col + e);
          if (COUNT) c.bytes += 4;
          if (jmeta & (1u << 11)) {
            const uint32_t ex = __ldg(p.exp + e);
            if (COUNT) c.bytes += 4;
            active = ex == 0 || ex > p.now;
          }
        }
        // the 51st hop is an error (pkg/spicedb/spicedb.go:33)
        const bool too_deep = active && cdepth > ZG_MAX_DEPTH;
        c.err |= __reduce_or_sync(kFull, too_deep ? (1u << jslot) : 0u);
        active = active && !too_deep;
        if (use_memo && active) {
          // exact key (a visit at another depth has another hop budget, so depth is part of it);
          // direct-mapped and lossy: a lost entry only costs a repeated visit
          const unsigned long long key = (1ull << 59) | (static_cast<unsigned long long>(jslot) << 54) |
                                         (static_cast<unsigned long long>(cdepth) << 48) |
                                         (static_cast<unsigned long long>(tslot) << 32) | child;
          unsigned long long hsh = key * 0x9E3779B97F4A7C15ull;
          const uint32_t at = static_cast<uint32_t>(hsh >> 40) & (p.memo_entries - 1u);
          if (memo[at] == key) active = false;
          else memo[at] = key;
        }
        uint32_t cunit = 0;
        bool raise = false;
        if (active) {
          const DSlot s = pr.slots()[tslot];
          const bool remote = p.shard_count > 1 && (child % p.shard_count) != p.shard_rank;
          if (s.kind == SK_NONPURE || remote) {
            // defer: Check(child#tslot @ S) becomes a sub-query of the next pass
            const unsigned long long at = atomicAdd(p.subq_count, 1ull);
            if (at < p.subq_cap) {
              zg_check sq;
              sq.res = child;
              sq.subj = sq_subj;
              sq.perm = static_cast<uint16_t>(tslot);
              sq.stype = static_cast<uint16_t>(sq_ss >> 16);
              sq.srel = static_cast<uint16_t>(sq_ss & 0xFFFFu);
              sq.flags = static_cast<uint16_t>(cdepth);
              p.subq[at] = sq;
              p.subq_parent[at] = static_cast<uint32_t>((base + jslot) * p.L + l);
            } else {
              atomicOr(p.flags, 2u);
            }
            raise = true;
            active = false;
          } else {
            cunit = s.unit;
          }
        }
        c.raised |= __reduce_or_sync(kFull, raise ? (1u << jslot) : 0u);
        visit(p, pr, c, active, jslot, child, cunit, cdepth);
      }

      if (act) {
        const bool t = (c.found >> lane) & 1u, e = (c.err >> lane) & 1u, r = (c.raised >> lane) & 1u;
        if (p.val) p.val[q * p.L + l] = t ? kValT : (e ? kValE : 0);
        // a leaf that raised sub-queries is not final unless it is already true
        const uint32_t v = t ? 1u : ((e || r) ? 2u : 0u);
        const uint32_t w1 = state[32];
        if ((w1 >> 15) & 1u) {
          unsigned long long vals = (static_cast<unsigned long long>(state[128]) << 32) | state[96];
          vals = (vals & ~(3ull << (2 * l))) | (static_cast<unsigned long long>(v) << (2 * l));
          state[96] = static_cast<uint32_t>(vals);
          state[128] = static_cast<uint32_t>(vals >> 32);
          const uint32_t srel = c.my_ss & 0xFFFFu;
          res = eval_tree(pr, pr.trees()[state[64]], vals, c.my_subj == state[0], srel);
          if (res != 2u || l + 1 == static_cast<int>(nl)) undet = false;
        } else {
          res = v;
          undet = false;
        }
      }
    }
    if (valid && p.out) {
      if (p.final_codes) p.out[q] = res == 1 ? ZG_HAS_PERMISSION : (res == 2 ? ZG_ITEM_ERROR : ZG_NO_PERMISSION);
      else p.out[q] = res == 1 ? kValT : (res == 2 ? kValE : 0);
    }
    __syncwarp();
  }
  if (COUNT) {
    unsigned long long b = c.bytes;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) b += __shfl_xor_sync(kFull, b, d);
    if (lane == 0 && b) atomicAdd(p.alg_bytes, b);
  }
}

// ---- multi-pass runs: the boolean fold of a level whose queries raised sub-queries -------------

// Evaluates each query's boolean tree over its leaf values val[q * L + l] (the check kernel wrote
// them, the folds of deeper levels OR-ed the sub-query results in). out != null: write the v1 code
// (or the raw value bits) to out[q]; otherwise OR the value into the (query, leaf) of the previous
// level that raised this query.
__global__ void fold_kernel(const uint8_t* prog, const zg_check* queries, unsigned long long n, uint32_t L,
                            const uint8_t* val, uint8_t* out, const uint32_t* parent, uint8_t* parent_val, int raw_out) {
  const Prog pr = make_prog(prog);
  unsigned long long q = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  const zg_check it = queries[q];
  uint32_t r = 2;  // invalid query -> error
  const bool ok = it.perm < pr.hdr()->n_slots && it.stype < pr.hdr()->n_types &&
                  (it.srel == kNone || (it.srel < pr.hdr()->n_slots && pr.slots()[it.srel].type == it.stype));
  if (ok) {
    const DSlot s = pr.slots()[it.perm];
    auto leafval = [&](uint32_t l) -> unsigned long long {
      const uint8_t v = val[q * L + l];
      return (v & kValT) ? 1ull : ((v & kValE) ? 2ull : 0ull);
    };
    if (s.kind != SK_NONPURE) {
      r = static_cast<uint32_t>(leafval(0));
    } else {
      const DTree t = pr.trees()[s.unit];
      unsigned long long vals = 0;
      for (uint32_t l = 0; l < t.n_leaves; ++l) vals |= leafval(l) << (2 * l);
      r = eval_tree(pr, t, vals, it.subj == it.res, it.srel);
    }
  }
  if (out) {
    if (raw_out) out[q] = r == 1 ? kValT : (r == 2 ? kValE : 0);  // value of a routed sub-query
    else out[q] = r == 1 ? ZG_HAS_PERMISSION : (r == 2 ? ZG_ITEM_ERROR : ZG_NO_PERMISSION);
  } else if (r) {
    const uint32_t pj = parent[q];
    const uint32_t bits = r == 1 ? kValT : kValE;
    atomicOr(reinterpret_cast<unsigned int*>(parent_val + (pj & ~3u)), bits << (8 * (pj & 3u)));
  }
}

// Sharded store: values of routed sub-queries come back from their owners and are OR-ed into
// the jobs that raised them.
__global__ void or_children_kernel(const uint32_t* parent, const uint8_t* child_val, unsigned long long n,
                                   uint8_t* parent_val) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t bits = child_val[i] & (kValT | kValE);
  if (!bits) return;
  const uint32_t pj = parent[i];
  atomicOr(reinterpret_cast<unsigned int*>(parent_val + (pj & ~3u)), bits << (8 * (pj & 3u)));
}

// ---- sharded store: device-side routing of checks / raised sub-queries to their owners ---------------
//
// owner(item) = item.res % n_dest. Two passes of a counting sort: per-destination counts (shared-memory
// histogram, one atomic per block and destination), then a scatter that writes every item behind its
// destination's cursor together with its source index, so that the values coming back in ROUTED order can be
// folded into the (query, leaf) that raised them without un-permuting anything.
constexpr int kMaxRouteDest = 64;
__global__ void __launch_bounds__(256) route_count_kernel(const zg_check* items, unsigned long long n, uint32_t n_dest,
                                                          unsigned long long* counts) {
  ZG_BLOCK_SHARED_U32(hist, kMaxRouteDest);
  for (uint32_t i = threadIdx.x; i < n_dest; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    atomicAdd(&hist[items[i].res % n_dest], 1u);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_dest; i += blockDim.x)
    if (hist[i]) atomicAdd(counts + i, static_cast<unsigned long long>(hist[i]));
}
__global__ void __launch_bounds__(256) route_scatter_kernel(const zg_check* items, unsigned long long n, uint32_t n_dest,
                                                            unsigned long long* cursor, zg_check* routed, uint32_t* src) {
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const zg_check it = items[i];
    const unsigned long long at = atomicAdd(cursor + (it.res % n_dest), 1ull);
    routed[at] = it;
    src[at] = static_cast<uint32_t>(i);
  }
}
// values of routed sub-queries (routed order) OR-ed into the (query, leaf) that raised sub-query src[i]
__global__ void or_children_src_kernel(const uint32_t* parent, const uint32_t* src, const uint8_t* child_val, unsigned long long n,
                                       uint8_t* parent_val) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t bits = child_val[i] & (kValT | kValE);
  if (!bits) return;
  const uint32_t pj = parent[src[i]];
  atomicOr(reinterpret_cast<unsigned int*>(parent_val + (pj & ~3u)), bits << (8 * (pj & 3u)));
}
// out[src[i]] = val[i]: answers that came back in routed order, restored to the caller's order
__global__ void unroute_kernel(const uint32_t* src, const uint8_t* val, unsigned long long n, uint8_t* out) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[src[i]] = val[i];
}

// ---- LookupResources: candidate generation by reverse BFS ----------------------------
//
// Every true result r of LookupResources(T, P, S) has a forward path r -> o1 -> ... -> ok
// where each o(i+1) is the subject object of a relationship on o(i) and S is a direct (or
// wildcard) subject on ok. Walking the reverse CSR from S therefore reaches a SUPERSET of
// the results (relation names, expiry and & / - are ignored here); the check kernel then
// verifies every candidate of type T, so the answer is exact for any schema.

struct RbfsParams {
  const uint32_t* rrow_ptr;
  const uint32_t* rcol;
  const uint8_t* prog;
  const unsigned long long* frontier;  // (type << 32 | object)
  unsigned long long n_in;
  unsigned long long* next;            // output frontier
  unsigned long long* next_count;
  unsigned long long next_cap;
  uint32_t* visited;                   // bitmap over all objects, bit = type_bit_base[type] + object
  const unsigned long long* type_bit_base;
  uint32_t want_type;                  // candidates of this type go to cand[]
  uint32_t* cand;
  unsigned long long* cand_count;
  unsigned long long cand_cap;
  uint32_t* flags;                     // bit 4: frontier / candidate overflow
  int wildcard_level;                  // level 0: also take the wildcard classes of the subject's type
};

// One warp per frontier object: for every class whose subjects have the object's type,
// read its reverse row (coalesced) and mark the resources.
__global__ void __launch_bounds__(256) rbfs_expand_kernel(const RbfsParams p) {
  const Prog pr = make_prog(p.prog);
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long w = (blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x) >> 5;
  if (w >= p.n_in) return;
  const unsigned long long item = p.frontier[w];
  const uint32_t type = static_cast<uint32_t>(item >> 32), obj = static_cast<uint32_t>(item);
  const DTypeInv tc = pr.type_rcls()[type];
  for (int ci = tc.begin; ci < tc.end; ++ci) {
    const DCls cl = pr.cls()[pr.rcls()[ci]];
    if (cl.flags & CF_EMPTY) continue;
    uint32_t row = obj;
    if (cl.sslot == kWildcard) {
      if (!p.wildcard_level) continue;  // type:* only matches the original, relation-less subject
      row = 0;
    } else if (obj >= cl.nsubj) {
      continue;
    }
    const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(row) * cl.rstride;
    const uint32_t b = __ldg(p.rrow_ptr + ri), e = __ldg(p.rrow_ptr + ri + 1);
    const unsigned long long bit_base = p.type_bit_base[cl.rtype];
    for (uint32_t i0 = b; i0 < e; i0 += 32) {
      const uint32_t i = i0 + lane;
      bool fresh = false;
      uint32_t r = 0;
      if (i < e) {
        r = __ldg(p.rcol + i);
        const unsigned long long bit = bit_base + r;
        const uint32_t mask = 1u << (bit & 31);
        fresh = !(atomicOr(p.visited + (bit >> 5), mask) & mask);
      }
      const unsigned fm = __ballot_sync(kFull, fresh);
      if (fm) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(p.next_count, static_cast<unsigned long long>(__popc(fm)));
        at = __shfl_sync(kFull, at, 0) + __popc(fm & ((1u << lane) - 1u));
        if (fresh) {
          if (at < p.next_cap) p.next[at] = (static_cast<unsigned long long>(cl.rtype) << 32) | r;
          else atomicOr(p.flags, 16u);
        }
      }
      const bool is_cand = fresh && cl.rtype == p.want_type;
      const unsigned cm = __ballot_sync(kFull, is_cand);
      if (cm) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(p.cand_count, static_cast<unsigned long long>(__popc(cm)));
        at = __shfl_sync(kFull, at, 0) + __popc(cm & ((1u << lane) - 1u));
        if (is_cand) {
          if (at < p.cand_cap) p.cand[at] = r;
          else atomicOr(p.flags, 16u);
        }
      }
    }
  }
}

// ---- batched LookupResources: K reverse walks in one frontier ---------------------------------
//
// The proxy runs one LookupResources per list request, concurrently (pkg/authz/responsefilterer.go:165,
// lookups.go:49-65). K of them share every launch: frontier entries carry the lookup they belong to,
// each lookup has its own visited bitmap slice, candidates are emitted directly as check items (the
// lookup's subject and permission filled in) with their owner beside them, and the level loop needs no
// host round trip: every level kernel reads its input count from device memory and an empty level is a
// no-op launch.
struct LookupParam {   // one per lookup of the batch
  uint32_t subj;
  uint16_t want_type, perm, stype, srel;
};
constexpr int kMaxLookupBatch = 64;
struct MrbfsParams {
  const uint32_t* rrow_ptr;
  const uint32_t* rcol;
  const uint8_t* prog;
  const LookupParam* lk;
  const unsigned long long* in;       // lookup << 48 | type << 32 | object
  unsigned long long* out;
  unsigned long long* counts;         // counts[level] = entries of `in`, counts[level + 1] = entries of `out`
  int level;
  unsigned long long cap;             // frontier capacity
  uint32_t* visited;                  // K slices of `words` u32
  unsigned long long words;
  const unsigned long long* type_bit_base;
  zg_check* cand;                     // candidate checks
  uint8_t* cand_owner;
  unsigned long long* cand_count;
  unsigned long long cand_cap;
  uint32_t* flags;                    // bit 4: frontier / candidate overflow
};

__global__ void __launch_bounds__(256) mrbfs_expand_kernel(const MrbfsParams p) {
  const unsigned long long n_in = p.counts[p.level];
  if (n_in == 0 || n_in > p.cap) return;
  const Prog pr = make_prog(p.prog);
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long n_warps = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 5;
  for (unsigned long long w = (blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x) >> 5; w < n_in;
       w += n_warps) {
    const unsigned long long item = p.in[w];
    const uint32_t k = static_cast<uint32_t>(item >> 48), type = static_cast<uint32_t>(item >> 32) & 0xFFFFu,
                   obj = static_cast<uint32_t>(item);
    const LookupParam lk = p.lk[k];
    uint32_t* const vis = p.visited + k * p.words;
    const DTypeInv tc = pr.type_rcls()[type];
    for (int ci = tc.begin; ci < tc.end; ++ci) {
      const DCls cl = pr.cls()[pr.rcls()[ci]];
      if (cl.flags & CF_EMPTY) continue;
      uint32_t row = obj;
      if (cl.sslot == kWildcard) {
        if (p.level != 0 || lk.srel != kNone) continue;  // type:* only matches the original, relation-less subject
        row = 0;
      } else if (obj >= cl.nsubj) {
        continue;
      }
      const unsigned long long ri = cl.rrow_base + static_cast<unsigned long long>(row) * cl.rstride;
      const uint32_t b = __ldg(p.rrow_ptr + ri), e = __ldg(p.rrow_ptr + ri + 1);
      const unsigned long long bit_base = p.type_bit_base[cl.rtype];
      for (uint32_t i0 = b; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        bool fresh = false;
        uint32_t r = 0;
        if (i < e) {
          r = __ldg(p.rcol + i);
          const unsigned long long bit = bit_base + r;
          const uint32_t mask = 1u << (bit & 31);
          fresh = !(atomicOr(vis + (bit >> 5), mask) & mask);
        }
        const unsigned fm = __ballot_sync(kFull, fresh);
        if (fm) {
          unsigned long long at = 0;
          if (lane == 0) at = atomicAdd(p.counts + p.level + 1, static_cast<unsigned long long>(__popc(fm)));
          at = __shfl_sync(kFull, at, 0) + __popc(fm & ((1u << lane) - 1u));
          if (fresh) {
            if (at < p.cap) p.out[at] = (static_cast<unsigned long long>(k) << 48) | (static_cast<unsigned long long>(cl.rtype) << 32) | r;
            else atomicOr(p.flags, 16u);
          }
        }
        const bool is_cand = fresh && cl.rtype == lk.want_type;
        const unsigned cm = __ballot_sync(kFull, is_cand);
        if (cm) {
          unsigned long long at = 0;
          if (lane == 0) at = atomicAdd(p.cand_count, static_cast<unsigned long long>(__popc(cm)));
          at = __shfl_sync(kFull, at, 0) + __popc(cm & ((1u << lane) - 1u));
          if (is_cand) {
            if (at < p.cand_cap) {
              zg_check q;
              q.res = r;
              q.subj = lk.subj;
              q.perm = lk.perm;
              q.stype = lk.stype;
              q.srel = lk.srel;
              q.flags = 0;
              p.cand[at] = q;
              p.cand_owner[at] = static_cast<uint8_t>(k);
            } else {
              atomicOr(p.flags, 16u);
            }
          }
        }
      }
    }
  }
}

// HAS candidates -> keys (owner << 32 | resource id), compacted; per-lookup result counts; lookups with an
// undecidable candidate (ITEM_ERROR) are flagged in err_mask.
__global__ void lookup_keys_kernel(const zg_check* cand, const uint8_t* owner, const uint8_t* codes, unsigned long long n,
                                   unsigned long long* keys, unsigned long long* n_keys, unsigned long long* per_lookup,
                                   unsigned long long* err_mask) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  const bool in = i < n;
  const uint8_t code = in ? codes[i] : 0;
  const bool has = code == ZG_HAS_PERMISSION;
  if (in && code == ZG_ITEM_ERROR) atomicOr(err_mask, 1ull << owner[i]);
  const unsigned m = __ballot_sync(kFull, has);
  if (!m) return;
  const unsigned lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(n_keys, static_cast<unsigned long long>(__popc(m)));
  base = __shfl_sync(kFull, base, 0);
  if (has) {
    keys[base + __popc(m & ((1u << lane) - 1u))] = (static_cast<unsigned long long>(owner[i]) << 32) | cand[i].res;
    atomicAdd(per_lookup + owner[i], 1ull);
  }
}
__global__ void low_words_kernel(const unsigned long long* keys, unsigned long long n, uint32_t* out) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<uint32_t>(keys[i]);
}

// LookupResources when the permission is a flat union of direct relations (the reference's
// own schema: `permission view = viewer + creator`, pkg/spicedb/bootstrap.yaml:13): the answer
// is the union of the subject's reverse rows of those classes. One warp copies them out.
struct FlatLookupClass {
  unsigned long long rrow_base;
  uint32_t nsubj;
  uint16_t stype;     // subject type the class accepts
  uint16_t wildcard;  // row 0 regardless of the subject id
  uint32_t rstride;   // reverse row stride per subject
  uint32_t pad;
};
__global__ void lookup_flat_kernel(const uint32_t* rrow_ptr, const uint32_t* rcol, const FlatLookupClass* cls, int ncls,
                                   uint32_t stype, uint32_t subj, uint32_t* out, unsigned long long cap,
                                   unsigned long long* count) {
  const unsigned lane = threadIdx.x & 31;
  unsigned long long w = 0;
  for (int c = 0; c < ncls; ++c) {
    if (cls[c].stype != stype) continue;  // relationships with another subject type cannot match
    const uint32_t row = cls[c].wildcard ? 0u : subj;
    if (row >= cls[c].nsubj) continue;
    const unsigned long long ri = cls[c].rrow_base + static_cast<unsigned long long>(row) * cls[c].rstride;
    const uint32_t b = rrow_ptr[ri], e = rrow_ptr[ri + 1];
    for (uint32_t i = b + lane; i < e; i += 32)
      if (w + (i - b) < cap) out[w + (i - b)] = rcol[i];
    w += e - b;
  }
  if (lane == 0) *count = w;
}

// ---- LookupResources helpers ---------------------------------------------------------

__global__ void lookup_fill_kernel(const uint32_t* cand, unsigned long long n, zg_check proto, zg_check* jobs) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  proto.res = cand[i];
  jobs[i] = proto;
}

// ids of candidates whose code is HAS, compacted with ballot + popc; order is
// restored by the caller (results are a set: pkg/authz/lookups.go:129).
__global__ void lookup_compact_kernel(const uint32_t* cand, const uint8_t* codes, unsigned long long n, uint32_t* out,
                                      unsigned long long cap, unsigned long long* count, uint32_t* flags) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  const bool has = i < n && codes[i] == ZG_HAS_PERMISSION;
  const bool bad = i < n && codes[i] == ZG_ITEM_ERROR;
  const unsigned m = __ballot_sync(kFull, has);
  if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 8u);
  if (!m) return;
  unsigned long long base = 0;
  const unsigned lane = threadIdx.x & 31;
  if (lane == 0) base = atomicAdd(count, static_cast<unsigned long long>(__popc(m)));
  base = __shfl_sync(kFull, base, 0);
  if (has) {
    const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
    if (at < cap) out[at] = cand[i];
  }
}

}  // namespace zg
