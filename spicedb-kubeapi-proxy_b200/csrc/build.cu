// build.cu -- builds the CSR snapshot ON THE GPU from the raw relationship list.
//
// What it replaces: the reference's embedded SpiceDB keeps relationships in an
// in-memory datastore that every write updates (pkg/spicedb/spicedb.go:50); here a write
// republishes the snapshot, so build time is write-visibility latency
// (pkg/authz/distributedtx/activity.go:54-76 expects the write to be readable on return).
//
// Pipeline (all on one stream; cub radix sorts + scans, HBM-bound streaming passes):
//   1. key kernel: relationship -> row index (type_base + object * stride + class), subject
//   2. stable radix sort by subject, then stable radix sort by row index
//      => ordered by (row, subject), equal keys in load order (TOUCH: the last one wins)
//   3. flag the last entry of every (row, subject) run, exclusive scan, compact
//      => col / exp;   histogram of rows + inclusive scan => row_ptr
//   4. reverse CSR: one more stable radix sort of (class rrow_base + subject * rstride) keys; the
//      entries are already in ascending resource order within a class; histogram + scan
//   5. per type: objects that own >= 1 relationship (cub::DeviceSelect over row_ptr)
// The host builder (store.cc) produces identical arrays; ZGPU_VERIFY_BUILD=1 compares them.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "delta.cuh"
#include "device.h"

namespace zg {

namespace {

struct BSlot {  // per relation slot
  unsigned long long row_base;
  uint32_t stride;
  uint16_t ncls;
  uint16_t cls_begin;
};

#define BCUDA(expr)                                                              \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return std::string(#expr) + ": " + cudaGetErrorString(_e); \
  } while (0)


// dead_key = pool: tombstoned relationships sort after every real row
__global__ void key_kernel(const zg_tuple* t, unsigned long long n, const BSlot* slots, const DCls* cls,
                           unsigned long long dead_key, unsigned long long* rowkey, uint32_t* subj, uint32_t* idx) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const zg_tuple x = t[i];
  idx[i] = static_cast<uint32_t>(i);
  subj[i] = x.srel == kWildcard ? 0u : x.subj;
  unsigned long long key = dead_key;
  if (!(x.flags & 1)) {
    const BSlot s = slots[x.rel];
    for (uint32_t k = 0; k < s.ncls; ++k) {
      const DCls c = cls[s.cls_begin + k];
      if (c.stype == x.stype && c.sslot == x.srel) {
        key = s.row_base + static_cast<unsigned long long>(x.res) * s.stride + k;
        break;
      }
    }
  }
  rowkey[i] = key;
}

__global__ void gather_u64_kernel(const unsigned long long* src, const uint32_t* idx, unsigned long long n,
                                  unsigned long long* dst) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// last entry of each (row, subject) run among live entries
__global__ void flag_kernel(const unsigned long long* rowkey, const uint32_t* idx, const uint32_t* subj_by_tuple,
                            unsigned long long n, unsigned long long dead_key, uint32_t* flag) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = rowkey[i];
  uint32_t f = 0;
  if (k != dead_key) {
    f = 1;
    if (i + 1 < n && rowkey[i + 1] == k && subj_by_tuple[idx[i + 1]] == subj_by_tuple[idx[i]]) f = 0;
  }
  flag[i] = f;
}

// compact unique entries; count rows and classes; emit reverse keys
__global__ void emit_kernel(const unsigned long long* rowkey, const uint32_t* idx, const uint32_t* flag,
                            const uint32_t* pos, unsigned long long n, const zg_tuple* t, const uint32_t* expires,
                            const BSlot* slots, const DCls* cls, uint32_t* col, uint32_t* exp, uint32_t* row_cnt,
                            uint32_t* cls_cnt, unsigned long long* rkey, uint32_t* rres, uint32_t* rrow_cnt) {
  unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const uint32_t p = pos[i], ti = idx[i];
  const zg_tuple x = t[ti];
  const uint32_t s = x.srel == kWildcard ? 0u : x.subj;
  col[p] = s;
  if (exp) exp[p] = expires[ti];
  const unsigned long long row = rowkey[i];
  atomicAdd(row_cnt + row + 1, 1u);
  const BSlot bs = slots[x.rel];
  const uint32_t gc = bs.cls_begin + static_cast<uint32_t>((row - bs.row_base) % bs.stride);
  atomicAdd(cls_cnt + gc, 1u);
  const unsigned long long rk = cls[gc].rrow_base + static_cast<unsigned long long>(s) * cls[gc].rstride;
  rkey[p] = rk;
  rres[p] = x.res;
  atomicAdd(rrow_cnt + rk + 1, 1u);
}

__global__ void owner_flag_kernel(const uint32_t* row_ptr, unsigned long long type_base, uint32_t stride, uint32_t n_obj,
                                  uint8_t* flag) {
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_obj) return;
  flag[o] = stride && row_ptr[type_base + static_cast<unsigned long long>(o + 1) * stride] >
                          row_ptr[type_base + static_cast<unsigned long long>(o) * stride];
}

int bits_for(unsigned long long max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) ++b;
  return b;
}

}  // namespace

std::string sort_u32(const uint32_t* d_in, uint32_t* d_out, uint64_t n, cudaStream_t st) {
  static thread_local DevBuf tmp;  // grow-only scratch per calling thread
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_in, d_out, n, 0, 32, st);
  if (!tmp.ensure(bytes + 256)) return "out of device memory (sort scratch)";
  bytes = tmp.cap;
  BCUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_in, d_out, n, 0, 32, st));
  return "";
}

std::string sort_u64(const unsigned long long* d_in, unsigned long long* d_out, uint64_t n, int end_bit, cudaStream_t st) {
  static thread_local DevBuf tmp;
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_in, d_out, n, 0, end_bit, st);
  if (!tmp.ensure(bytes + 256)) return "out of device memory (sort scratch)";
  bytes = tmp.cap;
  BCUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_in, d_out, n, 0, end_bit, st));
  return "";
}

// per type: objects that are the resource of >= 1 relationship (ascending ids), from the forward row table
std::string gpu_resource_lists(Snapshot* s, cudaStream_t st) {
  const size_t nt = s->n_objects.size();
  const unsigned blk = 256;
  s->resources.resize(nt);
  s->n_resources.assign(nt, 0);
  DevBuf d_of, d_num, d_tmp;
  struct Release {
    std::vector<DevBuf*> v;
    ~Release() {
      for (DevBuf* b : v) b->release();
    }
  } rel2{{&d_of, &d_num, &d_tmp}};
  if (!d_num.ensure(16)) return "out of device memory";
  for (size_t t = 0; t < nt; ++t) {
    const uint32_t no = s->n_objects[t];
    if (!s->resources[t].ensure(std::max<size_t>(size_t(no) * 4, 16))) return "out of device memory (resource lists)";
    if (!no || !s->type_ncls[t]) continue;
    if (!d_of.ensure(no)) return "out of device memory";
    owner_flag_kernel<<<(no + blk - 1) / blk, blk, 0, st>>>(s->row_ptr.as<uint32_t>(), s->type_base[t], s->type_ncls[t], no,
                                                            d_of.as<uint8_t>());
    size_t tb = 0;
    thrust::counting_iterator<uint32_t> ids(0);
    cub::DeviceSelect::Flagged(nullptr, tb, ids, d_of.as<uint8_t>(), s->resources[t].as<uint32_t>(), d_num.as<uint32_t>(), no, st);
    if (!d_tmp.ensure(tb + 256)) return "out of device memory (select scratch)";
    tb = d_tmp.cap;
    BCUDA(cub::DeviceSelect::Flagged(d_tmp.p, tb, ids, d_of.as<uint8_t>(), s->resources[t].as<uint32_t>(),
                                     d_num.as<uint32_t>(), no, st));
    uint32_t cnt = 0;
    BCUDA(cudaMemcpyAsync(&cnt, d_num.p, 4, cudaMemcpyDeviceToHost, st));
    BCUDA(cudaStreamSynchronize(st));
    s->n_resources[t] = cnt;
  }
  s->resources_stale = false;
  return "";
}

// Fills `s` (row_ptr, col, exp, rrow_ptr, rcol, resources) from the store's relationship list.
// `lay` comes from Store::layout(); on return lay->cls has CF_EMPTY cleared for non-empty classes.
std::string gpu_build_snapshot(const Store& store, const Schema& sc, HostSnapshot* lay, cudaStream_t st, Snapshot* s) {
  const uint64_t n = store.tuples.size();
  if (n >= 0xFFFFFFF0ull) return "more than 2^32 relationships in one snapshot";
  const uint64_t pool = lay->pool, rpool = lay->rpool;
  const bool with_exp = sc.has_expiry;

  std::vector<BSlot> bslots(sc.slots.size(), BSlot{0, 0, 0, 0});
  for (size_t i = 0; i < sc.rel_slots.size(); ++i) {
    const DRel& r = lay->rels[i];
    bslots[sc.rel_slots[i]] = BSlot{r.row_base, r.stride, r.ncls, r.cls_begin};
  }
  DevBuf d_t, d_e, d_slots, d_cls, d_key[2], d_idx[2], d_subj[2], d_flag, d_pos, d_tmp, d_rkey[2], d_rres[2], d_clscnt;
  struct Release {
    std::vector<DevBuf*> v;
    ~Release() {
      for (DevBuf* b : v) b->release();
    }
  } rel{{&d_t, &d_e, &d_slots, &d_cls, &d_key[0], &d_key[1], &d_idx[0], &d_idx[1], &d_subj[0], &d_subj[1], &d_flag,
         &d_pos, &d_tmp, &d_rkey[0], &d_rkey[1], &d_rres[0], &d_rres[1], &d_clscnt}};
  auto need = [&](DevBuf& b, size_t bytes) { return b.ensure(bytes ? bytes : 16); };
  const size_t nn = n ? n : 1;
  if (!need(d_t, nn * sizeof(zg_tuple)) || !need(d_slots, bslots.size() * sizeof(BSlot)) ||
      !need(d_cls, lay->cls.size() * sizeof(DCls)) || !need(d_key[0], nn * 8) || !need(d_key[1], nn * 8) ||
      !need(d_idx[0], nn * 4) || !need(d_idx[1], nn * 4) || !need(d_subj[0], nn * 4) || !need(d_subj[1], nn * 4) ||
      !need(d_flag, nn * 4) || !need(d_pos, nn * 4) || !need(d_clscnt, (lay->cls.size() + 1) * 4) ||
      (with_exp && !need(d_e, nn * 4)))
    return "out of device memory (snapshot build)";
  BCUDA(cudaMemcpyAsync(d_t.p, store.tuples.data(), n * sizeof(zg_tuple), cudaMemcpyHostToDevice, st));
  if (with_exp) BCUDA(cudaMemcpyAsync(d_e.p, store.expires.data(), n * 4, cudaMemcpyHostToDevice, st));
  BCUDA(cudaMemcpyAsync(d_slots.p, bslots.data(), bslots.size() * sizeof(BSlot), cudaMemcpyHostToDevice, st));
  BCUDA(cudaMemcpyAsync(d_cls.p, lay->cls.data(), lay->cls.size() * sizeof(DCls), cudaMemcpyHostToDevice, st));
  BCUDA(cudaMemsetAsync(d_clscnt.p, 0, (lay->cls.size() + 1) * 4, st));

  if (!s->row_ptr.ensure((pool + 1) * 4) || !s->rrow_ptr.ensure((rpool + 1) * 4))
    return "out of device memory (row tables)";
  BCUDA(cudaMemsetAsync(s->row_ptr.p, 0, (pool + 1) * 4, st));
  BCUDA(cudaMemsetAsync(s->rrow_ptr.p, 0, (rpool + 1) * 4, st));

  const unsigned blk = 256;
  const unsigned grid = static_cast<unsigned>((nn + blk - 1) / blk);
  uint64_t n_unique = 0;
  if (n) {
    const int kbits = bits_for(pool);  // keys are < pool, tombstones == pool
    key_kernel<<<grid, blk, 0, st>>>(d_t.as<zg_tuple>(), n, d_slots.as<BSlot>(), d_cls.as<DCls>(), pool,
                                     d_key[0].as<unsigned long long>(), d_subj[0].as<uint32_t>(), d_idx[0].as<uint32_t>());
    // pass A: stable sort of tuple indices by subject
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_subj[0].as<uint32_t>(), d_subj[1].as<uint32_t>(),
                                    d_idx[0].as<uint32_t>(), d_idx[1].as<uint32_t>(), n, 0, 32, st);
    size_t tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp2, d_key[0].as<unsigned long long>(), d_key[1].as<unsigned long long>(),
                                    d_idx[0].as<uint32_t>(), d_idx[1].as<uint32_t>(), n, 0, kbits, st);
    size_t tmp3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp3, d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), n, st);
    size_t tmp4 = 0, tmp5 = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp4, s->row_ptr.as<uint32_t>(), s->row_ptr.as<uint32_t>(), pool + 1, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmp5, s->rrow_ptr.as<uint32_t>(), s->rrow_ptr.as<uint32_t>(), rpool + 1, st);
    tmp_bytes = std::max({tmp_bytes, tmp2, tmp3, tmp4, tmp5}) + 256;
    if (!d_tmp.ensure(tmp_bytes)) return "out of device memory (sort scratch)";
    BCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_subj[0].as<uint32_t>(), d_subj[1].as<uint32_t>(),
                                          d_idx[0].as<uint32_t>(), d_idx[1].as<uint32_t>(), n, 0, 32, st));
    // pass B: stable sort by row index (dead relationships sort last)
    gather_u64_kernel<<<grid, blk, 0, st>>>(d_key[0].as<unsigned long long>(), d_idx[1].as<uint32_t>(), n,
                                            d_key[1].as<unsigned long long>());
    BCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_key[1].as<unsigned long long>(),
                                          d_key[0].as<unsigned long long>(), d_idx[1].as<uint32_t>(),
                                          d_idx[0].as<uint32_t>(), n, 0, kbits, st));
    // now d_key[0] = sorted row keys, d_idx[0] = tuple indices in (row, subject, load order);
    // d_subj[0] still holds the subject per tuple index (sort inputs are left intact)
    flag_kernel<<<grid, blk, 0, st>>>(d_key[0].as<unsigned long long>(), d_idx[0].as<uint32_t>(), d_subj[0].as<uint32_t>(), n,
                                      pool, d_flag.as<uint32_t>());
    BCUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), n, st));
    uint32_t last_pos = 0, last_flag = 0;
    BCUDA(cudaMemcpyAsync(&last_pos, d_pos.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    BCUDA(cudaMemcpyAsync(&last_flag, d_flag.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    BCUDA(cudaStreamSynchronize(st));
    n_unique = static_cast<uint64_t>(last_pos) + last_flag;
  }
  const size_t nu = n_unique ? n_unique : 1;
  // + 64: the check kernel streams rows with aligned 128-bit loads, which may read (and ignore) up to 12 bytes past a row
  // edge arrays of a large store get the growth headroom of the incremental publish from the start (see below)
  const size_t edge_bytes = static_cast<size_t>(nu) * 4 + 64,
               edge_cap = nu >= (1u << 20) ? edge_bytes + edge_bytes / 32 + (4u << 20) : edge_bytes;
  if (!s->col.ensure(edge_cap) || !s->rcol.ensure(edge_cap) || (with_exp && !s->exp.ensure(edge_cap)) ||
      !d_rkey[0].ensure(nu * 8) || !d_rkey[1].ensure(nu * 8) || !d_rres[0].ensure(nu * 4) || !d_rres[1].ensure(nu * 4))
    return "out of device memory (edge arrays)";
  if (n) {
    emit_kernel<<<grid, blk, 0, st>>>(d_key[0].as<unsigned long long>(), d_idx[0].as<uint32_t>(), d_flag.as<uint32_t>(),
                                      d_pos.as<uint32_t>(), n, d_t.as<zg_tuple>(), with_exp ? d_e.as<uint32_t>() : nullptr,
                                      d_slots.as<BSlot>(), d_cls.as<DCls>(), s->col.as<uint32_t>(),
                                      with_exp ? s->exp.as<uint32_t>() : nullptr, s->row_ptr.as<uint32_t>(),
                                      d_clscnt.as<uint32_t>(), d_rkey[0].as<unsigned long long>(), d_rres[0].as<uint32_t>(),
                                      s->rrow_ptr.as<uint32_t>());
    size_t tb = d_tmp.cap;
    BCUDA(cub::DeviceScan::InclusiveSum(d_tmp.p, tb, s->row_ptr.as<uint32_t>(), s->row_ptr.as<uint32_t>(), pool + 1, st));
    tb = d_tmp.cap;
    BCUDA(cub::DeviceScan::InclusiveSum(d_tmp.p, tb, s->rrow_ptr.as<uint32_t>(), s->rrow_ptr.as<uint32_t>(), rpool + 1, st));
  }
  if (n_unique) {
    // reverse CSR: entries are in (type, object, class) order, i.e. ascending resource within a
    // class; a stable sort by (class reverse base + subject) finishes the job
    size_t tb = 0;
    const int rbits = bits_for(rpool);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, d_rkey[0].as<unsigned long long>(), d_rkey[1].as<unsigned long long>(),
                                    d_rres[0].as<uint32_t>(), s->rcol.as<uint32_t>(), n_unique, 0, rbits, st);
    if (!d_tmp.ensure(tb + 256)) return "out of device memory (sort scratch)";
    tb = d_tmp.cap;
    BCUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, d_rkey[0].as<unsigned long long>(),
                                          d_rkey[1].as<unsigned long long>(), d_rres[0].as<uint32_t>(),
                                          s->rcol.as<uint32_t>(), n_unique, 0, rbits, st));
  }
  // classes that hold at least one relationship
  std::vector<uint32_t> cls_cnt(lay->cls.size() + 1, 0);
  BCUDA(cudaMemcpyAsync(cls_cnt.data(), d_clscnt.p, cls_cnt.size() * 4, cudaMemcpyDeviceToHost, st));
  BCUDA(cudaStreamSynchronize(st));
  for (size_t c = 0; c < lay->cls.size(); ++c)
    if (cls_cnt[c]) lay->cls[c].flags &= static_cast<uint16_t>(~CF_EMPTY);

  s->cls_count.assign(cls_cnt.begin(), cls_cnt.end() - 1);
  s->type_base = lay->type_base;
  s->type_ncls = lay->type_ncls;
  s->n_objects = lay->n_objects;
  s->delta_ok = true;
  if (nu >= (1u << 20)) {
    // The alternates of the incremental publish (merge_direction), with its headroom, come with the snapshot: the first
    // write after a bulk load must not pay for three store-sized allocations (8 .. 560 ms, profiles/r2l_write_bench.json).
    // Best effort: a failure here is met again, and reported, by the write that needs them.
    const size_t need = static_cast<size_t>(nu) * 4 + 64, roomy = need + need / 32 + (4u << 20);
    s->col_alt.ensure(roomy);
    s->rcol_alt.ensure(roomy);
    if (with_exp) s->exp_alt.ensure(roomy);
    cudaGetLastError();
  }
  {
    std::string rerr = gpu_resource_lists(s, st);
    if (!rerr.empty()) return rerr;
  }
  BCUDA(cudaGetLastError());
  lay->n_tuples = n_unique;
  s->n_tuples = n_unique;
  s->bytes += (pool + 1) * 4 + (rpool + 1) * 4 + n_unique * (with_exp ? 12 : 8);
  return "";
}


namespace {

// one direction (forward or reverse) of the merge
std::string merge_direction(const HostDelta& h, bool with_exp, DevBuf& row_ptr, uint64_t pool, DevBuf& edges, DevBuf& exp,
                            DevBuf& edges_alt, DevBuf& exp_alt, uint64_t n_old, cudaStream_t st, DevBuf& scratch,
                            uint32_t* d_bad) {
  const uint32_t ni = static_cast<uint32_t>(h.ins_key.size()), nd = static_cast<uint32_t>(h.del_key.size()),
                 nt = with_exp ? static_cast<uint32_t>(h.tch_key.size()) : 0u;
  const uint64_t n_new = n_old + ni - nd;
  if (n_new >= 0xFFFFFFF0ull) return "more than 2^32 relationships in one snapshot";
  // scratch layout (u64 keys first, then u32 arrays)
  const size_t keys = size_t(ni) + nd + nt;
  const size_t bytes = keys * 8 + (size_t(ni) * 3 + size_t(nd) * 2 + size_t(nt) * 3) * 4 + 64;
  if (!scratch.ensure(bytes)) return "out of device memory (delta scratch)";
  uint8_t* base = scratch.as<uint8_t>();
  auto* d_ins_key = reinterpret_cast<unsigned long long*>(base);
  auto* d_del_key = d_ins_key + ni;
  auto* d_tch_key = d_del_key + nd;
  auto* u32p = reinterpret_cast<uint32_t*>(d_tch_key + nt);
  uint32_t* d_ins_val = u32p; u32p += ni;
  uint32_t* d_ins_exp = u32p; u32p += ni;
  uint32_t* d_ins_pos = u32p; u32p += ni;
  uint32_t* d_del_val = u32p; u32p += nd;
  uint32_t* d_del_pos = u32p; u32p += nd;
  uint32_t* d_tch_val = u32p; u32p += nt;
  uint32_t* d_tch_exp = u32p; u32p += nt;
  uint32_t* d_tch_pos = u32p; u32p += nt;
  auto up = [&](void* dst, const void* src, size_t b) -> cudaError_t {
    return b ? cudaMemcpyAsync(dst, src, b, cudaMemcpyHostToDevice, st) : cudaSuccess;
  };
  BCUDA(up(d_ins_key, h.ins_key.data(), size_t(ni) * 8));
  BCUDA(up(d_del_key, h.del_key.data(), size_t(nd) * 8));
  BCUDA(up(d_ins_val, h.ins_val.data(), size_t(ni) * 4));
  BCUDA(up(d_del_val, h.del_val.data(), size_t(nd) * 4));
  if (with_exp) {
    BCUDA(up(d_ins_exp, h.ins_exp.data(), size_t(ni) * 4));
    BCUDA(up(d_tch_key, h.tch_key.data(), size_t(nt) * 8));
    BCUDA(up(d_tch_val, h.tch_val.data(), size_t(nt) * 4));
    BCUDA(up(d_tch_exp, h.tch_exp.data(), size_t(nt) * 4));
  }
  DeltaDev d{d_ins_key, d_ins_val, with_exp ? d_ins_exp : nullptr, d_ins_pos, ni, d_del_key, d_del_val, d_del_pos, nd,
             d_tch_key, d_tch_val, d_tch_exp, d_tch_pos, nt};
  const unsigned blk = 256;
  if (keys) delta_locate_kernel<<<static_cast<unsigned>((keys + blk - 1) / blk), blk, 0, st>>>(row_ptr.as<uint32_t>(),
                                                                                             edges.as<uint32_t>(), d, d_bad);
  if (ni || nd) {
    // headroom when the alternate has to grow (3 % + 4 MB): without it every write that adds edges re-allocates a
    // store-sized buffer, and a cudaMalloc / cudaFree of 400 MB costs 100 ms on some boxes (profiles/r2k_write_bench.json)
    const size_t need = std::max<uint64_t>(n_new, 1) * 4 + 64, roomy = need + need / 32 + (4u << 20);
    if (!edges_alt.ensure(edges_alt.cap < need ? roomy : need) || (with_exp && !exp_alt.ensure(exp_alt.cap < need ? roomy : need)))
      return "out of device memory (edge arrays)";
    if (n_old)
      delta_merge_kernel<<<static_cast<unsigned>((n_old + kMergeTile - 1) / kMergeTile), 256, 0, st>>>(
          edges.as<uint32_t>(), edges_alt.as<uint32_t>(), with_exp ? exp.as<uint32_t>() : nullptr,
          with_exp ? exp_alt.as<uint32_t>() : nullptr, static_cast<uint32_t>(n_old), d_ins_pos, ni, d_del_pos, nd);
    if (ni)
      delta_insert_kernel<<<(ni + blk - 1) / blk, blk, 0, st>>>(edges_alt.as<uint32_t>(), with_exp ? exp_alt.as<uint32_t>() : nullptr,
                                                              d_ins_pos, d_ins_val, with_exp ? d_ins_exp : nullptr, ni, d_del_pos, nd);
    if (nt)
      delta_touch_kernel<<<(nt + blk - 1) / blk, blk, 0, st>>>(exp_alt.as<uint32_t>(), d_tch_pos, d_tch_exp, nt, d_ins_pos, ni,
                                                             d_del_pos, nd);
    unsigned long long first = pool;
    if (ni) first = std::min<unsigned long long>(first, h.ins_key.front());
    if (nd) first = std::min<unsigned long long>(first, h.del_key.front());
    ++first;  // row_ptr[key] itself (the start of the first touched row) does not move
    if (first <= pool)
      delta_rows_kernel<<<static_cast<unsigned>((pool + 1 - first + kMergeTile - 1) / kMergeTile), 256, 0, st>>>(
          row_ptr.as<uint32_t>(), first, pool, d_ins_key, ni, d_del_key, nd);
    std::swap(edges, edges_alt);
    if (with_exp) std::swap(exp, exp_alt);
  } else if (nt) {
    // expirations only: patch in place (positions do not move)
    delta_touch_kernel<<<(nt + blk - 1) / blk, blk, 0, st>>>(exp.as<uint32_t>(), d_tch_pos, d_tch_exp, nt, d_ins_pos, 0, d_del_pos, 0);
  }
  BCUDA(cudaGetLastError());
  return "";
}

}  // namespace

std::string gpu_apply_delta(const Store& store, const Schema& sc, const HostSnapshot& lay, cudaStream_t st, Snapshot* s,
                            std::vector<uint32_t>* cls_delta) {
  const bool with_exp = sc.has_expiry;
  HostDelta f, r;
  {
    std::string perr = prepare_delta(store, sc, lay, &f, &r, cls_delta);
    if (!perr.empty()) return perr;
  }
  if (f.ins_key.empty() && f.del_key.empty() && f.tch_key.empty()) return "";
  static thread_local DevBuf scratch_f, scratch_r, bad;
  if (!bad.ensure(16)) return "out of device memory";
  BCUDA(cudaMemsetAsync(bad.p, 0, 16, st));
  const uint64_t n_old = s->n_tuples;
  DevBuf none1, none2;
  std::string err = merge_direction(f, with_exp, s->row_ptr, lay.pool, s->col, s->exp, s->col_alt, s->exp_alt, n_old, st,
                                    scratch_f, bad.as<uint32_t>());
  if (!err.empty()) return err;
  err = merge_direction(r, false, s->rrow_ptr, lay.rpool, s->rcol, none1, s->rcol_alt, none2, n_old, st, scratch_r,
                        bad.as<uint32_t>() + 1);
  if (!err.empty()) return err;
  uint32_t hb[2] = {0, 0};
  BCUDA(cudaMemcpyAsync(hb, bad.p, 8, cudaMemcpyDeviceToHost, st));
  BCUDA(cudaStreamSynchronize(st));
  if (hb[0] || hb[1]) return "corrupt";  // the resident snapshot disagrees with the store: rebuild it
  s->n_tuples = n_old + f.ins_key.size() - f.del_key.size();
  return "";
}

}  // namespace zg
