// Mutation fuzzer for zg_list_scan / zg_list_filter (csrc/listfilter.cc), built with
// -fsanitize=address,undefined by tests/test_listfilter.py::test_scanner_fuzz_under_sanitizers.
// The scanner reads bytes that come from the network; every input, however broken, must end in a
// return code, with every reported range inside the body and the splice exactly as long as announced.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zgpu.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}

static const char* kSeeds[] = {
    R"({"kind":"PodList","apiVersion":"v1","metadata":{"resourceVersion":"1"},"items":[{"metadata":{"name":"a","namespace":"n"},"spec":{"x":[1,2.5e-3,true,null,{"y":"]}"}]}},{"metadata":{"name":"b\u00e9\"","labels":{"k":"v"}}},3,"s",null,[{"metadata":{"name":"z"}}]]})",
    R"({"kind":"Table","rows":[{"cells":["a",1],"object":{"kind":"PartialObjectMetadata","metadata":{"name":"a","namespace":"n"}}},{"cells":[],"object":null},{"object":{"metadata":7}}],"columnDefinitions":[{"name":"Name"}]})",
    R"(  { "items" : [ { "metadata" : { "name" : "sp aced" , "namespace" : "\\n\t" } , "metadata" : { "name" : "dup" } } ] , "items" : [ ] } )",
    R"({"\u0069tems":[{"\u006detadata":{"n\u0061me":"esc"}}],"x":-0.0e+10,"y":[[[[[[]]]]]],"z":{"a":{"b":{"c":{}}}}})",
    R"({"items":[]})",
    R"({})",
};

int main(int argc, char** argv) {
  const long iters = argc > 1 ? std::atol(argv[1]) : 200000;
  const char structural[] = "{}[]\",:\\ \n0-e.tfnu";
  long accepted = 0, filtered = 0;
  std::vector<zg_list_item> items(64);
  std::vector<uint8_t> keep(64);
  std::string out;
  for (long it = 0; it < iters; ++it) {
    std::string s = kSeeds[rnd() % (sizeof(kSeeds) / sizeof(kSeeds[0]))];
    const int muts = static_cast<int>(rnd() % 4);  // 0 = the seed itself
    for (int m = 0; m < muts && !s.empty(); ++m) {
      const size_t p = rnd() % s.size();
      switch (rnd() % 6) {
        case 0: s[p] = structural[rnd() % (sizeof(structural) - 1)]; break;
        case 1: s.erase(p, 1 + rnd() % 3); break;
        case 2: s.insert(p, 1, structural[rnd() % (sizeof(structural) - 1)]); break;
        case 3: s.resize(p); break;
        case 4: s[p] = static_cast<char>(rnd()); break;
        default: s.insert(p, s.substr(p / 2, rnd() % 24)); break;
      }
    }
    // exact-size heap copy, NOT NUL terminated: an over-read is an ASan report
    char* body = static_cast<char*>(std::malloc(s.size() ? s.size() : 1));
    std::memcpy(body, s.data(), s.size());
    for (uint32_t mode = 0; mode < 2; ++mode) {
      const uint64_t cap = rnd() % 3 == 0 ? rnd() % 3 : items.size();
      uint64_t ib = 0, ie = 0;
      const int64_t n = zg_list_scan(body, s.size(), mode, cap ? items.data() : nullptr, cap, &ib, &ie);
      if (n < 0) {
        if (n != ZG_EINVAL && n != ZG_E2BIG) return std::printf("unexpected code %lld\n", static_cast<long long>(n)), 1;
        continue;
      }
      ++accepted;
      if (ib > ie || ie > s.size()) return std::printf("array range out of bounds\n"), 1;
      if (n == 0 || !cap) continue;
      uint64_t prev = ib;
      for (int64_t i = 0; i < n; ++i) {
        const zg_list_item& x = items[static_cast<size_t>(i)];
        if (x.begin < prev || x.end < x.begin || x.end > ie) return std::printf("item range out of order\n"), 1;
        if (x.name_len && (x.name_off < x.begin || x.name_off + x.name_len > x.end)) return std::printf("name range\n"), 1;
        if (x.ns_len && (x.ns_off < x.begin || x.ns_off + x.ns_len > x.end)) return std::printf("ns range\n"), 1;
        prev = x.end;
        keep[static_cast<size_t>(i)] = rnd() & 1;
      }
      size_t need = 0;
      const uint32_t flags = rnd() & 1;
      if (zg_list_filter(body, s.size(), items.data(), static_cast<uint64_t>(n), keep.data(), ib, ie, flags, nullptr, 0, &need) != ZG_E2BIG)
        return std::printf("size query did not return E2BIG\n"), 1;
      char* o = static_cast<char*>(std::malloc(need ? need : 1));
      size_t wrote = 0;
      if (zg_list_filter(body, s.size(), items.data(), static_cast<uint64_t>(n), keep.data(), ib, ie, flags, o, need, &wrote) != ZG_OK || wrote != need)
        return std::printf("filter wrote %zu, announced %zu\n", wrote, need), 1;
      // the filtered body must itself scan, and to the number of kept items
      uint64_t ib2, ie2;
      int64_t kept = 0;
      for (int64_t i = 0; i < n; ++i) kept += keep[static_cast<size_t>(i)];
      const int64_t n2 = zg_list_scan(o, wrote, mode, nullptr, 0, &ib2, &ie2);
      if (n2 != kept && !(kept == 0 && n2 == 0)) {
        // a later duplicate "items" key can shadow the filtered one only if the input had duplicates:
        // the scanner reports the LAST array, the splice rewrote that same one, so counts must agree
        return std::printf("rescan found %lld items, kept %lld\n", static_cast<long long>(n2), static_cast<long long>(kept)), 1;
      }
      std::free(o);
      ++filtered;
    }
    std::free(body);
  }
  // ---- protobuf-encoded lists (mode ZG_LIST_PROTOBUF): byte-level mutations of small valid envelopes
  auto put_varint = [](std::string& o, uint64_t v) {
    while (v >= 0x80) {
      o.push_back(static_cast<char>((v & 0x7F) | 0x80));
      v >>= 7;
    }
    o.push_back(static_cast<char>(v));
  };
  auto field = [&](std::string& o, uint32_t num, const std::string& payload) {
    put_varint(o, (static_cast<uint64_t>(num) << 3) | 2);
    put_varint(o, payload.size());
    o += payload;
  };
  long pb_accepted = 0, pb_filtered = 0;
  for (long it = 0; it < iters / 2; ++it) {
    std::string raw, lm;
    field(lm, 2, "12345");  // ListMeta.resourceVersion
    field(raw, 1, lm);
    const int ni = static_cast<int>(rnd() % 5);
    for (int i = 0; i < ni; ++i) {
      std::string om, item;
      if (rnd() % 8) field(om, 1, std::string(1 + rnd() % 140, static_cast<char>('a' + i)));
      if (rnd() % 3) field(om, 3, "ns-" + std::to_string(rnd() % 3));
      field(om, 5, "uid");
      if (rnd() % 8) field(item, 1, om);
      field(item, 2, std::string(rnd() % 200, 's'));
      if (rnd() % 4 == 0) {  // a varint and a fixed64 field the scanner must skip
        put_varint(item, (7u << 3) | 0);
        put_varint(item, rnd());
        put_varint(item, (9u << 3) | 1);
        item += std::string(8, '\x01');
      }
      field(raw, 2, item);
    }
    std::string s("k8s\0", 4), tm;
    field(tm, 1, "v1");
    field(tm, 2, "PodList");
    field(s, 1, tm);
    field(s, 2, raw);
    field(s, 3, "");
    field(s, 4, "");
    const int muts = static_cast<int>(rnd() % 3);
    for (int m = 0; m < muts && !s.empty(); ++m) {
      const size_t p = rnd() % s.size();
      switch (rnd() % 5) {
        case 0: s[p] = static_cast<char>(rnd()); break;
        case 1: s.erase(p, 1 + rnd() % 3); break;
        case 2: s.insert(p, 1, static_cast<char>(rnd())); break;
        case 3: s.resize(p); break;
        default: s[p] = static_cast<char>(s[p] ^ (1 << (rnd() % 8))); break;
      }
    }
    char* body = static_cast<char*>(std::malloc(s.size() ? s.size() : 1));
    std::memcpy(body, s.data(), s.size());
    uint64_t ib = 0, ie = 0;
    {  // the single-object view of the same bytes: one item, inside raw
      zg_list_item one;
      const int64_t n1 = zg_list_scan(body, s.size(), ZG_LIST_PROTOBUF_OBJECT, &one, 1, &ib, &ie);
      if (n1 > 1 || (n1 < 0 && n1 != ZG_EINVAL)) return std::printf("pb object: code %lld\n", static_cast<long long>(n1)), 1;
      if (n1 == 1 && (one.begin > one.end || one.end > ie || ie > s.size() ||
                      (one.name_len && (one.name_off < one.begin || one.name_off + one.name_len > one.end))))
        return std::printf("pb object: ranges\n"), 1;
    }
    const int64_t n = zg_list_scan(body, s.size(), ZG_LIST_PROTOBUF, items.data(), items.size(), &ib, &ie);
    if (n < 0) {
      if (n != ZG_EINVAL && n != ZG_E2BIG) return std::printf("pb: unexpected code %lld\n", static_cast<long long>(n)), 1;
      std::free(body);
      continue;
    }
    ++pb_accepted;
    if (ib > ie || ie > s.size()) return std::printf("pb: raw range out of bounds\n"), 1;
    if (n > 0) {
      uint64_t prev = ib;
      int64_t kept = 0;
      for (int64_t i = 0; i < n; ++i) {
        const zg_list_item& x = items[static_cast<size_t>(i)];
        if (x.begin < prev || x.end < x.begin || x.end > ie) return std::printf("pb: item range out of order\n"), 1;
        if (x.name_len && (x.name_off < x.begin || x.name_off + x.name_len > x.end)) return std::printf("pb: name range\n"), 1;
        if (x.ns_len && (x.ns_off < x.begin || x.ns_off + x.ns_len > x.end)) return std::printf("pb: ns range\n"), 1;
        prev = x.end;
        kept += (keep[static_cast<size_t>(i)] = rnd() & 1);
      }
      size_t need = 0;
      if (zg_list_filter(body, s.size(), items.data(), static_cast<uint64_t>(n), keep.data(), ib, ie, 0, nullptr, 0, &need) != ZG_E2BIG)
        return std::printf("pb: size query did not return E2BIG\n"), 1;
      char* o = static_cast<char*>(std::malloc(need ? need : 1));
      size_t wrote = 0;
      if (zg_list_filter(body, s.size(), items.data(), static_cast<uint64_t>(n), keep.data(), ib, ie, 0, o, need, &wrote) != ZG_OK || wrote != need)
        return std::printf("pb: filter wrote %zu, announced %zu\n", wrote, need), 1;
      uint64_t ib2, ie2;
      const int64_t n2 = zg_list_scan(o, wrote, ZG_LIST_PROTOBUF, nullptr, 0, &ib2, &ie2);
      if (n2 != kept) return std::printf("pb: rescan found %lld items, kept %lld\n", static_cast<long long>(n2), static_cast<long long>(kept)), 1;
      std::free(o);
      ++pb_filtered;
    }
    std::free(body);
  }
  std::printf("ok iterations=%ld accepted=%ld filtered=%ld pb_accepted=%ld pb_filtered=%ld\n", iters, accepted, filtered, pb_accepted,
              pb_filtered);
  return 0;
}
