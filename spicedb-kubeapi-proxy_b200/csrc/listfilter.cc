// listfilter.cc -- SURVEY.md 8(f) rank 1: the list-response filter either side of the engine.
//
// Reference behaviour (pkg/authz/postfilter.go:17-55, 67-119): json.Unmarshal the whole kube
// List body into map[string]interface{}, read metadata.name / metadata.namespace of every
// item, check, then json.Marshal the kept items back. Here: ONE pass over the bytes that only
// tokenises structure -- it records each item's byte range and the (still escaped) name /
// namespace strings -- and a splice that copies the kept ranges. Nothing is materialised, no
// floats are re-printed, unknown fields survive byte for byte.
//
// Pure host code (no CUDA): the checks in between go through zg_check_bulk_str.
#include <cstdint>
#include <cstring>
#include <initializer_list>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/zgpu.h"

namespace {

struct Cur {
  const char* s;
  size_t n, i;
  bool ok;
};

inline void ws(Cur& c) {
  while (c.i < c.n && (c.s[c.i] == ' ' || c.s[c.i] == '\n' || c.s[c.i] == '\t' || c.s[c.i] == '\r')) ++c.i;
}

inline int hexv(char h) {
  if (h >= '0' && h <= '9') return h - '0';
  if (h >= 'a' && h <= 'f') return h - 'a' + 10;
  if (h >= 'A' && h <= 'F') return h - 'A' + 10;
  return -1;
}

// Pointer-style primitives (everything stays in registers): each returns the position after what
// it consumed, or nullptr on malformed input.

// p at the opening quote; returns the position after the closing quote, *close = its position.
inline const char* str_end(const char* p, const char* end, const char** close) {
  if (p >= end || *p != '"') return nullptr;
  ++p;
  while (p < end) {
#if defined(__SSE2__)
    // 16 bytes at a time up to the next quote, backslash or control character
    while (p + 16 <= end) {
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
      const __m128i special = _mm_or_si128(
          _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('"')), _mm_cmpeq_epi8(v, _mm_set1_epi8('\\'))),
          _mm_cmpeq_epi8(_mm_max_epu8(v, _mm_set1_epi8(0x1f)), _mm_set1_epi8(0x1f)));  // byte <= 0x1f
      const int m = _mm_movemask_epi8(special);
      if (m) {
        p += __builtin_ctz(m);
        break;
      }
      p += 16;
    }
    if (p >= end) break;
#endif
    const char ch = *p;
    if (ch == '"') {
      *close = p;
      return p + 1;
    }
    if (ch == '\\') {
      if (p + 1 >= end) break;
      const char x = p[1];
      if (x == 'u') {
        if (p + 6 > end || hexv(p[2]) < 0 || hexv(p[3]) < 0 || hexv(p[4]) < 0 || hexv(p[5]) < 0) break;
        p += 6;
        continue;
      }
      if (x != '"' && x != '\\' && x != '/' && x != 'b' && x != 'f' && x != 'n' && x != 'r' && x != 't') break;
      p += 2;
      continue;
    }
    if (static_cast<unsigned char>(ch) < 0x20) break;  // raw control characters are not JSON
    ++p;
  }
  return nullptr;
}

inline const char* skip_ws(const char* p, const char* end) {
  if (p < end && static_cast<unsigned char>(*p) > ' ') return p;  // compact bodies: nothing to skip
  while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  return p;
}

// true / false / null / number: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?
inline const char* scalar_end(const char* p, const char* end) {
  const char ch = *p;
  if (ch == 't') return (end - p >= 4 && std::memcmp(p, "true", 4) == 0) ? p + 4 : nullptr;
  if (ch == 'f') return (end - p >= 5 && std::memcmp(p, "false", 5) == 0) ? p + 5 : nullptr;
  if (ch == 'n') return (end - p >= 4 && std::memcmp(p, "null", 4) == 0) ? p + 4 : nullptr;
  if (p < end && *p == '-') ++p;
  if (p < end && *p == '0') {
    ++p;
  } else {
    const char* d = p;
    while (p < end && *p >= '0' && *p <= '9') ++p;
    if (p == d) return nullptr;
  }
  if (p < end && *p == '.') {
    const char* d = ++p;
    while (p < end && *p >= '0' && *p <= '9') ++p;
    if (p == d) return nullptr;
  }
  if (p < end && (*p == 'e' || *p == 'E')) {
    ++p;
    if (p < end && (*p == '+' || *p == '-')) ++p;
    const char* d = p;
    while (p < end && *p >= '0' && *p <= '9') ++p;
    if (p == d) return nullptr;
  }
  return p;
}

constexpr int kMaxDepth = 512;  // containers nested below the value a skip starts at

// Validates ONE value of any shape starting at p and returns the position after it. Iterative (an
// explicit stack of container kinds), strict: commas, colons, literals, numbers and string escapes
// are all checked -- this is where nearly all bytes of a list body are spent.
const char* value_end(const char* p, const char* end, int depth0) {
  char stack[kMaxDepth + 2];
  int sp = 0;
  for (;;) {
    // ---- a value is expected at p
    p = skip_ws(p, end);
    if (p >= end) return nullptr;
    const char ch = *p;
    if (ch == '"') {
      const char* q;
      if (!(p = str_end(p, end, &q))) return nullptr;
    } else if (ch == '{' || ch == '[') {
      if (depth0 + sp > kMaxDepth) return nullptr;
      stack[sp++] = ch;
      p = skip_ws(p + 1, end);
      if (p >= end) return nullptr;
      if (ch == '[') {
        if (*p != ']') continue;  // first element
        ++p;
        --sp;
      } else if (*p == '}') {
        ++p;
        --sp;
      } else {
        goto key;
      }
    } else if (!(p = scalar_end(p, end))) {
      return nullptr;
    }
    // ---- a value just ended
    for (;;) {
      if (sp == 0) return p;
      p = skip_ws(p, end);
      if (p >= end) return nullptr;
      const char c2 = *p++;
      if (stack[sp - 1] == '{') {
        if (c2 == ',') goto key_ws;
        if (c2 != '}') return nullptr;
      } else {
        if (c2 == ',') break;  // next element
        if (c2 != ']') return nullptr;
      }
      --sp;
    }
    continue;
  key_ws:
    p = skip_ws(p, end);
  key : {
    const char* q;
    if (!(p = str_end(p, end, &q))) return nullptr;
    p = skip_ws(p, end);
    if (p >= end || *p != ':') return nullptr;
    ++p;
  }
  }
}

// Cursor adapters for the outer walkers below.
inline bool skip_string(Cur& c, size_t* b, size_t* e) {
  const char* q;
  const char* p = str_end(c.s + c.i, c.s + c.n, &q);
  if (!p) return c.ok = false;
  *b = c.i + 1;
  *e = static_cast<size_t>(q - c.s);
  c.i = static_cast<size_t>(p - c.s);
  return true;
}

inline bool skip_value(Cur& c, int depth) {
  const char* p = value_end(c.s + c.i, c.s + c.n, depth);
  if (!p) return c.ok = false;
  c.i = static_cast<size_t>(p - c.s);
  return true;
}

// Key comparison after JSON unescaping (encoding/json compares decoded keys): the keys looked
// for are ASCII, so only escapes that decode to ASCII can match.
inline bool key_is(const Cur& c, size_t b, size_t e, const char* k) {
  const size_t n = std::strlen(k);
  if (e - b == n) return std::memcmp(c.s + b, k, n) == 0;
  if (e - b < n || !std::memchr(c.s + b, '\\', e - b)) return false;
  size_t j = 0;
  for (size_t i = b; i < e; ++j) {
    if (j >= n) return false;
    char ch = c.s[i++];
    if (ch == '\\') {
      if (i >= e) return false;
      const char x = c.s[i++];
      if (x == 'u') {
        if (i + 4 > e) return false;
        int v = 0;
        for (int t = 0; t < 4; ++t) {
          const int h = hexv(c.s[i + t]);
          if (h < 0) return false;
          v = v * 16 + h;
        }
        i += 4;
        if (v >= 0x80) return false;
        ch = static_cast<char>(v);
      } else if (x == '/' || x == '\\' || x == '"') {
        ch = x;
      } else {
        return false;  // \b \f \n \r \t never occur in the keys looked for
      }
    }
    if (ch != k[j]) return false;
  }
  return j == n;
}

// Inside an object positioned at '{': calls visit(key_begin, key_end) before each value; visit
// returns true if it consumed the value itself.
template <class F>
bool walk_object(Cur& c, F visit) {
  if (c.i >= c.n || c.s[c.i] != '{') return c.ok = false;
  ++c.i;
  ws(c);
  if (c.i < c.n && c.s[c.i] == '}') {
    ++c.i;
    return true;
  }
  for (;;) {
    ws(c);
    size_t b, e;
    if (!skip_string(c, &b, &e)) return false;
    ws(c);
    if (c.i >= c.n || c.s[c.i] != ':') return c.ok = false;
    ++c.i;
    ws(c);
    if (!visit(b, e)) {
      if (!c.ok) return false;
      if (!skip_value(c, 1)) return false;
    }
    ws(c);
    if (c.i >= c.n) return c.ok = false;
    if (c.s[c.i] == ',') {
      ++c.i;
      continue;
    }
    if (c.s[c.i] == '}') {
      ++c.i;
      return true;
    }
    return c.ok = false;
  }
}

inline void clear_meta(zg_list_item* it) {
  it->name_off = it->ns_off = 0;
  it->name_len = it->ns_len = 0;
  it->flags &= ~ZG_ITEM_HAS_METADATA;
}

// Walks the object at the cursor (the holder of "metadata": a list item, or a table row's
// "object") and records metadata.name / metadata.namespace. encoding/json keeps the LAST of
// duplicate keys, at every level; so does this.
bool scan_holder(Cur& c, zg_list_item* it) {
  return walk_object(c, [&](size_t kb, size_t ke) -> bool {
    if (!key_is(c, kb, ke, "metadata")) return false;
    clear_meta(it);
    if (c.i >= c.n || c.s[c.i] != '{') return false;  // metadata that is not an object: no ObjectMeta
    it->flags |= ZG_ITEM_HAS_METADATA;
    return walk_object(c, [&](size_t b2, size_t e2) -> bool {
      const bool is_name = key_is(c, b2, e2, "name"), is_ns = key_is(c, b2, e2, "namespace");
      if (!is_name && !is_ns) return false;
      uint64_t& off = is_name ? it->name_off : it->ns_off;
      uint32_t& ln = is_name ? it->name_len : it->ns_len;
      off = ln = 0;
      if (c.i >= c.n || c.s[c.i] != '"') return false;  // non-string: the type assertion fails, stays ""
      size_t vb, ve;
      if (!skip_string(c, &vb, &ve)) return false;
      off = vb;
      ln = static_cast<uint32_t>(ve - vb);
      return true;
    });
  });
}

bool scan_item(Cur& c, zg_list_item* it, uint32_t mode) {
  ws(c);
  it->begin = c.i;
  it->flags = it->reserved = 0;
  clear_meta(it);
  if (c.i < c.n && c.s[c.i] == '{') {
    it->flags = ZG_ITEM_IS_OBJECT;
    bool ok;
    if (mode == ZG_LIST_ITEMS) {
      ok = scan_holder(c, it);
    } else {  // ZG_LIST_TABLE_ROWS: the object sits under the row's "object" key
      ok = walk_object(c, [&](size_t kb, size_t ke) -> bool {
        if (!key_is(c, kb, ke, "object")) return false;
        clear_meta(it);
        it->flags |= ZG_ITEM_HAS_OBJECT;
        if (c.i >= c.n || c.s[c.i] != '{') return false;  // null (includeObject=None) or a scalar
        return scan_holder(c, it);
      });
    }
    if (!ok) return false;
  } else if (!skip_value(c, 1)) {  // a non-object element: no metadata
    return false;
  }
  it->end = c.i;
  return true;
}

}  // namespace


// ---- protobuf-encoded lists (application/vnd.kubernetes.protobuf) --------------------------------
// The reference decodes a list response with the serializer the Content-Type names
// (pkg/authz/responsefilterer.go:256-266, :301-313); for built-in types kube clients ask for protobuf. Wire format
// (k8s.io/apimachinery v0.34.1, pinned in the reference's go.mod; not vendored there, restated from the published
// .proto files: pkg/runtime/generated.proto, pkg/apis/meta/v1/generated.proto, pkg/runtime/serializer/protobuf):
//   body   = 'k' '8' 's' 0x00  Unknown
//   Unknown  { TypeMeta typeMeta = 1; bytes raw = 2; string contentEncoding = 3; string contentType = 4; }
//   raw      = <Kind>List { ListMeta metadata = 1; repeated <Kind> items = 2; }
//   <Kind>   { ObjectMeta metadata = 1; ... }      ObjectMeta { string name = 1; ... string namespace = 3; ... }
// The filter drops `items` entries and rewrites ONE length (that of `raw`); every other byte is preserved.
namespace {

struct Pb {
  const uint8_t* s;
  size_t n, i;
};
inline bool pb_varint(Pb& c, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (c.i >= c.n) return false;
    const uint8_t b = c.s[c.i++];
    r |= static_cast<uint64_t>(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;  // more than 10 bytes
}
// One field: *num / *wt, and for length-delimited fields the payload range [*b, *e). Skips the value.
inline bool pb_field(Pb& c, uint32_t* num, uint32_t* wt, size_t* b, size_t* e, size_t* len_at = nullptr) {
  uint64_t tag, v;
  if (!pb_varint(c, &tag) || (tag >> 3) == 0 || (tag >> 3) > 0x1FFFFFFFull) return false;
  *num = static_cast<uint32_t>(tag >> 3);
  *wt = static_cast<uint32_t>(tag & 7);
  *b = *e = c.i;
  if (len_at) *len_at = c.i;
  switch (*wt) {
    case 0: return pb_varint(c, &v);
    case 1:
      if (c.n - c.i < 8) return false;
      c.i += 8;
      return true;
    case 2:
      if (!pb_varint(c, &v) || v > c.n - c.i) return false;
      *b = c.i;
      *e = c.i + static_cast<size_t>(v);
      c.i = *e;
      return true;
    case 5:
      if (c.n - c.i < 4) return false;
      c.i += 4;
      return true;
    default: return false;  // groups: never emitted by kube's generated marshallers
  }
}
inline size_t pb_varint_len(uint64_t v) {
  size_t n = 1;
  while (v >= 0x80) {
    v >>= 7;
    ++n;
  }
  return n;
}
inline size_t pb_put_varint(char* out, uint64_t v) {
  size_t n = 0;
  while (v >= 0x80) {
    out[n++] = static_cast<char>((v & 0x7F) | 0x80);
    v >>= 7;
  }
  out[n++] = static_cast<char>(v);
  return n;
}
constexpr char kPbMagic[4] = {'k', '8', 's', 0};

// Locates `raw` inside the envelope: *len_pos = offset of its length varint, [*rb, *re) its payload.
// found = false when the Unknown has no raw field (an empty object).
bool pb_envelope(const char* body, size_t len, size_t* len_pos, size_t* rb, size_t* re, bool* found) {
  *found = false;
  if (len < 4 || std::memcmp(body, kPbMagic, 4) != 0) return false;
  Pb c{reinterpret_cast<const uint8_t*>(body), len, 4};
  while (c.i < c.n) {
    uint32_t num, wt;
    size_t b, e, at;
    if (!pb_field(c, &num, &wt, &b, &e, &at)) return false;
    if (num == 2) {
      if (wt != 2) return false;
      // proto3 / gogo semantics for a repeated scalar-bytes field seen twice: last wins. Kube emits it once;
      // a second one would make "the raw the reference filters" ambiguous for a splice: refuse (fail closed).
      if (*found) return false;
      *found = true;
      *rb = b;
      *re = e;
      *len_pos = at;
    }
  }
  return true;
}

// One object message [b, e): its ObjectMeta is field 1 (the last one wins; kube emits one), name = 1, namespace = 3.
bool pb_object_names(const uint8_t* s, size_t b, size_t e, zg_list_item* it) {
  it->flags = ZG_ITEM_IS_OBJECT | ZG_ITEM_RAW_NAMES;
  Pb m{s, e, b};
  while (m.i < m.n) {
    uint32_t fn, fw;
    size_t fb, fe;
    if (!pb_field(m, &fn, &fw, &fb, &fe)) return false;
    if (fn != 1) continue;
    if (fw != 2) return false;
    it->flags |= ZG_ITEM_HAS_METADATA;
    it->name_off = it->ns_off = 0;
    it->name_len = it->ns_len = 0;
    Pb o{s, fe, fb};
    while (o.i < o.n) {
      uint32_t on, ow;
      size_t ob, oe;
      if (!pb_field(o, &on, &ow, &ob, &oe)) return false;
      if (on != 1 && on != 3) continue;
      if (ow != 2 || oe - ob > 0xFFFFFFFFull) return false;
      if (on == 1) {
        it->name_off = ob;
        it->name_len = static_cast<uint32_t>(oe - ob);
      } else {
        it->ns_off = ob;
        it->ns_len = static_cast<uint32_t>(oe - ob);
      }
    }
  }
  return true;
}

// A single protobuf-encoded object (the `default:` branch of the reference's filter, responsefilterer.go:320-341):
// one item = the raw payload itself.
int64_t pb_scan_object(const char* body, size_t len, zg_list_item* out, uint64_t cap, uint64_t* items_begin,
                       uint64_t* items_end) {
  size_t len_pos = 0, rb = 0, re = 0;
  bool found = false;
  if (!pb_envelope(body, len, &len_pos, &rb, &re, &found)) return ZG_EINVAL;
  if (items_begin) *items_begin = found ? len_pos : 0;
  if (items_end) *items_end = found ? re : 0;
  if (!found) return 0;
  zg_list_item it{};
  it.begin = rb;
  it.end = re;
  if (!pb_object_names(reinterpret_cast<const uint8_t*>(body), rb, re, &it)) return ZG_EINVAL;
  if (out && cap < 1) return ZG_E2BIG;
  if (out) out[0] = it;
  return 1;
}

int64_t pb_scan(const char* body, size_t len, zg_list_item* out, uint64_t cap, uint64_t* items_begin, uint64_t* items_end) {
  size_t len_pos = 0, rb = 0, re = 0;
  bool found = false;
  if (!pb_envelope(body, len, &len_pos, &rb, &re, &found)) return ZG_EINVAL;
  if (items_begin) *items_begin = found ? len_pos : 0;
  if (items_end) *items_end = found ? re : 0;
  if (!found) return 0;
  Pb c{reinterpret_cast<const uint8_t*>(body), re, rb};
  int64_t n = 0;
  while (c.i < c.n) {
    const size_t at = c.i;
    uint32_t num, wt;
    size_t b, e;
    if (!pb_field(c, &num, &wt, &b, &e)) return ZG_EINVAL;
    if (num != 2) continue;  // ListMeta (1) and anything newer: kept as they are
    if (wt != 2) return ZG_EINVAL;
    zg_list_item it{};
    it.begin = at;  // the whole entry: tag, length, message
    it.end = e;
    if (!pb_object_names(c.s, b, e, &it)) return ZG_EINVAL;
    if (static_cast<uint64_t>(n) < cap && out) out[n] = it;
    ++n;
  }
  if (out && static_cast<uint64_t>(n) > cap) return ZG_E2BIG;
  return n;
}

int pb_filter(const char* body, size_t len, const zg_list_item* items, uint64_t n, const uint8_t* keep, uint64_t len_pos,
              uint64_t raw_end, char* out, size_t cap, size_t* out_len) {
  if (len_pos < 5 || len_pos >= raw_end || raw_end > len) return ZG_EINVAL;
  Pb c{reinterpret_cast<const uint8_t*>(body), static_cast<size_t>(raw_end), static_cast<size_t>(len_pos)};
  uint64_t raw_len;
  if (!pb_varint(c, &raw_len) || raw_len != raw_end - c.i) return ZG_EINVAL;
  const size_t rb = c.i;
  uint64_t dropped = 0, prev_end = rb;
  for (uint64_t i = 0; i < n; ++i) {  // entries in body order, inside raw, not overlapping
    if (items[i].begin < prev_end || items[i].begin > items[i].end || items[i].end > raw_end) return ZG_EINVAL;
    prev_end = items[i].end;
    if (!keep[i]) dropped += items[i].end - items[i].begin;
  }
  const uint64_t new_raw = raw_len - dropped;
  const size_t need = len_pos + pb_varint_len(new_raw) + static_cast<size_t>(new_raw) + (len - raw_end);
  *out_len = need;
  if (need > cap || !out) return ZG_E2BIG;
  size_t w = 0;
  std::memcpy(out, body, len_pos);
  w = len_pos;
  w += pb_put_varint(out + w, new_raw);
  size_t from = rb;
  for (uint64_t i = 0; i < n; ++i)
    if (!keep[i]) {
      std::memcpy(out + w, body + from, items[i].begin - from);
      w += items[i].begin - from;
      from = items[i].end;
    }
  std::memcpy(out + w, body + from, raw_end - from);
  w += raw_end - from;
  std::memcpy(out + w, body + raw_end, len - raw_end);
  w += len - raw_end;
  *out_len = w;
  return ZG_OK;
}

}  // namespace

extern "C" int64_t zg_list_scan(const char* body, size_t len, uint32_t mode, zg_list_item* out, uint64_t cap,
                                uint64_t* items_begin, uint64_t* items_end) {
  if (!body || mode > ZG_LIST_PROTOBUF_OBJECT) return ZG_EINVAL;
  if (mode == ZG_LIST_PROTOBUF) return pb_scan(body, len, out, cap, items_begin, items_end);
  if (mode == ZG_LIST_PROTOBUF_OBJECT) return pb_scan_object(body, len, out, cap, items_begin, items_end);
  const char* const array_key = mode == ZG_LIST_ITEMS ? "items" : "rows";
  Cur c{body, len, 0, true};
  ws(c);
  int64_t n = 0;
  bool found = false, seen_key = false;
  uint64_t ib = 0, ie = 0;
  const bool ok = walk_object(c, [&](size_t kb, size_t ke) -> bool {
    if (!key_is(c, kb, ke, array_key)) return false;
    // A second top-level "items" ("rows") key: the reference re-marshals a map, so only the LAST array would
    // survive, filtered; a byte splice would ship the earlier array unfiltered to a first-key-wins consumer.
    // Kube never emits this: refuse the body instead of guessing (fail closed).
    if (seen_key) return c.ok = false;
    seen_key = true;
    n = 0;
    found = c.i < c.n && c.s[c.i] == '[';
    if (!found) return false;  // not an array: the reference passes the body through
    ib = c.i;
    ++c.i;
    ws(c);
    if (c.i < c.n && c.s[c.i] == ']') {
      ++c.i;
      ie = c.i;
      return true;
    }
    for (;;) {
      zg_list_item tmp;
      if (!scan_item(c, &tmp, mode)) return c.ok = false;
      if (static_cast<uint64_t>(n) < cap && out) out[n] = tmp;
      ++n;
      ws(c);
      if (c.i >= c.n) return c.ok = false;
      if (c.s[c.i] == ',') {
        ++c.i;
        continue;
      }
      if (c.s[c.i] == ']') {
        ++c.i;
        ie = c.i;
        return true;
      }
      return c.ok = false;
    }
  });
  if (!ok || !c.ok) return ZG_EINVAL;
  ws(c);
  if (c.i != c.n) return ZG_EINVAL;  // trailing bytes after the document
  if (!found) ib = ie = 0;
  if (items_begin) *items_begin = ib;
  if (items_end) *items_end = ie;
  if (!found) return 0;  // no "items" key: nothing to filter (postfilter.go:27-31 returns the body unchanged)
  if (out && static_cast<uint64_t>(n) > cap) return ZG_E2BIG;
  return n;
}

extern "C" int zg_list_filter(const char* body, size_t len, const zg_list_item* items, uint64_t n, const uint8_t* keep,
                              uint64_t items_begin, uint64_t items_end, uint32_t flags, char* out, size_t cap,
                              size_t* out_len) {
  if (!body || (!items && n) || (!keep && n) || !out_len) return ZG_EINVAL;
  if (items_begin > items_end || items_end > len) return ZG_EINVAL;
  if (len >= 4 && std::memcmp(body, kPbMagic, 4) == 0)  // a protobuf body (no JSON document starts with 'k')
    return pb_filter(body, len, items, n, keep, items_begin, items_end, out, cap, out_len);
  size_t need = items_begin + (len - items_end);
  uint64_t kept = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (keep[i]) {
      if (items[i].begin > items[i].end || items[i].end > len) return ZG_EINVAL;
      need += (items[i].end - items[i].begin) + (kept ? 1 : 0);
      ++kept;
    }
  // Nothing kept: the post-filter marshals a nil slice, i.e. "items":null (postfilter.go:138,
  // `var allowedItems []interface{}` is never appended to) -- ZG_LIST_EMPTY_AS_NULL; the
  // pre-filter's list/table paths start from make(..., 0) and give [] (responsefilterer.go:356,377).
  const bool as_null = !kept && (flags & ZG_LIST_EMPTY_AS_NULL);
  need += as_null ? 4 : 2;
  *out_len = need;
  if (need > cap || !out) return ZG_E2BIG;
  size_t w = 0;
  std::memcpy(out + w, body, items_begin);
  w += items_begin;
  if (as_null) {
    std::memcpy(out + w, "null", 4);
    w += 4;
  } else {
    out[w++] = '[';
    kept = 0;
    for (uint64_t i = 0; i < n; ++i)
      if (keep[i]) {
        if (kept++) out[w++] = ',';
        std::memcpy(out + w, body + items[i].begin, items[i].end - items[i].begin);
        w += items[i].end - items[i].begin;
      }
    out[w++] = ']';
  }
  std::memcpy(out + w, body + items_end, len - items_end);
  w += len - items_end;
  *out_len = w;
  return ZG_OK;
}
