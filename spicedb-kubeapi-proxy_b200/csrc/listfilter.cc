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

// Skips a string starting at the opening quote; returns [begin, end) of its raw contents.
inline bool skip_string(Cur& c, size_t* b, size_t* e) {
  if (c.i >= c.n || c.s[c.i] != '"') return c.ok = false;
  size_t i = c.i + 1;
  *b = i;
  while (i < c.n) {
#if defined(__SSE2__)
    // 16 bytes at a time up to the next quote, backslash or control character (most of a kube
    // body is string contents)
    while (i + 16 <= c.n) {
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c.s + i));
      const __m128i special = _mm_or_si128(
          _mm_or_si128(_mm_cmpeq_epi8(v, _mm_set1_epi8('"')), _mm_cmpeq_epi8(v, _mm_set1_epi8('\\'))),
          _mm_cmpeq_epi8(_mm_max_epu8(v, _mm_set1_epi8(0x1f)), _mm_set1_epi8(0x1f)));  // byte <= 0x1f
      const int m = _mm_movemask_epi8(special);
      if (m) {
        i += __builtin_ctz(m);
        break;
      }
      i += 16;
    }
    if (i >= c.n) break;
#endif
    const char ch = c.s[i];
    if (ch == '"') {
      *e = i;
      c.i = i + 1;
      return true;
    }
    if (ch == '\\') {
      if (i + 1 >= c.n) break;
      const char x = c.s[i + 1];
      if (x == 'u') {
        if (i + 6 > c.n || hexv(c.s[i + 2]) < 0 || hexv(c.s[i + 3]) < 0 || hexv(c.s[i + 4]) < 0 || hexv(c.s[i + 5]) < 0)
          break;
        i += 6;
        continue;
      }
      if (x != '"' && x != '\\' && x != '/' && x != 'b' && x != 'f' && x != 'n' && x != 'r' && x != 't') break;
      i += 2;
      continue;
    }
    if (static_cast<unsigned char>(ch) < 0x20) break;  // raw control characters are not JSON
    ++i;
  }
  return c.ok = false;
}

bool skip_value(Cur& c, int depth);

bool skip_container(Cur& c, char open, char close, int depth) {
  if (depth > 512) return c.ok = false;
  ++c.i;  // opening bracket
  ws(c);
  if (c.i < c.n && c.s[c.i] == close) {
    ++c.i;
    return true;
  }
  for (;;) {
    ws(c);
    if (open == '{') {
      size_t b, e;
      if (!skip_string(c, &b, &e)) return false;
      ws(c);
      if (c.i >= c.n || c.s[c.i] != ':') return c.ok = false;
      ++c.i;
    }
    if (!skip_value(c, depth + 1)) return false;
    ws(c);
    if (c.i >= c.n) return c.ok = false;
    if (c.s[c.i] == ',') {
      ++c.i;
      continue;
    }
    if (c.s[c.i] == close) {
      ++c.i;
      return true;
    }
    return c.ok = false;
  }
}

bool skip_value(Cur& c, int depth) {
  ws(c);
  if (c.i >= c.n) return c.ok = false;
  const char ch = c.s[c.i];
  if (ch == '{') return skip_container(c, '{', '}', depth);
  if (ch == '[') return skip_container(c, '[', ']', depth);
  if (ch == '"') {
    size_t b, e;
    return skip_string(c, &b, &e);
  }
  for (const char* lit : {"true", "false", "null"}) {
    const size_t n = std::strlen(lit);
    if (ch == lit[0]) {
      if (c.i + n > c.n || std::memcmp(c.s + c.i, lit, n) != 0) return c.ok = false;
      c.i += n;
      return true;
    }
  }
  // number: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?
  size_t i = c.i;
  auto digits = [&]() {
    const size_t s0 = i;
    while (i < c.n && c.s[i] >= '0' && c.s[i] <= '9') ++i;
    return i - s0;
  };
  if (i < c.n && c.s[i] == '-') ++i;
  if (i < c.n && c.s[i] == '0') ++i;
  else if (!digits()) return c.ok = false;
  if (i < c.n && c.s[i] == '.') {
    ++i;
    if (!digits()) return c.ok = false;
  }
  if (i < c.n && (c.s[i] == 'e' || c.s[i] == 'E')) {
    ++i;
    if (i < c.n && (c.s[i] == '+' || c.s[i] == '-')) ++i;
    if (!digits()) return c.ok = false;
  }
  c.i = i;
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

extern "C" int64_t zg_list_scan(const char* body, size_t len, uint32_t mode, zg_list_item* out, uint64_t cap,
                                uint64_t* items_begin, uint64_t* items_end) {
  if (!body || mode > ZG_LIST_TABLE_ROWS) return ZG_EINVAL;
  const char* const array_key = mode == ZG_LIST_ITEMS ? "items" : "rows";
  Cur c{body, len, 0, true};
  ws(c);
  int64_t n = 0;
  bool found = false;
  uint64_t ib = 0, ie = 0;
  const bool ok = walk_object(c, [&](size_t kb, size_t ke) -> bool {
    if (!key_is(c, kb, ke, array_key)) return false;
    n = 0;  // a later duplicate "items" key replaces an earlier one
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
