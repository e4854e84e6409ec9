"""List-response filter: zg_list_scan / zg_list_filter and the host mirror of
pkg/authz/postfilter.go. The checker for the byte-level work is Python's json module (decode both
sides, compare values); the reference's own cases (postfilter_test.go:151-323) are transcribed with
a mock client shaped like its mockPermissionsClient (postfilter_test.go:19-66)."""
import json
import random

import numpy as np
import pytest

import zgpu  # noqa: F401  (registers the package under an importable name)
from spicedb_kubeapi_proxy_b200 import _lib, client as cl, postfilter as pf


class MockPermissionsClient:
    """postfilter_test.go:19-66: a map from 'type:id#perm@stype:sid' to a permissionship, default NO."""

    def __init__(self, responses, errors=(), fail=False, short=False):
        self.responses, self.errors, self.fail, self.short = responses, set(errors), fail, short
        self.calls = []

    def CheckBulkPermissions(self, req):
        if self.fail:
            raise cl.RpcError("UNAVAILABLE", "down")
        self.calls.append(len(req.items))
        pairs = []
        for it in req.items:
            key = f"{it.resource.object_type}:{it.resource.object_id}#{it.permission}@" \
                  f"{it.subject.object.object_type}:{it.subject.object.object_id}"
            if key in self.errors:
                pairs.append(cl.CheckBulkPermissionsPair(it, error="boom"))
            else:
                pairs.append(cl.CheckBulkPermissionsPair(it, item=cl.CheckBulkPermissionsResponseItem(
                    self.responses.get(key, cl.PERMISSIONSHIP_NO_PERMISSION))))
        return cl.CheckBulkPermissionsResponse(pairs[:-1] if self.short else pairs)


HAS, NO = cl.PERMISSIONSHIP_HAS_PERMISSION, cl.PERMISSIONSHIP_NO_PERMISSION
TPL = "pod:{{name}}#view@user:{{user.name}}"
REQ, USER = pf.RequestInfo(verb="list"), pf.UserInfo(name="testuser")


def pod(name, ns="default", **extra):
    return {"metadata": {"name": name, "namespace": ns}, **extra}


def test_reference_filter_list_response():
    # postfilter_test.go:151-244 TestFilterListResponse
    body = json.dumps({"apiVersion": "v1", "kind": "PodList", "items": [pod("pod1"), pod("pod2")]}).encode()
    mock = MockPermissionsClient({"pod:pod1#view@user:testuser": HAS, "pod:pod2#view@user:testuser": NO})
    out = json.loads(pf.filter_list_response(body, [TPL], REQ, USER, mock))
    assert len(out["items"]) == 1 and out["items"][0]["metadata"]["name"] == "pod1"
    assert out["apiVersion"] == "v1" and out["kind"] == "PodList"
    assert mock.calls == [2]  # ONE bulk call for the whole list


def test_reference_filter_items_with_bulk_permissions():
    # postfilter_test.go:246-323 TestFilterItemsWithBulkPermissions (incl. the empty-items leg)
    mock = MockPermissionsClient({"pod:testpod1#view@user:testuser": HAS, "pod:testpod2#view@user:testuser": NO})
    body = json.dumps({"items": [pod("testpod1"), pod("testpod2")]}).encode()
    out = json.loads(pf.filter_list_response(body, [TPL], REQ, USER, mock))
    assert [i["metadata"]["name"] for i in out["items"]] == ["testpod1"]
    empty = b'{"items": []}'
    assert pf.filter_list_response(empty, [TPL], REQ, USER, mock) == empty
    assert mock.calls == [2]  # no call for an empty list


def test_mirrored_decisions():
    mock = MockPermissionsClient({"pod:a#view@user:testuser": HAS, "pod:a#edit@user:testuser": HAS,
                                  "pod:b#view@user:testuser": HAS, "pod:req#view@user:testuser": HAS},
                                 errors={"pod:e#view@user:testuser"})
    # no "items" array / items of another type: passthrough, byte for byte
    for raw in (b'{"kind":"Status","code":403}', b'{"items":"x"}', b'{"items":{"a":1}}', b' {"items" : null} '):
        assert pf.filter_list_response(raw, [TPL], REQ, USER, mock) is raw
    # non-object items are kept; an item without metadata takes the request's name (rules.go:321-326)
    body = json.dumps({"items": [3, "s", None, [pod("zz")], pod("a"), pod("nope"), {"spec": 1}]}).encode()
    out = json.loads(pf.filter_list_response(body, [TPL], pf.RequestInfo(name="req"), USER, mock))
    assert out["items"] == [3, "s", None, [pod("zz")], pod("a"), {"spec": 1}]
    # all post-filters must pass; per-pair error drops the item; unresolvable template = check skipped
    body = json.dumps({"items": [pod("a"), pod("b"), pod("e")]}).encode()
    both = [TPL, "pod:{{name}}#edit@user:{{user.name}}"]
    assert [i["metadata"]["name"] for i in json.loads(pf.filter_list_response(body, both, REQ, USER, mock))["items"]] == ["a"]
    assert [i["metadata"]["name"] for i in json.loads(pf.filter_list_response(body, [TPL], REQ, USER, mock))["items"]] == ["a", "b"]
    skipped = ["pod:{{nosuchfield}}#view@user:{{user.name}}"]
    assert json.loads(pf.filter_list_response(body, skipped, REQ, USER, mock))["items"] == [pod("a"), pod("b"), pod("e")]
    # nothing kept: "items": null, as a re-marshalled nil slice (postfilter.go:138)
    none = pf.filter_list_response(json.dumps({"items": [pod("x")], "kind": "L"}).encode(), [TPL], REQ, USER, mock)
    assert json.loads(none) == {"items": None, "kind": "L"}
    # a short response drops the items whose pair is missing (postfilter.go:151-155)
    short = MockPermissionsClient({"pod:a#view@user:testuser": HAS, "pod:b#view@user:testuser": HAS}, short=True)
    body = json.dumps({"items": [pod("a"), pod("b")]}).encode()
    assert json.loads(pf.filter_list_response(body, [TPL], REQ, USER, short))["items"] == [pod("a")]
    # the bulk call failing fails the filter; a malformed body too
    with pytest.raises(cl.RpcError):
        pf.filter_list_response(body, [TPL], REQ, USER, MockPermissionsClient({}, fail=True))
    for bad in (b'{"items":[', b'{"items":[{]}', b'{"items":[1,]}', b'{"items":[]} x', b'', b'[1]', b'{"a":"\x01"}',
                b'{"a":tru}', b'{"a":nulll}', b'{"a":01}', b'{"a":1.}', b'{"a":-}', b'{"a":1e}', b'{"a":.5}',
                b'{"a":"\\x"}', b'{"a":"\\u12g4"}', b'{"a":1 "b":2}', b'{"a"}', b'{a:1}', b'{"items":[1 2]}'):
        with pytest.raises(ValueError):
            pf.filter_list_response(bad, [TPL], REQ, USER, mock)


def test_namespaced_name_normalisation():
    # pkg/rules/rules.go:312-339: ns/name, request fallbacks, `namespaces` clears the namespace
    seen = []

    def probe(fields):
        seen.append((fields["name"], fields["namespace"], fields["namespacedName"]))
        return f"pod:{fields['namespacedName']}#view@user:u"
    mock = MockPermissionsClient({})
    body = json.dumps({"items": [pod("p", "ns1"), {"metadata": {"name": "q"}}, {"metadata": {"namespace": "ns2"}},
                                 {"metadata": {"name": 5, "namespace": ["x"]}}]}).encode()
    pf.filter_list_response(body, [probe], pf.RequestInfo(name="rn", namespace="rns"), USER, mock)
    assert seen == [("p", "ns1", "ns1/p"), ("q", "rns", "rns/q"), ("rn", "ns2", "ns2/rn"), ("rn", "rns", "rns/rn")]
    seen.clear()
    pf.filter_list_response(json.dumps({"items": [pod("team-a", "team-a")]}).encode(), [probe],
                            pf.RequestInfo(resource="namespaces"), USER, mock)
    assert seen == [("team-a", "", "team-a")]


def test_duplicate_and_escaped_keys_follow_encoding_json():
    # encoding/json: the last duplicate key wins; keys are compared after unescaping
    raw = (b'{"x":1,'
           b'"\\u0069tems":[{"metadata":{"name":"first","name":"n\\u0061me2"},"\\u006detadata":{"n\\u0061me":"last"}},'
           b'{"metadata":{"name":"gone"},"metadata":7}]}')
    # a second top-level "items" key is refused: a splice would ship the earlier array unfiltered, while the
    # reference (re-marshalling a map) keeps only the last one -- kube never emits this, so fail closed
    dup = b'{"items":[{"metadata":{"name":"old"}}],' + raw[1:]
    with pytest.raises(_lib.ZgpuError):
        _lib.list_scan(dup)
    items, ib, ie = _lib.list_scan(raw)
    got = [(raw[i["name_off"]:i["name_off"] + i["name_len"]], int(i["flags"])) for i in items]
    assert got == [(b"last", 3), (b"", 1)]
    ref = json.loads(raw)  # Python's json keeps the last duplicate too
    assert [i["metadata"]["name"] if isinstance(i["metadata"], dict) else None for i in ref["items"]] == ["last", None]
    assert raw[ib:ie] == raw[raw.index(b"[", raw.index(b"\\u0069tems")):-1]
    # escaped VALUES are handed over raw and decoded by the caller
    seen = []
    pf.filter_list_response(b'{"items":[{"metadata":{"name":"a\\u00e9\\"b","namespace":"n\\\\s"}}]}',
                            [lambda f: seen.append(f["namespacedName"]) or "pod:x#view@user:u"], REQ, USER,
                            MockPermissionsClient({}))
    assert seen == ['n\\s/aé"b']


def _rand_value(rng, depth=0):
    k = rng.randrange(9 if depth < 4 else 6)
    if k == 0:
        return rng.choice([None, True, False])
    if k == 1:
        return rng.choice([0, -1, 17, 2**53, 1.5, -2.25e-7, 1e300, 12345678901234567890])
    if k in (2, 3, 4, 5):
        alphabet = ['a', 'Z', '0', ' ', '"', '\\', '/', '{', '}', '[', ']', ',', ':', '\n', '\t', 'é', '中',
                    '\U0001f600', 'items', 'metadata', 'name']
        return "".join(rng.choice(alphabet) for _ in range(rng.randrange(8)))
    if k in (6, 7):
        return {str(_rand_value(rng, 9)) if rng.random() < 0.3 else rng.choice(["a", "name", "metadata", "items", "spec"]):
                _rand_value(rng, depth + 1) for _ in range(rng.randrange(4))}
    return [_rand_value(rng, depth + 1) for _ in range(rng.randrange(4))]


def _rand_item(rng):
    r = rng.random()
    if r < 0.1:
        return _rand_value(rng, 2)
    meta = {}
    if rng.random() < 0.9:
        meta["name"] = rng.choice(["pod-%d" % rng.randrange(50), _rand_value(rng, 9), "né\"\\x"])
    if rng.random() < 0.7:
        meta["namespace"] = rng.choice(["ns-%d" % rng.randrange(5), _rand_value(rng, 9)])
    meta["labels"] = _rand_value(rng, 2)
    item = {"spec": _rand_value(rng, 1), "metadata": meta if rng.random() < 0.9 else _rand_value(rng, 3),
            "status": _rand_value(rng, 1)}
    keys = list(item)
    rng.shuffle(keys)
    return {k: item[k] for k in keys}


@pytest.mark.parametrize("seed", range(40))
def test_scan_and_filter_against_json_module(seed):
    rng = random.Random(seed)
    doc = {"kind": "PodList", "apiVersion": "v1", "metadata": {"resourceVersion": "42", "name": "decoy"}}
    items = json.loads(json.dumps([_rand_item(rng) for _ in range(rng.randrange(0, 40))]))  # keys -> str
    pos = rng.randrange(len(doc) + 1)
    doc = dict(list(doc.items())[:pos] + [("items", items)] + list(doc.items())[pos:])
    style = rng.randrange(3)
    body = json.dumps(doc, ensure_ascii=rng.random() < 0.5, indent=[None, 2, None][style],
                      separators=[None, None, (",", ":")][style]).encode()
    got, ib, ie = _lib.list_scan(body)
    assert len(got) == len(items) and json.loads(body[ib:ie]) == items
    for g, it in zip(got, items):
        assert json.loads(body[g["begin"]:g["end"]]) == it
        is_obj = isinstance(it, dict)
        meta = it.get("metadata") if is_obj else None
        assert int(g["flags"]) == (1 if is_obj else 0) | (2 if isinstance(meta, dict) else 0)
        for key, off, ln in (("name", "name_off", "name_len"), ("namespace", "ns_off", "ns_len")):
            want = meta.get(key) if isinstance(meta, dict) and isinstance(meta.get(key), str) else ""
            assert json.loads(b'"' + body[g[off]:g[off] + g[ln]] + b'"') == want
    for _ in range(3):
        keep = np.array([rng.random() < 0.5 for _ in items], dtype=np.uint8)
        as_null = rng.random() < 0.5
        out = json.loads(_lib.list_filter(body, got, keep, ib, ie, _lib.LIST_EMPTY_AS_NULL if as_null else 0))
        want = dict(doc)
        want["items"] = [it for it, k in zip(items, keep) if k] or (None if as_null else [])
        assert out == want and list(out) == list(want)  # values AND key order preserved


def test_filter_buffer_contract():
    import ctypes as C
    body = b'{"items":[{"a":1},{"b":2}],"k":1}'
    items, ib, ie = _lib.list_scan(body)
    L = _lib.lib()
    need = C.c_size_t(0)
    keep = np.array([1, 0], dtype=np.uint8)
    rc = L.zg_list_filter(body, len(body), items.ctypes.data, 2, keep.ctypes.data, ib, ie, 0, None, 0, C.byref(need))
    assert rc == -7 and need.value == len(b'{"items":[{"a":1}],"k":1}')
    buf = C.create_string_buffer(need.value)
    assert L.zg_list_filter(body, len(body), items.ctypes.data, 2, keep.ctypes.data, ib, ie, 0, buf, need.value,
                            C.byref(need)) == 0
    assert buf.raw == b'{"items":[{"a":1}],"k":1}'
    # capacity too small for the scan: E2BIG, and a count-only call (out = NULL) reports the size
    small = np.zeros(1, dtype=_lib.LIST_ITEM_DTYPE)
    assert L.zg_list_scan(body, len(body), 0, small.ctypes.data, 1, None, None) == -7
    assert L.zg_list_scan(body, len(body), 0, None, 0, None, None) == 2
    assert L.zg_list_scan(body, len(body), 2, None, 0, None, None) == -1  # unknown mode
    # deep nesting is bounded, not a stack overflow
    deep = b'{"items":[' + b'[' * 100000 + b']' * 100000 + b']}'
    assert L.zg_list_scan(deep, len(deep), 0, None, 0, None, None) == -1


def test_list_resolve_matches_the_string_path():
    """zg_list_resolve builds the checks of a template straight from the body bytes; the checker is the string
    path (postfilter._fields -> template -> zg_resolve_checks) on the same items."""
    from spicedb_kubeapi_proxy_b200 import workloads
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    c = cl.PermissionsClient(workloads.BOOTSTRAP_SCHEMA, engine=e)
    c.WriteRelationships(cl.WriteRelationshipsRequest(
        [cl.RelationshipUpdate(cl.OPERATION_TOUCH, cl.Relationship.parse(r)) for r in
         [f"pod:ns{i % 3}/p{i}#viewer@user:alice" for i in range(20)] + ["pod:solo#viewer@user:alice", "pod:rn#viewer@user:bob",
                                                                        "pod:rns/rn#viewer@user:bob", "namespace:ns1#creator@user:alice"]]))
    rng = random.Random(11)
    items = []
    for i in range(300):
        r = rng.random()
        if r < 0.08:
            items.append(rng.choice([1, "s", None, [1], True]))
            continue
        meta = {}
        if rng.random() < 0.85:
            meta["name"] = rng.choice([f"p{rng.randrange(25)}", "solo", "", 'q"uo\\te', "é中\U0001f600", "tab\tnl\n", 7])
        if rng.random() < 0.7:
            meta["namespace"] = rng.choice([f"ns{rng.randrange(4)}", "", None, "n/s"])
        items.append({"spec": {"metadata": {"name": "decoy"}}, "metadata": meta} if rng.random() < 0.9 else {"kind": "NoMeta"})
    for ensure_ascii in (True, False):
        body = json.dumps({"items": items}, ensure_ascii=ensure_ascii).encode()
        scanned, ib, ie = _lib.list_scan(body)
        for kind, tplstr in ((_lib.ID_NAMESPACED_NAME, "pod:{{namespacedName}}#view@user:alice"), (_lib.ID_NAME, "pod:{{name}}#view@user:alice")):
            for req, clear in ((pf.RequestInfo(), False), (pf.RequestInfo(name="rn", namespace="rns"), False),
                               (pf.RequestInfo(resource="namespaces", namespace="rns"), True)):
                tpl = e.list_template("pod", "view", "user", "alice", "", kind, req.name, req.namespace, clear)
                got, checked = e.list_resolve(body, scanned, tpl)
                want_rels, want_checked = [], []
                for it in scanned:
                    flags = int(it["flags"])
                    rel = None
                    if flags & _lib.ITEM_IS_OBJECT:
                        f = pf._fields(req, pf.UserInfo(name="alice"), bool(flags & _lib.ITEM_HAS_METADATA),
                                       pf._text(body, int(it["name_off"]), int(it["name_len"])),
                                       pf._text(body, int(it["ns_off"]), int(it["ns_len"])))
                        try:
                            rel = pf.resolve_rel(tplstr, f)
                        except pf.ResolveError:
                            rel = None
                    want_checked.append(rel is not None)
                    if rel is not None:
                        want_rels.append(rel)
                assert checked.astype(bool).tolist() == want_checked
                assert got[checked.astype(bool)].tobytes() == e.resolve_checks(want_rels).tobytes()
                assert 20 < sum(want_checked) < 300 and any(int(g["res"]) != 0xFFFFFFFF for g in got[checked.astype(bool)])
    # unknown permission / type: checked, and the item answers ZG_ITEM_ERROR like the string path (perm = 0xFFFF)
    got, checked = e.list_resolve(body, scanned, e.list_template("pod", "nosuch", "user", "alice"))
    assert checked.any() and (got["perm"] == 0xFFFF).all()
    # never-written resource and subject that are the same object share the sentinel
    b2 = b'{"items":[{"metadata":{"name":"ghost"}}]}'
    sc2 = _lib.list_scan(b2)[0]
    g2, _ = e.list_resolve(b2, sc2, e.list_template("pod", "view", "pod", "ghost", "viewer", _lib.ID_NAME))
    assert (int(g2[0]["res"]), int(g2[0]["subj"])) == (0xFFFFFFFF, 0xFFFFFFFF)
    g3, _ = e.list_resolve(b2, sc2, e.list_template("pod", "view", "pod", "other", "viewer", _lib.ID_NAME))
    assert (int(g3[0]["res"]), int(g3[0]["subj"])) == (0xFFFFFFFF, 0xFFFFFFFE)
    # item ranges handed back by the caller are bounds-checked (overflow-safe) before any byte is read
    for off, ln in ((10**9, 1), (2**64 - 2, 5), (len(b2) - 1, 2)):
        bad = sc2.copy()
        bad["name_off"], bad["name_len"] = off, ln
        with pytest.raises(_lib.ZgpuError, match="outside the body"):
            e.list_resolve(b2, bad, e.list_template("pod", "view", "user", "u", "", _lib.ID_NAME))
    # the fused call: passthrough cases need no GPU; anything with checks fails loudly without one
    tpl = e.list_template("pod", "view", "user", "alice")
    for raw in (b'{"kind":"Status"}', b'{"items":[]}', b'{"items":null}'):
        assert e.list_postfilter(raw, [tpl]) == raw
    assert json.loads(e.list_postfilter(b'{"items":[1,"x"]}', [tpl])) == {"items": [1, "x"]}  # nothing to check
    tiny = json.dumps({"items": [1] * 5000, "k": "v"}, separators=(",", ":")).encode()  # more items than the first guess
    assert json.loads(e.list_postfilter(tiny, [tpl])) == json.loads(tiny)
    with pytest.raises(_lib.ZgpuError, match="no CPU fallback"):
        e.list_postfilter(body, [tpl])
    with pytest.raises(_lib.ZgpuError):
        e.list_postfilter(b'{"items":[', [tpl])


def test_scanner_fuzz_under_sanitizers(tmp_path):
    """tests/fuzz/listfilter_fuzz.cc: mutated bodies through zg_list_scan / zg_list_filter under ASan + UBSan
    (exact-size heap buffers, so any over-read is a report), with range and length invariants checked."""
    import os
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "lf_fuzz")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            "-o", exe, os.path.join(root, "tests", "fuzz", "listfilter_fuzz.cc"),
                            os.path.join(root, "spicedb-kubeapi-proxy_b200", "csrc", "listfilter.cc")],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, "400000"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and run.stdout.startswith("ok "), (run.stdout[-500:], run.stderr[-2000:])


def test_scanner_accepts_exactly_what_a_json_parser_accepts():
    """Differential against Python's json on mutated documents: the scanner is a strict validator, so it must
    reject what a JSON parser rejects (it sits in front of one) and accept what it accepts. Known, deliberate
    differences are normalised: Python accepts NaN/Infinity literals and lone surrogate escapes; RFC 8259 (and
    encoding/json) do not."""
    import ctypes as C
    L = _lib.lib()
    rng = random.Random(2024)
    seeds = [json.dumps({"kind": "L", "items": [pod("a", "n", spec={"x": [1, 2.5e-3, True, None, {"y": "]}"}]}), pod("b"), 3, "s"],
                         "metadata": {"k": "v\u00e9\"\\"}}).encode(),
             b' { "items" : [ { "metadata" : { "name" : "x" } } , [ ] , { } ] , "n" : -0.5E+3 } ',
             b'{"a":[[[[]]]],"items":[{"metadata":{"name":"q","namespace":"w"},"z":{"a":{"b":[1,{"c":"d"}]}}}]}']
    alphabet = b'{}[]",:\\ \n0123456789-+.eEtrufalsn\x00\x7f\xc3\xa9'
    agree = rejected = 0
    for _ in range(6000):
        b = bytearray(rng.choice(seeds))
        for _m in range(rng.randrange(1, 4)):
            p = rng.randrange(len(b)) if b else 0
            k = rng.randrange(4)
            if k == 0 and b:
                b[p] = rng.choice(alphabet)
            elif k == 1 and b:
                del b[p:p + rng.randrange(1, 3)]
            elif k == 2:
                b.insert(p, rng.choice(alphabet))
            elif b:
                del b[p:]
        raw = bytes(b)
        try:
            doc = json.loads(raw.decode("utf-8"), parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
            ok_py = isinstance(doc, dict)  # the scanner only takes a top-level object (a List is one)
        except (ValueError, UnicodeDecodeError, RecursionError):
            ok_py = None
        if ok_py is None:
            try:  # invalid UTF-8 is not the scanner's business (encoding/json replaces it): retry leniently
                doc = json.loads(raw.decode("utf-8", "replace"), parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
                ok_py = isinstance(doc, dict) if b"\xef\xbf\xbd" not in raw else None
            except (ValueError, RecursionError):
                ok_py = False
        if ok_py is None:
            continue
        n = L.zg_list_scan(raw, len(raw), 0, None, 0, None, None)
        assert (n >= 0) == ok_py, (raw, n, ok_py)
        agree += 1
        rejected += not ok_py
        if ok_py and isinstance(doc.get("items"), list):
            assert n == len(doc["items"]), raw
    assert agree > 5000 and 500 < rejected < agree - 500


class MockLookupClient:
    def __init__(self, ids, conditional=()):
        self.ids, self.conditional = ids, set(conditional)

    def LookupResources(self, req):
        self.req = req
        return iter([cl.LookupResourcesResponse(i, 2 if i in self.conditional else cl.LOOKUP_PERMISSIONSHIP_HAS_PERMISSION)
                     for i in self.ids])


def test_prefilter_lookup_to_allowed_set():
    # lookups.go:44-132
    mock = MockLookupClient(["ns1/a", "ns2/b", "clusterwide", "ns3/cond"], conditional={"ns3/cond"})
    rel = ("pod", "$", "view", "user", "alice", "")
    res = pf.run_lookup_resources(mock, rel, pf.RequestInfo(namespace="reqns"))
    assert res.allowed_results == {("ns1", "a"), ("ns2", "b"), ("reqns", "clusterwide")}
    assert (mock.req.resource_object_type, mock.req.permission, mock.req.subject.object.object_id) == ("pod", "view", "alice")
    assert res.IsAllowed("ns1", "a") and not res.IsAllowed("ns1", "b") and pf.PrefilterResult(all_allowed=True).IsAllowed("x", "y")
    with pytest.raises(ValueError):
        pf.run_lookup_resources(mock, ("pod", "p1", "view", "user", "alice", ""), pf.RequestInfo())
    with pytest.raises(ValueError):  # a name expression that yields nothing fails the whole pre-filter
        pf.run_lookup_resources(MockLookupClient(["ns/"]), rel, pf.RequestInfo())


def test_prefilter_list_table_object():
    res = pf.PrefilterResult(allowed_results={("ns1", "a"), ("", "node1")})
    body = json.dumps({"kind": "PodList", "items": [pod("a", "ns1"), pod("a", "ns2"), {"metadata": {"name": "node1"}},
                                                    {"spec": {}}], "metadata": {}}).encode()
    out = json.loads(pf.filter_list(body, res))
    assert out == {"kind": "PodList", "items": [pod("a", "ns1"), {"metadata": {"name": "node1"}}], "metadata": {}}
    # nothing allowed: [] (make(..., 0), responsefilterer.go:377), not the post-filter's null
    assert json.loads(pf.filter_list(body, pf.PrefilterResult()))["items"] == []
    assert json.loads(pf.filter_list(body, pf.PrefilterResult(all_allowed=True))) == json.loads(body)
    with pytest.raises(ValueError):
        pf.filter_list(b'{"items":[1]}', res)
    # metav1.Table: rows[i].object carries the PartialObjectMetadata (responsefilterer.go:349-374)
    row = lambda n, ns, **kw: {"cells": [n, "Running", "metadata"], "object": {"kind": "PartialObjectMetadata", **pod(n, ns)}, **kw}
    table = {"kind": "Table", "columnDefinitions": [{"name": "Name", "type": "string"}],
             "rows": [row("a", "ns1"), row("a", "ns2", metadata={"name": "decoy"}), row("b", "ns1"),
                      {"cells": [], "object": None}]}
    tb = json.dumps(table).encode()
    out = json.loads(pf.filter_table(tb, res))
    assert out["rows"] == [row("a", "ns1")] and out["columnDefinitions"] == table["columnDefinitions"]
    assert json.loads(pf.filter_table(tb, pf.PrefilterResult()))["rows"] == []
    with pytest.raises(ValueError):  # a row without "object": empty RawExtension does not decode
        pf.filter_table(json.dumps({"rows": [{"cells": []}]}).encode(), res)
    # a row-level "metadata" is not the object's: only rows[i].object.metadata counts
    items, _, _ = _lib.list_scan(json.dumps({"rows": [{"metadata": {"name": "x"}, "object": {"metadata": {"name": "y"}}}]}).encode(),
                                 _lib.LIST_TABLE_ROWS)
    assert int(items[0]["name_len"]) == 1 and int(items[0]["flags"]) == 7
    # single object (responsefilterer.go:403-415)
    one = json.dumps(pod("a", "ns1", spec={"x": [1, 2]})).encode()
    assert pf.filter_object(one, res) is one
    with pytest.raises(pf.Unauthorized):
        pf.filter_object(json.dumps(pod("zz", "ns1")).encode(), res)
    with pytest.raises(ValueError):
        pf.filter_object(b'{"a":1},{"b":2}', res)


def test_list_keep_allowed_matches_the_prefilter_mirror():
    """zg_list_keep_allowed (ids from LookupResources -> keep mask, in C) against the mirror of lookups.go +
    responsefilterer.go (PrefilterResult built from the same ids by name, filter_list / filter_table)."""
    from spicedb_kubeapi_proxy_b200 import workloads
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    names = [f"ns{i % 3}/p{i}" for i in range(30)] + ["cw1", "cw2", "rns/inreq", "/odd"]
    ids = {n: e.intern("pod", n) for n in names}
    rng = random.Random(5)
    kept_total = seen_total = 0
    for trial in range(30):
        allowed_names = sorted(rng.sample(names, rng.randrange(0, len(names))))
        allowed = np.array(sorted(ids[n] for n in allowed_names), dtype=np.uint32)
        req_ns = rng.choice(["", "rns", "ns1"])
        result = pf.run_lookup_resources(MockLookupClient(allowed_names), ("pod", "$", "view", "user", "u", ""),
                                         pf.RequestInfo(namespace=req_ns))
        objs = []
        for _ in range(60):
            meta = {}
            if rng.random() < 0.9:
                meta["name"] = rng.choice([f"p{rng.randrange(34)}", "cw1", "cw2", "inreq", "odd", "a/b", "", "never"])
            if rng.random() < 0.8:
                meta["namespace"] = rng.choice([f"ns{rng.randrange(4)}", "rns", "", "ns1"])
            objs.append({"metadata": meta, "spec": {"x": 1}} if rng.random() < 0.95 else {"spec": 2})
        for mode, body, mirror in (
                (_lib.LIST_ITEMS, json.dumps({"kind": "PodList", "items": objs}).encode(), pf.filter_list),
                (_lib.LIST_TABLE_ROWS, json.dumps({"kind": "Table", "rows": [{"cells": [1], "object": o} for o in objs]}).encode(),
                 pf.filter_table)):
            scanned, ib, ie = _lib.list_scan(body, mode)
            keep = e.list_keep_allowed(body, scanned, "pod", allowed, mode, req_ns)
            got = _lib.list_filter(body, scanned, keep, ib, ie)
            assert got == mirror(body, result), (trial, mode)
        kept_total += int(keep.sum())
        seen_total += len(objs)
    assert 100 < kept_total < seen_total - 100
    # a never-written userset subject naming itself is allowed by name
    b = json.dumps({"items": [pod("ghost", ""), pod("x", "")]}).encode()
    sc = _lib.list_scan(b)[0]
    assert e.list_keep_allowed(b, sc, "pod", np.zeros(0, np.uint32), self_name="ghost").tolist() == [1, 0]
    # decode failures of the reference: non-object element, table row without "object"
    with pytest.raises(_lib.ZgpuError, match="not an object"):
        e.list_keep_allowed(b'{"items":[1]}', _lib.list_scan(b'{"items":[1]}')[0], "pod", allowed)
    rows = b'{"rows":[{"cells":[]}]}'
    with pytest.raises(_lib.ZgpuError, match="table row"):
        e.list_keep_allowed(rows, _lib.list_scan(rows, _lib.LIST_TABLE_ROWS)[0], "pod", allowed, _lib.LIST_TABLE_ROWS)
    with pytest.raises(_lib.ZgpuError, match="not found"):
        e.list_keep_allowed(b, sc, "nosuch", allowed)
    # the fused call needs the GPU for its LookupResources: loud failure here
    with pytest.raises(_lib.ZgpuError, match="no CPU fallback"):
        e.list_prefilter(b, e.list_template("pod", "view", "user", "u"))


def test_run_watch_rechecks_every_update_in_one_bulk_call():
    # watch.go:48-107 with the per-update CheckPermission folded into one bulk call per WatchResponse
    up = lambda op, rel: cl.RelationshipUpdate(op, cl.Relationship.parse(rel))
    stream = [cl.WatchResponse([up(cl.OPERATION_TOUCH, "pod:ns1/a#viewer@user:x"), up(cl.OPERATION_DELETE, "pod:ns1/b#viewer@user:y")], 7),
              cl.WatchResponse([], 8), cl.WatchResponse([up(cl.OPERATION_CREATE, "pod:cw#viewer@user:z")], 9)]
    mock = MockPermissionsClient({"pod:ns1/a#view@user:alice": HAS, "pod:cw#view@user:alice": HAS})
    got = pf.run_watch(stream, mock, ("pod", "$", "view", "user", "alice", ""))
    assert got == [pf.ResultChange(True, ("ns1", "a")), pf.ResultChange(False, ("ns1", "b")), pf.ResultChange(True, ("", "cw"))]
    assert mock.calls == [2, 1]
    with pytest.raises(RuntimeError):
        pf.run_watch(stream, MockPermissionsClient({}, errors={"pod:ns1/b#view@user:alice"}), ("pod", "$", "view", "user", "alice", ""))


def test_watch_frame_filter_event_loop():
    """responsefilterer.go:487-714 restated as a state machine: frames wait for the tracker."""
    frame = lambda typ, name, ns="ns", **kw: json.dumps({"type": typ, "object": {"kind": "Pod", **pod(name, ns), **kw}}).encode()
    f = pf.WatchFrameFilter()
    a1, a2, b1 = frame("ADDED", "a"), frame("MODIFIED", "a", spec=2), frame("ADDED", "b")
    assert f.on_frame(a1) == [] and f.on_frame(b1) == []           # nothing allowed yet: buffered
    assert f.on_frame(a2) == []                                      # a newer frame replaces the buffered one
    assert f.on_change(pf.ResultChange(True, ("ns", "a"))) == [a2]   # allowed: the buffered frame is flushed, once
    assert f.on_change(pf.ResultChange(True, ("ns", "a"))) == []
    a3 = frame("MODIFIED", "a", spec=3)
    assert f.on_frame(a3) == [a3]                                    # allowed objects pass at once
    assert f.on_frame(frame("DELETED", "a")) == [] and f.on_frame(frame("BOOKMARK", "a")) == []  # never written
    assert f.on_change(pf.ResultChange(False, ("ns", "b"))) == []    # denied: the buffered frame is forgotten
    assert f.on_change(pf.ResultChange(True, ("ns", "b"))) == []
    assert f.on_change(pf.ResultChange(False, ("ns", "a"))) == [] and f.on_frame(a3) == []  # revoked: buffered again
    # a Table frame is keyed by its first row's object
    table = json.dumps({"type": "ADDED", "object": {"kind": "Table", "apiVersion": "meta.k8s.io/v1",
                                                    "rows": [{"cells": ["t"], "object": pod("t", "ns")}]}}).encode()
    assert f.on_frame(table) == [] and f.on_change(pf.ResultChange(True, ("ns", "t"))) == [table]
    # cluster-scoped objects have an empty namespace
    node = json.dumps({"type": "ADDED", "object": {"metadata": {"name": "node1"}}}).encode()
    assert f.on_change(pf.ResultChange(True, ("", "node1"))) == [] and f.on_frame(node) == [node]
    # a Status passes through and ends the stream; so does a frame that does not decode
    status = b'{"kind":"Status","apiVersion":"v1","status":"Failure","code":410}'
    assert f.on_frame(status) == [status] and f.closed and f.on_frame(a3) == [] and f.on_change(pf.ResultChange(True, ("x", "y"))) == []
    g = pf.WatchFrameFilter()
    assert g.on_frame(b'{"type":"ADDED","object":') == [] and g.closed
