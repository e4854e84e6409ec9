"""Protobuf-encoded List responses (Content-Type application/vnd.kubernetes.protobuf): the pre-filter's list path
(pkg/authz/responsefilterer.go:241-313, :376-400) decodes with the negotiated serializer, and kube clients ask for
protobuf on built-in types. The checker is an independent encoder / decoder of the wire format written here (kube's
generated.proto files are not in /root/reference: k8s.io/apimachinery v0.34.1 is a go.mod dependency, not vendored):
filtering a body must give, byte for byte, the encoding of the list with only the kept items."""
import random

import numpy as np
import pytest

import zgpu  # noqa: F401
from spicedb_kubeapi_proxy_b200 import _lib, postfilter as pf

MAGIC = b"k8s\x00"


def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def ld(num, payload: bytes):  # length-delimited field
    return varint((num << 3) | 2) + varint(len(payload)) + payload


def vi(num, v):
    return varint(num << 3) + varint(v)


def object_meta(name=None, namespace=None, uid="u-1", labels=()):
    m = b""
    if name is not None:
        m += ld(1, name.encode())
    m += ld(2, b"")  # generateName
    if namespace is not None:
        m += ld(3, namespace.encode())
    m += ld(5, uid.encode()) + ld(6, b"12345") + vi(7, 3)
    for k, v in labels:
        m += ld(11, ld(1, k.encode()) + ld(2, v.encode()))
    return m


def pod(name=None, namespace=None, meta=True, spec_bytes=40):
    msg = ld(1, object_meta(name, namespace)) if meta else b""
    return msg + ld(2, b"s" * spec_bytes) + ld(3, ld(1, b"Running"))


def pod_list(pods, trailing=True):
    raw = ld(1, ld(2, b"987654") + ld(3, b""))  # ListMeta{resourceVersion, continue}
    for p in pods:
        raw += ld(2, p)
    unknown = ld(1, ld(1, b"v1") + ld(2, b"PodList")) + ld(2, raw)
    if trailing:
        unknown += ld(3, b"") + ld(4, b"")
    return MAGIC + unknown


def decode_items(body):
    """-> [(namespace, name)] by an independent walk of the wire format."""
    assert body[:4] == MAGIC

    def fields(b):
        i = 0
        while i < len(b):
            tag, i = rd(b, i)
            num, wt = tag >> 3, tag & 7
            if wt == 0:
                v, i = rd(b, i)
                yield num, v
            elif wt == 2:
                n, i = rd(b, i)
                yield num, b[i:i + n]
                i += n
            else:
                raise AssertionError(wt)

    def rd(b, i):
        v = s = 0
        while True:
            c = b[i]
            i += 1
            v |= (c & 0x7F) << s
            s += 7
            if not c & 0x80:
                return v, i

    raw = [v for n, v in fields(body[4:]) if n == 2][0]
    out = []
    for n, item in fields(raw):
        if n != 2:
            continue
        metas = [v for k, v in fields(item) if k == 1]
        name = ns = ""
        if metas:
            for k, v in fields(metas[-1]):
                if k == 1:
                    name = v.decode()
                if k == 3:
                    ns = v.decode()
        out.append((ns, name))
    return out


def scan(body):
    r = _lib.list_scan(body, _lib.LIST_PROTOBUF)
    assert r is not None
    return r


def names_of(body, items):
    return [(body[int(i["ns_off"]):int(i["ns_off"]) + int(i["ns_len"])].decode(),
             body[int(i["name_off"]):int(i["name_off"]) + int(i["name_len"])].decode()) for i in items]


def test_scan_finds_every_item_and_its_names():
    pods = [pod("a", "ns1"), pod("b-é中", "ns2"), pod("cluster-scoped"), pod(None, "only-ns"), pod(meta=False),
            pod("x" * 300, "y" * 200, spec_bytes=70000)]
    body = pod_list(pods)
    items, ib, ie = scan(body)
    assert names_of(body, items) == decode_items(body) == [("ns1", "a"), ("ns2", "b-é中"), ("", "cluster-scoped"),
                                                           ("only-ns", ""), ("", ""), ("y" * 200, "x" * 300)]
    assert all(int(f) & _lib.ITEM_IS_OBJECT and int(f) & _lib.ITEM_RAW_NAMES for f in items["flags"])
    assert [bool(int(f) & _lib.ITEM_HAS_METADATA) for f in items["flags"]] == [True, True, True, True, False, True]
    # an entry is the whole field: tag, length, message
    for it, p in zip(items, pods):
        assert body[int(it["begin"]):int(it["end"])] == ld(2, p)


@pytest.mark.parametrize("trailing", [True, False])
def test_filter_equals_encoding_the_kept_items(trailing):
    rng = random.Random(5)
    for n in (0, 1, 2, 7, 60):
        # sizes straddle the 1-/2-/3-byte length varints of an entry and of raw (127/128, 16383/16384)
        pods = [pod(f"p{i}", f"ns{i % 3}", spec_bytes=rng.choice([0, 1, 60, 90, 127, 128, 300, 16300, 16400])) for i in range(n)]
        body = pod_list(pods, trailing)
        items, ib, ie = scan(body) if n else (np.zeros(0, dtype=_lib.LIST_ITEM_DTYPE), *scan(body)[1:])
        for keep in ([1] * n, [0] * n, [rng.random() < 0.5 for _ in range(n)], [i % 7 == 0 for i in range(n)]):
            got = _lib.list_filter(body, items, np.array(keep, dtype=np.uint8), ib, ie)
            assert got == pod_list([p for p, k in zip(pods, keep) if k], trailing)
        assert _lib.list_filter(body, items, np.ones(n, dtype=np.uint8), ib, ie) == body


def test_prefilter_mirror_on_a_protobuf_list():
    """filter_list with the protobuf media type = the reference's filterList on the decoded list
    (responsefilterer.go:376-400): allowed (namespace, name) pairs stay."""
    pods = [pod("a", "n1"), pod("b", "n1"), pod("a", "n2"), pod("node-1"), pod(meta=False)]
    body = pod_list(pods)
    res = pf.PrefilterResult(allowed_results={("n1", "a"), ("n2", "a"), ("", "node-1")})
    out = pf.filter_list(body, res, "application/vnd.kubernetes.protobuf")
    assert decode_items(out) == [("n1", "a"), ("n2", "a"), ("", "node-1")]
    assert out == pod_list([pods[0], pods[2], pods[3]])
    assert pf.filter_list(body, pf.PrefilterResult(all_allowed=True), "application/vnd.kubernetes.protobuf;stream=watch") == body
    assert decode_items(pf.filter_list(body, pf.PrefilterResult(), "application/vnd.kubernetes.protobuf")) == []


def test_keep_allowed_resolves_protobuf_names_against_the_store():
    """zg_list_keep_allowed on protobuf items: names are plain bytes (no JSON unescaping), ids come from a
    LookupResources answer. Host-only engine: no GPU work on this path."""
    e = zgpu.Engine("definition user {}\ndefinition pod { relation viewer: user  permission view = viewer }", host_only=True)
    ids = {n: e.intern("pod", n) for n in ("n1/a", "n1/b", "n2/a", "node-1", r"n3/back\\slash")}
    pods = [pod("a", "n1"), pod("b", "n1"), pod("a", "n2"), pod("node-1"), pod("node-1", "other"), pod(r"back\\slash", "n3")]
    body = pod_list(pods)
    items, ib, ie = scan(body)
    allowed = np.sort(np.array([ids["n1/a"], ids["node-1"], ids[r"n3/back\\slash"]], dtype=np.uint32))
    keep = e.list_keep_allowed(body, items, "pod", allowed, _lib.LIST_PROTOBUF)
    assert keep.tolist() == [1, 0, 0, 1, 0, 1]
    keep = e.list_keep_allowed(body, items, "pod", allowed, _lib.LIST_PROTOBUF, req_namespace="other")
    assert keep.tolist() == [1, 0, 0, 0, 1, 1]  # cluster-scoped ids take the request's namespace (lookups.go:117-127)


def test_malformed_bodies_are_refused():
    good = pod_list([pod("a", "n"), pod("b", "n")])
    raw_at = good.index(b"\x12", 4 + 2)  # not relied on below beyond being inside the envelope
    assert raw_at > 0
    bad = [b"", b"k8s", b"k9s\x00" + good[4:], good[:-3], good[:len(good) // 2],
           MAGIC + ld(1, b"") + varint((2 << 3) | 0) + varint(5),          # raw with the wrong wire type
           MAGIC + ld(2, ld(2, b"\x0a\x05abc")),                           # item's metadata runs past the item
           MAGIC + ld(2, varint((2 << 3) | 0) + varint(1)),                # items entry that is a varint
           MAGIC + ld(2, ld(2, varint((1 << 3) | 3))),                     # a group inside an item
           MAGIC + ld(2, b"") + ld(2, b""),                                # raw twice
           MAGIC + varint((2 << 3) | 2) + b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01",  # varint too long
           MAGIC + varint(2)]                                              # field number 0
    for b in bad:
        with pytest.raises(_lib.ZgpuError):
            _lib.list_scan(b, _lib.LIST_PROTOBUF)
    # JSON scanners refuse a protobuf body and vice versa (the post-filter json.Unmarshals: postfilter.go:19)
    with pytest.raises(_lib.ZgpuError):
        _lib.list_scan(good, _lib.LIST_ITEMS)
    with pytest.raises(_lib.ZgpuError):
        _lib.list_scan(b'{"items":[]}', _lib.LIST_PROTOBUF)
    assert _lib.list_scan(MAGIC + ld(1, ld(2, b"Status")), _lib.LIST_PROTOBUF) is None  # no raw: passes through


def test_filter_refuses_item_ranges_that_are_not_the_scanned_ones():
    body = pod_list([pod("a", "n"), pod("b", "n")])
    items, ib, ie = scan(body)
    sw = items[::-1].copy()
    with pytest.raises(_lib.ZgpuError):
        _lib.list_filter(body, sw, np.array([1, 0], dtype=np.uint8), ib, ie)
    with pytest.raises(_lib.ZgpuError):
        _lib.list_filter(body, items, np.array([1, 0], dtype=np.uint8), ib + 1, ie)
    with pytest.raises(_lib.ZgpuError):
        _lib.list_filter(body, items, np.array([1, 0], dtype=np.uint8), ib, ie - 1)


def test_random_lists_round_trip():
    rng = random.Random(77)
    for _ in range(300):
        n = rng.randrange(0, 12)
        pods = [pod(None if rng.random() < 0.1 else "p" * rng.randrange(1, 40) + str(i),
                    None if rng.random() < 0.3 else "ns" + str(rng.randrange(4)),
                    meta=rng.random() > 0.05, spec_bytes=rng.randrange(0, 400)) for i in range(n)]
        body = pod_list(pods, rng.random() < 0.5)
        r = _lib.list_scan(body, _lib.LIST_PROTOBUF)
        items, ib, ie = r
        assert names_of(body, items) == decode_items(body)
        keep = [rng.random() < 0.6 for _ in range(n)]
        out = _lib.list_filter(body, items, np.array(keep, dtype=np.uint8), ib, ie)
        assert out == pod_list([p for p, k in zip(pods, keep) if k], body.endswith(ld(4, b"")))


def pod_object(p):
    return MAGIC + ld(1, ld(1, b"v1") + ld(2, b"Pod")) + ld(2, p) + ld(3, b"") + ld(4, b"")


def test_single_protobuf_object_passes_or_is_unauthorized():
    """The `default:` branch (a get) on a protobuf body: responsefilterer.go:320-341, filterObject :403-415."""
    body = pod_object(pod("a", "n1"))
    r = _lib.list_scan(body, _lib.LIST_PROTOBUF_OBJECT)
    items, ib, ie = r
    assert len(items) == 1 and names_of(body, items) == [("n1", "a")]
    assert body[int(items[0]["begin"]):int(items[0]["end"])] == pod("a", "n1")
    ct = "application/vnd.kubernetes.protobuf"
    assert pf.filter_object(body, pf.PrefilterResult(allowed_results={("n1", "a")}), ct) == body
    with pytest.raises(pf.Unauthorized):
        pf.filter_object(body, pf.PrefilterResult(allowed_results={("n2", "a")}), ct)
    with pytest.raises(pf.Unauthorized):
        pf.filter_object(pod_object(pod(meta=False)), pf.PrefilterResult(allowed_results={("n1", "a")}), ct)
    assert pf.filter_object(pod_object(pod("node-1")), pf.PrefilterResult(allowed_results={("", "node-1")}), ct)
    with pytest.raises(ValueError):
        pf.filter_object(body[:-9], pf.PrefilterResult(all_allowed=True), ct)
