"""CPU-only tests of the product's HOST logic (no GPU, no compute calls):
the C ABI library loads and exports every symbol include/zgpu.h declares, the schema
compiler agrees with the oracle's slot numbering, the CSR snapshot builder lays rows
out as documented, and the relationship store implements the v1 write / read /
delete / precondition contract the reference relies on
(pkg/authz/distributedtx/activity.go:54-76,128-171; pkg/authz/update.go:207-271)."""
import ctypes
import os
import re

import numpy as np
import pytest

import randgen
import zgpu
from oracle.pyoracle import Oracle
from spicedb_kubeapi_proxy_b200 import _lib, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zgpu.h")).read()
    declared = set(re.findall(r"\b(zg_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed from include/zgpu.h"
    L = ctypes.CDLL(zgpu.build_library())
    for name in sorted(declared):
        assert hasattr(L, name), f"libzgpu.so does not export {name}"
    # the python binding binds exactly the declared set
    assert set(_lib.exported_symbols()) == declared


def test_no_cpu_fallback_on_hot_path():
    e = zgpu.Engine(workloads.CFG2_SCHEMA, host_only=True)
    e.add_bulk("pod", "viewer", "user", [1, 2], [3, 4])
    e.publish()
    items = np.zeros(2, dtype=zgpu.CHECK_DTYPE)
    with pytest.raises(zgpu.ZgpuError, match="no CPU fallback"):
        e.check_bulk(items)
    with pytest.raises(zgpu.ZgpuError, match="no CPU fallback"):
        e.lookup_resources_ids("pod", "view", "user", 3)
    with pytest.raises(zgpu.ZgpuError, match="no CPU fallback"):
        e.check_bulk_str(["pod:a#view@user:b"])


def test_engine_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zgpu.ZgpuError, match="no CUDA device"):
        zgpu.Engine(workloads.CFG2_SCHEMA)


@pytest.mark.parametrize("schema", [workloads.BOOTSTRAP_SCHEMA, workloads.CFG3_SCHEMA, workloads.CFG4_SCHEMA]
                         + list(randgen.FIXED_SCHEMAS.values()))
def test_slot_numbering_matches_oracle(schema):
    e, o = zgpu.Engine(schema, host_only=True), Oracle(schema)
    assert e.slot_table() == o.slot_table()
    for t in ("user", "group", "namespace", "document", "nope"):
        assert e.type_id(t) == o.type_id(t)


@pytest.mark.parametrize("seed", range(20))
def test_random_schemas_compile_like_the_oracle(seed):
    import random

    schema, _ = randgen.random_schema(random.Random(seed))
    e, o = zgpu.Engine(schema, host_only=True), Oracle(schema)
    assert e.slot_table() == o.slot_table()


BAD_SCHEMAS = {
    "definition a { relation r: nope }": "unknown subject type",
    "definition a { relation r: a#zz }": "does not exist",
    "definition a { permission p = q }": "unknown relation or permission",
    "definition a { relation r: a  permission p = p2->x  permission p2 = r }": "needs a relation on the left",
    "definition a { relation r: a with somecaveat }": "caveats are not supported",
    "caveat c(x int) { x > 1 }": "caveats are not supported",
    "definition a { permission p = p }": "refers to itself",
    "definition a { relation r: a  relation r: a }": "duplicate",
    "definition a {": "expected",
    "definition a { relation r: a  permission p = r.all(p) }": "only .any",
}


@pytest.mark.parametrize("text,msg", list(BAD_SCHEMAS.items()))
def test_schema_errors(text, msg):
    with pytest.raises(zgpu.ZgpuError, match=msg):
        zgpu.Engine(text, host_only=True)
    with pytest.raises(Exception):
        Oracle(text)


def test_operator_precedence_matches_oracle_and_mini():
    """'+' binds tightest, then '&', then '-' (SpiceDB DSL); same in all three parsers."""
    from oracle.mini_oracle import parse_schema

    s = "definition user {} definition d { relation a: user relation b: user relation c: user " \
        "permission p = a - b + c  permission q = a & b + c  permission r = a + b - c & a }"
    d = parse_schema(s)["d"]["permissions"]
    assert d["p"] == ("-", ("ref", "a"), ("+", ("ref", "b"), ("ref", "c")))
    assert d["q"] == ("&", ("ref", "a"), ("+", ("ref", "b"), ("ref", "c")))
    assert d["r"] == ("-", ("+", ("ref", "a"), ("ref", "b")), ("&", ("ref", "c"), ("ref", "a")))
    zgpu.Engine(s, host_only=True)


def test_csr_rows_match_numpy_reference():
    w = workloads.cfg3(scale=0.002)
    e = zgpu.Engine(w.schema, host_only=True)
    w.load_into(e)
    e.publish()
    g = w.groups[0]  # group#member@user -> class 0 of (user | group#member)
    for grp in np.unique(g.res)[:50]:
        want = np.unique(g.subj[g.res == grp])
        assert np.array_equal(e.debug_row("group", "member", int(grp), 0), want)
        assert e.debug_row("group", "member", int(grp), 1).size == 0
    t = w.groups[1]  # team#member@group#member -> class 0 of (group#member | user)
    for team in np.unique(t.res)[:20]:
        assert np.array_equal(e.debug_row("team", "member", int(team), 0), np.unique(t.subj[t.res == team]))
    # reverse CSR (subject -> resources), what the direction-optimised probes and LookupResources read
    for u in np.unique(g.subj)[:50]:
        assert np.array_equal(e.debug_row("group", "member", int(u), 0, reverse=True), np.unique(g.res[g.subj == u]))
    for grp in np.unique(t.subj)[:20]:
        assert np.array_equal(e.debug_row("team", "member", int(grp), 0, reverse=True), np.unique(t.res[t.subj == grp]))
    # duplicates in bulk loads fold (TOUCH semantics)
    e2 = zgpu.Engine(workloads.CFG2_SCHEMA, host_only=True)
    e2.add_bulk("pod", "viewer", "user", [5, 5, 5, 1], [9, 9, 2, 7])
    e2.publish()
    assert list(e2.debug_row("pod", "viewer", 5)) == [2, 9]
    assert e2.num_tuples() == 3


def test_write_read_delete_preconditions():
    C = zgpu.client
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    cl = C.PermissionsClient(workloads.BOOTSTRAP_SCHEMA, engine=e)
    up = lambda op, rel, exp=0: C.RelationshipUpdate(op, C.Relationship.parse(rel, exp))
    cl.WriteRelationships(C.WriteRelationshipsRequest([
        up(C.OPERATION_CREATE, "namespace:ns1#creator@user:paul"),
        up(C.OPERATION_CREATE, "namespace:ns1#cluster@cluster:cluster"),
        up(C.OPERATION_TOUCH, "pod:ns1/p1#viewer@user:app"),
    ]))
    read = lambda **f: sorted(r.relationship.text() for r in cl.ReadRelationships(
        C.ReadRelationshipsRequest(C.RelationshipFilter(**f))))
    assert read(resource_type="namespace") == ["namespace:ns1#cluster@cluster:cluster", "namespace:ns1#creator@user:paul"]
    assert read(resource_type="pod", optional_resource_id="ns1/p1", optional_relation="viewer") == ["pod:ns1/p1#viewer@user:app"]
    assert read(resource_type="namespace", optional_subject_filter=C.SubjectFilter("user", "paul")) == \
        ["namespace:ns1#creator@user:paul"]
    assert read(resource_type="pod", optional_resource_id="never-written") == []
    # CREATE of an existing relationship fails the whole write (nothing applied)
    with pytest.raises(C.RpcError, match="ALREADY_EXISTS"):
        cl.WriteRelationships(C.WriteRelationshipsRequest([
            up(C.OPERATION_TOUCH, "namespace:ns2#creator@user:chani"),
            up(C.OPERATION_CREATE, "namespace:ns1#creator@user:paul")]))
    assert read(resource_type="namespace", optional_resource_id="ns2") == []
    # preconditions (pkg/authz/distributedtx/workflow.go:147-182 builds MUST_NOT_MATCH ones)
    pre_absent = C.Precondition(C.PRECONDITION_MUST_NOT_MATCH, C.RelationshipFilter(
        "namespace", "ns1", "cluster", C.SubjectFilter("cluster", "cluster")))
    with pytest.raises(C.RpcError, match="FAILED_PRECONDITION"):
        cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "namespace:ns1#viewer@user:x")], [pre_absent]))
    assert read(resource_type="namespace", optional_relation="viewer") == []
    pre_present = C.Precondition(C.PRECONDITION_MUST_MATCH, pre_absent.filter)
    cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "namespace:ns1#viewer@user:x")], [pre_present]))
    assert read(resource_type="namespace", optional_relation="viewer") == ["namespace:ns1#viewer@user:x"]
    # schema validation of writes
    with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):
        cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "namespace:ns1#viewer@cluster:c")]))
    with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):
        cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "namespace:ns1#view@user:u")]))
    with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):  # expiration only where the schema allows it
        cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "namespace:ns1#viewer@user:u", 99)]))
    cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "workflow:w#idempotency_key@activity:a", 2_000_000_000)]))
    e.set_clock(1_999_999_999)
    assert read(resource_type="workflow") == ["workflow:w#idempotency_key@activity:a"]
    e.set_clock(2_000_000_000)
    assert read(resource_type="workflow") == []  # expired relationships are invisible
    # DELETE update + delete by filter
    cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_DELETE, "namespace:ns1#viewer@user:x")]))
    assert read(resource_type="namespace", optional_relation="viewer") == []
    n = cl.DeleteRelationships(C.DeleteRelationshipsRequest(C.RelationshipFilter("namespace", "ns1")))
    assert n == 2 and read(resource_type="namespace") == []
    # object ids: 1..1024 chars of [a-zA-Z0-9/_|\-=+]; '/' is legal and used (ns/name, pkg/rules/rules.go:335-339)
    cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, "pod:" + "a" * 1024 + "#viewer@user:x_y|z-1=2+3")]))
    for bad in ("pod:" + "a" * 1025 + "#viewer@user:u", "pod:has space#viewer@user:u", "pod:p#viewer@user:u!", "pod:*#viewer@user:u"):
        with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):
            cl.WriteRelationships(C.WriteRelationshipsRequest([up(C.OPERATION_TOUCH, bad)]))
    # more than 1000 updates per write is rejected (pkg/spicedb/spicedb.go:34)
    with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):
        cl.WriteRelationships(C.WriteRelationshipsRequest(
            [up(C.OPERATION_TOUCH, f"pod:p{i}#viewer@user:u") for i in range(1001)]))


def test_interleaved_rows_and_wildcard_reverse_row():
    """All relations of a type share one row table (object-major); wildcard classes keep one
    reverse row listing every resource that carries the wildcard."""
    w = workloads.cfg4(scale=0.0005)
    e = zgpu.Engine(w.schema, host_only=True)
    w.load_into(e)
    e.publish()
    by = {(g.res_type, g.rel, g.subj_type, g.srel, g.wildcard): g for g in w.groups}
    dv = by[("document", "viewer", "user", None, False)]
    dg = by[("document", "viewer", "group", "member", False)]
    dw = by[("document", "viewer", "user", None, True)]
    db = by[("document", "banned", "user", None, False)]
    for d in np.unique(dv.res)[:40]:
        assert np.array_equal(e.debug_row("document", "viewer", int(d), 0), np.unique(dv.subj[dv.res == d]))
        assert np.array_equal(e.debug_row("document", "viewer", int(d), 1), np.unique(dg.subj[dg.res == d]))
        assert e.debug_row("document", "viewer", int(d), 2).size == int((dw.res == d).any())
        assert np.array_equal(e.debug_row("document", "banned", int(d), 0), np.unique(db.subj[db.res == d]))
    assert np.array_equal(e.debug_row("document", "viewer", 0, 2, reverse=True), np.unique(dw.res))


def test_workload_generators_are_deterministic_and_sized():
    a, b = workloads.cfg3(scale=0.001), workloads.cfg3(scale=0.001)
    assert all(np.array_equal(x.res, y.res) and np.array_equal(x.subj, y.subj) for x, y in zip(a.groups, b.groups))
    assert np.array_equal(a.checks[0].res, b.checks[0].res)
    w1 = workloads.cfg1()
    assert 400 < w1.n_tuples() < 600 and w1.n_checks() == 1000
    w4 = workloads.cfg4(scale=0.0005)
    rels = {(g.res_type, g.rel, g.subj_type, g.srel, g.wildcard) for g in w4.groups}
    assert ("document", "viewer", "user", None, True) in rels and ("folder", "parent", "folder", None, False) in rels
    o = Oracle(w4.schema)
    w4.load_into(o)  # every generated relationship is valid under the schema


def test_watch_feed():
    """v1.WatchServiceClient.Watch as the proxy uses it (pkg/authz/watch.go:27-48): updates for one resource
    type, after the point the watch started, every applied change, grouped per write."""
    C = zgpu.client
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    cl = C.PermissionsClient(workloads.BOOTSTRAP_SCHEMA, engine=e)
    up = lambda op, rel, exp=0: C.RelationshipUpdate(op, C.Relationship.parse(rel, exp))
    write = lambda *u, pre=(): cl.WriteRelationships(C.WriteRelationshipsRequest(list(u), list(pre))).written_at
    r0 = write(up(C.OPERATION_TOUCH, "namespace:old#creator@user:paul"))
    pods = cl.Watch(C.WatchRequest(["pod"]))            # starts "now": the write above is not replayed
    everything = cl.Watch(C.WatchRequest([], optional_start_cursor=0))
    assert pods.Recv() is None
    r1 = write(up(C.OPERATION_CREATE, "pod:ns1/p1#viewer@user:app"), up(C.OPERATION_TOUCH, "namespace:ns1#creator@user:paul"),
               up(C.OPERATION_TOUCH, "pod:ns1/p2#viewer@user:app"))
    r2 = write(up(C.OPERATION_TOUCH, "pod:ns1/p1#viewer@user:app"),          # TOUCH of an existing one is still a change
               up(C.OPERATION_DELETE, "pod:ns1/never#viewer@user:app"),     # deleting nothing is not
               up(C.OPERATION_DELETE, "pod:ns1/p2#viewer@user:app"))
    assert r0 < r1 < r2
    got = [(r.changes_through, [(u.operation, u.relationship.text()) for u in r.updates]) for r in pods]
    assert got == [(r1, [(C.OPERATION_CREATE, "pod:ns1/p1#viewer@user:app"), (C.OPERATION_TOUCH, "pod:ns1/p2#viewer@user:app")]),
                   (r2, [(C.OPERATION_TOUCH, "pod:ns1/p1#viewer@user:app"), (C.OPERATION_DELETE, "pod:ns1/p2#viewer@user:app")])]
    assert pods.Recv() is None and pods.cursor == r2
    # a failed write (precondition, CREATE of an existing relationship) leaves no trace in the feed
    pre = C.Precondition(C.PRECONDITION_MUST_MATCH, C.RelationshipFilter("pod", "nope"))
    with pytest.raises(C.RpcError):
        write(up(C.OPERATION_TOUCH, "pod:ns1/p3#viewer@user:app"), pre=[pre])
    with pytest.raises(C.RpcError):
        write(up(C.OPERATION_TOUCH, "pod:ns1/p4#viewer@user:app"), up(C.OPERATION_CREATE, "pod:ns1/p1#viewer@user:app"))
    assert pods.Recv() is None
    # delete by filter reports every relationship it removed; expirations ride along
    write(up(C.OPERATION_TOUCH, "workflow:w#idempotency_key@activity:a", 2_000_000_000))
    assert cl.DeleteRelationships(C.DeleteRelationshipsRequest(C.RelationshipFilter("pod"))) == 1
    r = pods.Recv()
    assert [(u.operation, u.relationship.text()) for u in r.updates] == [(C.OPERATION_DELETE, "pod:ns1/p1#viewer@user:app")]
    allu = [(u.operation, u.relationship.text(), u.relationship.optional_expires_at) for r in everything for u in r.updates]
    assert allu[0] == (C.OPERATION_TOUCH, "namespace:old#creator@user:paul", 0) and len(allu) == 8
    assert (C.OPERATION_TOUCH, "workflow:w#idempotency_key@activity:a", 2_000_000_000) in allu
    # several types: filtered client-side; unknown type: INVALID_ARGUMENT
    two = cl.Watch(C.WatchRequest(["pod", "workflow"], optional_start_cursor=0))
    assert sum(len(r.updates) for r in two) == 6
    with pytest.raises(C.RpcError, match="INVALID_ARGUMENT"):
        cl.Watch(C.WatchRequest(["nosuchtype"]))
    # E2BIG protocol of the C entry point
    import ctypes as ct
    L = zgpu._lib.lib()
    need, n, through = ct.c_size_t(0), ct.c_uint64(0), ct.c_uint64(0)
    assert L.zg_watch_read(e._h, 0, b"", None, 0, ct.byref(need), ct.byref(n), ct.byref(through)) == -7
    assert n.value == 8 and through.value == e.stats()["revision"]
    buf = ct.create_string_buffer(need.value)
    assert L.zg_watch_read(e._h, 0, None, buf, need.value, ct.byref(need), ct.byref(n), ct.byref(through)) == 0
    assert buf.value.decode().splitlines()[0] == f"{r0} TOUCH namespace:old#creator@user:paul"
    assert buf.value.decode().splitlines()[-2].endswith("TOUCH workflow:w#idempotency_key@activity:a 2000000000")


def test_watch_drives_recheck_like_run_watch():
    """RunWatch's loop (watch.go:36-107) over the mirror: every update of the watched type names a resource to
    re-check; the operation is not inspected."""
    C = zgpu.client
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    cl = C.PermissionsClient(workloads.BOOTSTRAP_SCHEMA, engine=e)
    stream = cl.Watch(C.WatchRequest(["pod"]))
    for i in range(5):
        cl.WriteRelationships(C.WriteRelationshipsRequest([C.RelationshipUpdate(
            C.OPERATION_TOUCH, C.Relationship.parse(f"pod:ns/p{i}#viewer@user:u{i}"))]))
    rechecks = [(u.relationship.resource.object_id, u.relationship.subject.object.object_id)
                for r in stream for u in r.updates]
    assert rechecks == [(f"ns/p{i}", f"u{i}") for i in range(5)]


def test_resolve_checks_matches_item_by_item_resolution():
    """zg_resolve_checks (the ingress half of zg_check_bulk_str): the memo over repeated literal fields must not
    change any answer -- a shuffled batch resolves exactly like the same items one call at a time."""
    import random
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    cl = zgpu.client.PermissionsClient(workloads.BOOTSTRAP_SCHEMA, engine=e)
    C = zgpu.client
    cl.WriteRelationships(C.WriteRelationshipsRequest(
        [C.RelationshipUpdate(C.OPERATION_TOUCH, C.Relationship.parse(f"pod:ns{i % 5}/p{i}#viewer@user:u{i % 7}")) for i in range(60)] +
        [C.RelationshipUpdate(C.OPERATION_TOUCH, C.Relationship.parse(f"namespace:ns{i}#creator@user:u{i}")) for i in range(5)]))
    rng = random.Random(3)
    types = ["pod", "namespace", "user", "nosuch", "cluster"]
    rels = ["view", "viewer", "creator", "edit", "nosuch", "cluster"]
    items = []
    for _ in range(3000):
        rt, st = rng.choice(types[:3] + types), rng.choice(["user"] * 4 + types)
        rid = rng.choice([f"ns{rng.randrange(6)}/p{rng.randrange(70)}", f"ns{rng.randrange(6)}", f"u{rng.randrange(9)}", "never"])
        sid = rng.choice([f"u{rng.randrange(9)}", "never", rid])
        items.append((rt, rid, rng.choice(rels), st, sid, rng.choice(["", "", "", "...", "viewer", "nosuch"])))
    # runs of repeated templates, as a post-filter produces them
    items += [("pod", f"ns1/p{i}", "view", "user", "u3", "") for i in range(200)]
    batch = e.resolve_checks(items)
    single = np.concatenate([e.resolve_checks([it]) for it in items])
    assert batch.tobytes() == single.tobytes()
    known = [b for b in batch if b["perm"] != 0xFFFF]
    assert len(known) > 300 and any(b["res"] != 0xFFFFFFFF for b in known) and any(b["res"] == 0xFFFFFFFF for b in known)
    # two never-written names that are the same object share the sentinel; different ones do not
    same, diff = e.resolve_checks([("pod", "zz", "view", "pod", "zz", "viewer"), ("pod", "zz", "view", "pod", "yy", "viewer")])
    assert (same["res"], same["subj"]) == (0xFFFFFFFF, 0xFFFFFFFF) and (diff["res"], diff["subj"]) == (0xFFFFFFFF, 0xFFFFFFFE)


def test_interning_index_survives_growth_and_numeric_gaps():
    e = zgpu.Engine(workloads.CFG2_SCHEMA, host_only=True)
    t = next(iter(n for n in ("user", "document", "doc") if e.type_id(n) >= 0))
    ids = [e.intern(t, f"name-{i}-" + "x" * (i % 40)) for i in range(5000)]   # several table growths, short and long keys
    assert ids == list(range(5000))
    assert [e.find(t, f"name-{i}-" + "x" * (i % 40)) for i in range(0, 5000, 7)] == list(range(0, 5000, 7))
    NO = 0xFFFFFFFF
    assert e.find(t, "name-5000-") == NO and e.find(t, "") == NO and e.find(t, "name-1") == NO
    assert e.intern(t, "name-17-" + "x" * 17) == 17


def test_host_code_under_sanitizers(tmp_path):
    """tests/fuzz/host_fuzz.cc: the schema compiler on mutated texts, and the relationship store + CSR builder on
    random TOUCH/CREATE/DELETE streams checked row by row against a model -- the same .cc files libzgpu links,
    built here with ASan + UBSan."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "spicedb-kubeapi-proxy_b200", "csrc")
    exe = str(tmp_path / "host_fuzz")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            "-o", exe, os.path.join(root, "tests", "fuzz", "host_fuzz.cc"),
                            os.path.join(csrc, "schema.cc"), os.path.join(csrc, "store.cc")], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), (run.stdout[-800:], run.stderr[-2000:])


def test_batcher_under_concurrent_callers_without_a_gpu():
    """The group-commit batcher (csrc/capi.cu) runs before any CUDA call, so its queueing, leader hand-over and
    wake-ups can be stressed here: on a host-only engine every caller must come back with the loud ZG_ECUDA
    error -- its own, exactly once, no deadlock -- while a writer keeps taking the engine lock."""
    import threading
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    e.write_relationships([(zgpu._lib.OP_TOUCH, "pod:ns/p#viewer@user:u", 0)])
    items = np.zeros(64, dtype=zgpu.CHECK_DTYPE)
    results, stop = [], threading.Event()

    def caller(tid):
        got = 0
        for k in range(150):
            try:
                e.check_bulk(items[: 1 + (tid + k) % 64])
            except zgpu.ZgpuError as ex:
                got += "no CPU fallback" in str(ex)
        results.append(got)

    def writer():
        i = 0
        while not stop.is_set():
            e.write_relationships([(zgpu._lib.OP_TOUCH if i % 2 == 0 else zgpu._lib.OP_DELETE, "pod:ns/q#viewer@user:u", 0)])
            i += 1

    wt = threading.Thread(target=writer)
    wt.start()
    th = [threading.Thread(target=caller, args=(t,)) for t in range(24)]
    [t.start() for t in th]
    for t in th:
        t.join(timeout=120)
    stop.set()
    wt.join(timeout=30)
    assert not any(t.is_alive() for t in th) and not wt.is_alive(), "batcher deadlock"
    assert results == [150] * 24


def test_packed_resolution_equals_the_string_path():
    """zg_resolve_checks_packed (two buffers for n items) against zg_resolve_checks (6 n C strings)."""
    import random
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    for i in range(200):
        e.intern("pod", f"ns{i % 5}/p{i}")
        e.intern("user", f"u{i % 20}")
    e.intern("pod", "")  # the empty name is a legal key of the index
    rng = random.Random(9)
    res = [rng.choice([f"ns{rng.randrange(6)}/p{rng.randrange(230)}", "never", "", "u3"]) for _ in range(40000)]
    subs = [rng.choice([f"u{rng.randrange(25)}", "never", ""]) for _ in res]
    for rt, rel, st, srel in (("pod", "view", "user", ""), ("pod", "viewer", "user", "..."), ("pod", "view", "pod", "viewer"),
                              ("pod", "nosuch", "user", ""), ("nosuch", "view", "user", ""), ("pod", "view", "user", "nosuch")):
        want = e.resolve_checks([(rt, r, rel, st, s, srel) for r, s in zip(res, subs)])
        assert e.resolve_checks_packed(rt, rel, st, res, subs, srel).tobytes() == want.tobytes()
        one = e.resolve_checks([(rt, r, rel, st, "u7", srel) for r in res])
        assert e.resolve_checks_packed(rt, rel, st, res, "u7", srel).tobytes() == one.tobytes()
    # same never-written object on both sides: shared sentinel, per item and with one subject
    got = e.resolve_checks_packed("pod", "view", "pod", ["ghost", "ghost"], ["ghost", "other"], "viewer")
    assert [(int(g["res"]), int(g["subj"])) for g in got] == [(0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFE)]
    got = e.resolve_checks_packed("pod", "view", "pod", ["ghost", "x"], "ghost", "viewer")
    assert [(int(g["res"]), int(g["subj"])) for g in got] == [(0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFE)]
    assert e.resolve_checks_packed("pod", "view", "user", [], "u1").size == 0
    # decreasing offsets are refused; the fused call fails loudly without a GPU
    import ctypes as ct
    L = zgpu._lib.lib()
    off = np.array([0, 3, 1], dtype=np.uint32)
    out = np.zeros(2, dtype=zgpu.CHECK_DTYPE)
    assert L.zg_resolve_checks_packed(e._h, b"pod", b"view", b"user", b"", b"abc", off.ctypes.data, b"u", None, 2, out.ctypes.data) == -1
    with pytest.raises(zgpu.ZgpuError, match="no CPU fallback"):
        e.check_bulk_packed("pod", "view", "user", ["ns1/p1"], "u1")


def test_object_names_round_trip_across_numeric_gaps():
    import ctypes as ct
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    L, t = zgpu._lib.lib(), e.type_id("pod")
    a = e.intern("pod", "ns/a")
    e.add_bulk("pod", "viewer", "user", [5], [0])  # numeric ids 0..5 now exist for pods
    b = e.intern("pod", "after-gap")
    assert (a, b) == (0, 6)  # string ids and bulk numeric ids share one id space
    buf = ct.create_string_buffer(64)
    assert L.zg_object_name(e._h, t, a, buf, 64) == 4 and buf.value == b"ns/a"
    assert L.zg_object_name(e._h, t, b, buf, 64) == 9 and buf.value == b"after-gap"
    assert L.zg_object_name(e._h, t, 3, buf, 64) == -1 and L.zg_object_name(e._h, t, 99, buf, 64) == -1  # numeric-only / unknown
    assert L.zg_object_name(e._h, t, a, buf, 3) == -7 and L.zg_object_name(e._h, t, a, None, 0) == -7
    assert e.read_relationships(res_type="pod") == ["pod:5#viewer@user:0"]
