"""Cross-check the C oracle against the independent pure-Python oracle on random
schemas and graphs. This covers the operators the reference's own tests never pin
(intersection, exclusion, arrows, userset subjects, wildcards, depth cap,
expiration): "parity unpinned" against SpiceDB itself, but two independent
restatements of the documented semantics must agree."""
import random

import pytest

import randgen
from golden_runner import split_rel
from oracle.mini_oracle import MiniOracle
from oracle.pyoracle import Oracle


def _compare(schema, rels, checks, lookups=()):
    c, m = Oracle(schema), MiniOracle(schema)
    for r in rels:
        c.touch(r)
        m.write(r)
    for q in checks:
        a, b = c.check(*split_rel(q)), m.check(*split_rel(q))
        assert a == b, f"{q}: C={a} mini={b}\n{schema}\n" + "\n".join(rels)
    for (rt, perm, st, sid, srel) in lookups:
        a = sorted(c.lookup_resources(rt, perm, st, sid, srel))
        b = sorted(m.lookup_resources(rt, perm, st, sid, srel))
        assert a == b, f"lookup {rt}#{perm}@{st}:{sid}#{srel}: C={a} mini={b}"


@pytest.mark.parametrize("name", sorted(randgen.FIXED_SCHEMAS))
@pytest.mark.parametrize("seed", range(4))
def test_fixed_schemas(name, seed):
    rng = random.Random(1000 + seed)
    schema = randgen.FIXED_SCHEMAS[name]
    model = randgen.model_from_schema(schema)
    rels = randgen.random_relationships(rng, model, n_obj=5, n_user=5, density=0.3)
    checks = randgen.random_checks(rng, model, 300, n_obj=5, n_user=5)
    lookups = []
    for t, d in model["types"].items():
        for p in list(d["perms"])[:2]:
            lookups.append((t, p, "user", f"u{rng.randint(0, 5)}", ""))
    _compare(schema, rels, checks, lookups)


@pytest.mark.parametrize("seed", range(40))
def test_random_schemas(seed):
    rng = random.Random(seed)
    schema, model = randgen.random_schema(rng)
    rels = randgen.random_relationships(rng, model, n_obj=5, n_user=4, density=0.35)
    checks = randgen.random_checks(rng, model, 200, n_obj=5, n_user=4)
    lookups = []
    for t in model["tnames"]:
        for p in list(model["types"][t]["perms"])[:2]:
            lookups.append((t, p, "user", f"u{rng.randint(0, 4)}", ""))
    _compare(schema, rels, checks, lookups)


CHAIN = """
definition user {}
definition group { relation member: user | group#member }
definition folder {
  relation parent: folder
  relation viewer: user
  permission view = viewer + parent->view
  permission not_view = viewer - parent->view
}
"""


def _chain(o, n):
    for i in range(n):
        o(f"group:g{i}#member@group:g{i+1}#member")
    o(f"group:g{n}#member@user:deep")


@pytest.mark.parametrize("impl", ["c", "mini"])
def test_depth_cap(impl):
    """pkg/spicedb/spicedb.go:33 -- dispatch depth 50: the 51st hop is an error."""
    def mk():
        if impl == "c":
            o = Oracle(CHAIN)
            return o, o.touch
        o = MiniOracle(CHAIN)
        return o, o.write
    o, w = mk()
    _chain(w, 50)  # g0 -> ... -> g50 : 50 hops, allowed
    assert o.check("group", "g0", "member", "user", "deep") == 2
    assert o.check("group", "g0", "member", "user", "nobody") == 1
    o, w = mk()
    _chain(w, 51)  # needs a 51st hop
    assert o.check("group", "g0", "member", "user", "deep") == 255
    assert o.check("group", "g1", "member", "user", "deep") == 2
    # a cycle never terminates by itself: depth cap => error, unless found first
    o, w = mk()
    w("group:a#member@group:b#member")
    w("group:b#member@group:a#member")
    w("group:b#member@user:x")
    assert o.check("group", "a", "member", "user", "x") == 2
    assert o.check("group", "a", "member", "user", "y") == 255
    # errors under exclusion: NOT error = error, F AND error = F
    o, w = mk()
    for i in range(51):
        w(f"folder:f{i}#parent@folder:f{i+1}")
    w("folder:f0#viewer@user:v")
    assert o.check("folder", "f0", "view", "user", "v") == 2
    assert o.check("folder", "f0", "view", "user", "w") == 255
    assert o.check("folder", "f0", "not_view", "user", "v") == 255
    assert o.check("folder", "f0", "not_view", "user", "w") == 1


@pytest.mark.parametrize("impl", ["c", "mini"])
def test_expiration_and_wildcard_and_userset_subject(impl):
    schema = """
use expiration
definition user {}
definition group { relation member: user | group#member }
definition doc {
  relation viewer: user | user:* | group#member
  relation temp: user with expiration
  permission view = viewer + temp
}
"""
    if impl == "c":
        o = Oracle(schema)
        w = lambda r, e=0: o.touch(r, e)
    else:
        o = MiniOracle(schema)
        w = lambda r, e=0: o.write(r, e)
    w("doc:d1#temp@user:t", 1000)
    assert o.check("doc", "d1", "view", "user", "t", "", 999) == 2
    assert o.check("doc", "d1", "view", "user", "t", "", 1000) == 1
    assert o.lookup_resources("doc", "view", "user", "t", "", 999) == ["d1"]
    assert o.lookup_resources("doc", "view", "user", "t", "", 1001) == []
    w("doc:d2#viewer@user:*")
    assert o.check("doc", "d2", "view", "user", "anyone-at-all") == 2
    assert o.check("doc", "d2", "view", "group", "g", "member") == 1  # wildcard never matches usersets
    w("doc:d3#viewer@group:eng#member")
    w("group:eng#member@user:e1")
    assert o.check("doc", "d3", "view", "user", "e1") == 2
    assert o.check("doc", "d3", "view", "group", "eng", "member") == 2  # userset subject
    assert o.check("doc", "d3", "viewer", "group", "eng", "member") == 2
    assert o.check("doc", "d3", "view", "group", "ops", "member") == 1
    assert o.check("group", "eng", "member", "group", "eng", "member") == 2  # member of itself


SELF_SCHEMA = ("definition user {}\ndefinition doc {\n  relation viewer: user | doc#viewer\n  relation editor: user\n"
               "  relation banned: user | doc#viewer\n  permission view = viewer + editor\n  permission safe = view - banned\n}\n")
SELF_RELS = ["doc:1#viewer@user:a", "doc:2#editor@user:a", "doc:3#viewer@doc:5#viewer", "doc:7#banned@doc:7#viewer"]
# (resource type, permission, subject type, subject id, subject relation) -> expected ids
SELF_LOOKUPS = [
    (("doc", "view", "doc", "5", "viewer"), ["3", "5"]),    # doc:5#viewer is in doc:5#view (viewer is inlined into view)
    (("doc", "view", "doc", "9", "viewer"), ["9"]),          # never written, still a member of its own view
    (("doc", "viewer", "doc", "5", "viewer"), ["3", "5"]),
    (("doc", "view", "doc", "5", "editor"), []),             # doc:5#editor is in doc:5#view too...
    (("doc", "safe", "doc", "5", "viewer"), ["3", "5"]),
    (("doc", "safe", "doc", "7", "viewer"), []),             # ...but doc:7#viewer is banned from doc:7#safe
]


def test_lookup_resources_includes_the_userset_subject_itself_when_check_says_so():
    """LookupResources must agree with Check for userset subjects: T:x#r is a member of T:x#P for every relation r
    inlined into P's union, with or without relationships (round-1 advisor finding). Both oracles, and the GPU
    test of the same table in test_zz_gpu_new_paths.py."""
    from oracle.mini_oracle import MiniOracle
    from oracle.pyoracle import Oracle

    c, m = Oracle(SELF_SCHEMA), MiniOracle(SELF_SCHEMA)
    for r in SELF_RELS:
        c.touch(r), m.write(r)
    for (rt, perm, st, sid, srel), want in SELF_LOOKUPS:
        if srel == "editor":  # Check(doc:5#view@doc:5#editor) is HAS: the subject itself belongs to the answer
            want = ["5"]
        a, b = sorted(c.lookup_resources(rt, perm, st, sid, srel)), sorted(m.lookup_resources(rt, perm, st, sid, srel))
        assert a == b == want, f"{rt}#{perm}@{st}:{sid}#{srel}: C={a} mini={b} want={want}"
        for rid in ("3", "5", "7", "9"):
            assert (c.check(rt, rid, perm, st, sid, srel) == 2) == (rid in a)
