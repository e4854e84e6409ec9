"""Random Zanzibar schemas + relationship graphs for cross-checking implementations.

Used by: test_oracle_random.py (C oracle vs pure-Python oracle) and the GPU parity
tests (CUDA engine vs C oracle). Everything is seeded and deterministic.

Data is generated so that recursive usersets / arrows are ACYCLIC (edges go from a
higher object index to a lower one): forward evaluation with the reference's
cache-off configuration (pkg/spicedb/spicedb.go:44-46) is exponential on cyclic
data up to the depth cap. Cycles and the depth cap get their own small tests.
"""
from __future__ import annotations

import random

FIXED_SCHEMAS = {
    # SURVEY.md 8(d) cfg3
    "nesting": """
definition user {}
definition group { relation member: user | group#member }
definition team { relation member: group#member | user }
definition namespace {
  relation viewer: team#member | user
  permission view = viewer
}
""",
    # SURVEY.md 8(d) cfg4
    "docs": """
definition user {}
definition group { relation member: user | group#member }
definition org { relation member: user | group#member }
definition folder {
  relation parent: folder
  relation owner: user | group#member
  relation viewer: user | group#member
  permission view = viewer + owner + parent->view
}
definition document {
  relation parent: folder
  relation org: org
  relation owner: user | group#member | user:*
  relation editor: user | group#member | user:*
  relation viewer: user | group#member | user:*
  relation banned: user
  permission edit = owner + editor
  permission view = (viewer + edit + parent->view) - banned
  permission restricted_view = view & org->member
}
""",
    # non-pure permissions reached THROUGH arrows and usersets (sub-query passes)
    "deep_nonpure": """
definition user {}
definition group {
  relation member: user | group#member | group#active
  relation suspended: user
  permission active = member - suspended
}
definition folder {
  relation parent: folder
  relation viewer: user | group#active | group#member
  relation banned: user | group#member
  relation auditor: user
  permission view = (viewer + parent->view) - banned
  permission audit = auditor & view
  permission any = view + audit + parent->audit
}
definition document {
  relation folder: folder
  relation reader: user | folder#view | group#active
  permission read = reader + folder->view
  permission strict = folder->audit & reader
  permission nobody = nil
  permission weird = (read - folder->audit) + strict
}
""",
}


def random_schema(rng: random.Random):
    """-> (schema_text, model) where model describes types for data generation."""
    n_types = rng.randint(2, 4)
    tnames = [f"t{i}" for i in range(n_types)]
    perm_pool = ["p0", "p1", "p2"]
    types = {}
    for ti, t in enumerate(tnames):
        rels = {}
        for ri in range(rng.randint(1, 3)):
            allowed = []
            if rng.random() < 0.8:
                allowed.append(("user", None))
            if rng.random() < 0.25:
                allowed.append(("user", "*"))
            for _ in range(rng.randint(0, 2)):
                ot = rng.choice(tnames)
                allowed.append((ot, "?"))  # subject relation resolved below
            if not allowed:
                allowed.append(("user", None))
            rels[f"r{ri}"] = allowed
        # tupleset relations (object subjects of another type, no relation)
        for ai in range(rng.randint(0, 2)):
            rels[f"a{ai}"] = [(rng.choice(tnames), None)]
        types[t] = {"rels": rels, "perms": {}}
    # permissions: expressions over relations, earlier permissions, arrows
    for t in tnames:
        d = types[t]
        names_so_far = list(d["rels"].keys())
        arrows = [r for r in d["rels"] if r.startswith("a")]
        for p in perm_pool[: rng.randint(1, 3)]:
            def leaf():
                x = rng.random()
                if arrows and x < 0.3:
                    return f"{rng.choice(arrows)}->{rng.choice(perm_pool + ['r0'])}"
                if x < 0.35:
                    return "nil"
                return rng.choice(names_so_far)

            def expr(depth):
                if depth == 0 or rng.random() < 0.35:
                    return leaf()
                op = rng.choice(["+", "+", "&", "-"])
                return f"({expr(depth - 1)} {op} {expr(depth - 1)})"

            d["perms"][p] = expr(rng.randint(1, 3))
            names_so_far.append(p)
    # resolve userset subject relations: any relation/permission of the subject type
    for t in tnames:
        for r, allowed in types[t]["rels"].items():
            out = []
            for (ot, sr) in allowed:
                if sr == "?":
                    cands = list(types[ot]["rels"].keys()) + list(types[ot]["perms"].keys())
                    sr = rng.choice(cands)
                if (ot, sr) not in out:
                    out.append((ot, sr))
            types[t]["rels"][r] = out
    lines = ["definition user {}"]
    for t in tnames:
        lines.append(f"definition {t} {{")
        for r, allowed in types[t]["rels"].items():
            parts = [ot + (":*" if sr == "*" else (f"#{sr}" if sr else "")) for ot, sr in allowed]
            lines.append(f"  relation {r}: {' | '.join(parts)}")
        for p, e in types[t]["perms"].items():
            lines.append(f"  permission {p} = {e}")
        lines.append("}")
    return "\n".join(lines) + "\n", {"types": types, "tnames": tnames}


def model_from_schema(schema_text):
    """Derive the data-generation model from schema text (uses the mini parser)."""
    from oracle.mini_oracle import parse_schema

    defs = parse_schema(schema_text)
    types = {}
    for t, d in defs.items():
        rels = {r: [(st, sr) for (st, sr, _ex) in allowed] for r, allowed in d["relations"].items()}
        types[t] = {"rels": rels, "perms": {p: None for p in d["permissions"]}}
    return {"types": types, "tnames": [t for t in types]}


def random_relationships(rng: random.Random, model, n_obj=6, n_user=6, density=0.35, acyclic=True):
    """-> sorted list of relationship strings. Object ids are o0..o{n-1}, users u0.."""
    rels = set()
    for t, d in model["types"].items():
        for r, allowed in d["rels"].items():
            for (st, sr) in allowed:
                for oi in range(n_obj):
                    if sr == "*":
                        if rng.random() < density * 0.3:
                            rels.add(f"{t}:o{oi}#{r}@{st}:*")
                        continue
                    n_subj = n_user if st == "user" else n_obj
                    for si in range(n_subj):
                        if rng.random() >= density / (1 if st == "user" else 1.5):
                            continue
                        if st != "user" and acyclic and si >= oi:
                            continue  # edges only point to strictly lower indices
                        sname = f"u{si}" if st == "user" else f"o{si}"
                        rels.add(f"{t}:o{oi}#{r}@{st}:{sname}" + (f"#{sr}" if sr else ""))
    return sorted(rels)


def random_checks(rng: random.Random, model, n, n_obj=6, n_user=6):
    """-> list of 'type:id#perm@stype:sid[#srel]' (some with userset subjects,
    some naming unknown objects)."""
    out = []
    tn = [t for t in model["tnames"] if model["types"][t]["rels"] or model["types"][t]["perms"]]
    for _ in range(n):
        t = rng.choice(tn)
        d = model["types"][t]
        name = rng.choice(list(d["rels"].keys()) + list(d["perms"].keys()))
        oid = f"o{rng.randint(0, n_obj)}"  # o{n_obj} never written: unknown object
        if rng.random() < 0.85 or not tn:
            subj = f"user:u{rng.randint(0, n_user)}"
        else:
            st = rng.choice(tn)
            sd = model["types"][st]
            sr = rng.choice(list(sd["rels"].keys()) + list(sd["perms"].keys()))
            subj = f"{st}:o{rng.randint(0, n_obj - 1)}#{sr}"
        out.append(f"{t}:{oid}#{name}@{subj}")
    return out
