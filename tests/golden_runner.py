"""Runs the reference-derived golden cases (tests/golden/reference_cases.json)
against any backend exposing write/check/bulk/lookup/read. The same runner drives
the C oracle, the pure-Python mini oracle and the CUDA engine's client."""
CODES = {"NO": 1, "HAS": 2, "ERROR": 255}


def split_rel(rel):
    left, right = rel.split("@", 1)
    rt, rest = left.split(":", 1)
    rid, perm = rest.rsplit("#", 1)
    st, srest = right.split(":", 1)
    sid, _, srel = srest.partition("#")
    return rt, rid, perm, st, sid, srel


def run_case(backend, case):
    """backend: object with write(rel), check(rt,rid,perm,st,sid,srel)->code,
    bulk([rel...])->[code], lookup(rt,perm,st,sid,srel)->[ids], read(**filter)->[rel]."""
    for i, step in enumerate(case["steps"]):
        where = f"{case['id']} step {i} ({case['cite']})"
        if "write" in step:
            backend.write(step["write"])
        elif "check" in step:
            got = backend.check(*split_rel(step["check"]))
            assert got == CODES[step["expect"]], f"{where}: {step['check']} -> {got}"
        elif "bulk" in step:
            got = list(backend.bulk(step["bulk"]))
            assert got == [CODES[e] for e in step["expect"]], f"{where}: {got}"
        elif "lookup" in step:
            rt, rid, perm, st, sid, srel = split_rel(step["lookup"])
            assert rid == "$"
            got = sorted(backend.lookup(rt, perm, st, sid, srel))
            assert got == sorted(step["expect"]), f"{where}: {got}"
        elif "read" in step:
            got = sorted(backend.read(**step["read"]))
            assert got == sorted(step["expect"]), f"{where}: {got}"
        else:
            raise AssertionError(f"unknown step {step}")
