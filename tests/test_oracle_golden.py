"""The oracle is pinned here: every golden case the reference's own tests hold for
this path must hold for oracle/zanzibar_oracle.c and for
the independent pure-Python restatement oracle/mini_oracle.py."""
import pytest

from golden_runner import run_case, split_rel
from oracle.mini_oracle import MiniOracle
from oracle.pyoracle import Oracle


class COracleBackend:
    def __init__(self, schema):
        self.o = Oracle(schema)

    def write(self, rel):
        self.o.touch(rel)

    def check(self, *a):
        return self.o.check(*a)

    def bulk(self, rels):
        return [self.o.check(*split_rel(r)) for r in rels]

    def lookup(self, rt, perm, st, sid, srel):
        return self.o.lookup_resources(rt, perm, st, sid, srel)

    def read(self, **f):
        return self.o.read(**f)


class MiniBackend:
    def __init__(self, schema):
        self.o = MiniOracle(schema)

    def write(self, rel):
        self.o.write(rel)

    def check(self, *a):
        return self.o.check(*a)

    def bulk(self, rels):
        return [self.o.check(*split_rel(r)) for r in rels]

    def lookup(self, rt, perm, st, sid, srel):
        return self.o.lookup_resources(rt, perm, st, sid, srel)

    def read(self, res_type="", res_id="", rel="", **_):
        out = []
        for (rt, rid, rl), subs in self.o.rows.items():
            if res_type and rt != res_type or res_id and rid != res_id or rel and rl != rel:
                continue
            for (st, sid, srel) in subs:
                out.append(f"{rt}:{rid}#{rl}@{st}:{sid}" + (f"#{srel}" if srel else ""))
        return out


def _cases(golden):
    return golden["cases"]


@pytest.mark.parametrize("backend_cls", [COracleBackend, MiniBackend])
def test_golden_cases(golden, backend_cls):
    for case in _cases(golden):
        run_case(backend_cls(golden["schemas"][case["schema"]]), case)
