"""Pure-Python restatement of the same Check semantics, for SMALL cases only.

TEST INFRASTRUCTURE. A third, independently written implementation (strings and
Python sets, its own schema parser) used to cross-check oracle/zanzibar_oracle.c
on the golden cases and on random small graphs. Semantics: see zanzibar_oracle.c
header (SpiceDB v1.47.1 documented behaviour; reference call sites
pkg/authz/check.go:23-69, lookups.go:49-88, pkg/spicedb/spicedb.go:33,47).
"""
from __future__ import annotations

import re

NO, HAS, ERR = 1, 2, 255
F, T, E = 0, 1, 2
MAX_DEPTH = 50

_TOK = re.compile(r"\s*(?://[^\n]*|/\*.*?\*/|(->|[{}:|#=+&\-()*.]|[A-Za-z_][A-Za-z0-9_/]*))", re.S)


def _tokens(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOK.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"bad schema near {text[pos:pos+20]!r}")
        pos = m.end()
        if m.group(1):
            out.append(m.group(1))
    return out


class _P:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        v = self.peek()
        self.i += 1
        return v

    def expect(self, v):
        if self.next() != v:
            raise ValueError(f"expected {v!r} at token {self.i}")

    # '+' tightest, then '&', then '-' ; all left associative
    def expr(self, lvl=0):
        ops = ["-", "&", "+"]
        if lvl == 3:
            return self.primary()
        l = self.expr(lvl + 1)
        while self.peek() == ops[lvl]:
            self.next()
            l = (ops[lvl], l, self.expr(lvl + 1))
        return l

    def primary(self):
        t = self.next()
        if t == "(":
            e = self.expr()
            self.expect(")")
            return e
        if t == "nil":
            return ("nil",)
        if self.peek() == "->":
            self.next()
            return ("arrow", t, self.next())
        if self.peek() == ".":
            self.next()
            if self.next() != "any":
                raise ValueError("only .any() supported")
            self.expect("(")
            p = self.next()
            self.expect(")")
            return ("arrow", t, p)
        return ("ref", t)


def parse_schema(text):
    """-> {type: {"relations": {name: [(stype, srel|None|'*', expiry)]}, "permissions": {name: expr}}}"""
    p = _P(_tokens(text))
    defs = {}
    while p.peek() is not None:
        t = p.next()
        if t == "use":
            p.next()
            continue
        if t != "definition":
            raise ValueError(f"unexpected {t!r}")
        name = p.next()
        d = defs[name] = {"relations": {}, "permissions": {}}
        p.expect("{")
        while p.peek() != "}":
            k = p.next()
            if k == "relation":
                rn = p.next()
                p.expect(":")
                allowed = []
                while True:
                    st, sr, ex = p.next(), None, False
                    if p.peek() == ":":
                        p.next()
                        p.expect("*")
                        sr = "*"
                    elif p.peek() == "#":
                        p.next()
                        sr = p.next()
                    if p.peek() == "with":
                        p.next()
                        if p.next() != "expiration":
                            raise ValueError("caveats unsupported")
                        ex = True
                    allowed.append((st, sr, ex))
                    if p.peek() != "|":
                        break
                    p.next()
                d["relations"][rn] = allowed
            elif k == "permission":
                pn = p.next()
                p.expect("=")
                d["permissions"][pn] = p.expr()
            else:
                raise ValueError(f"unexpected {k!r}")
        p.expect("}")
    return defs


class MiniOracle:
    def __init__(self, schema: str):
        self.defs = parse_schema(schema)
        # (rtype, rid, rel) -> {(stype, sid, srel|None): expires_at}
        self.rows = {}

    def write(self, rel: str, expires_at=0, delete=False):
        m = re.match(r"^(.*?):(.*?)#(.*?)@(.*?):(.*?)(#(.*?))?$", rel)
        rt, rid, rl, st, sid, _, srel = m.groups()
        if rl not in self.defs[rt]["relations"]:
            raise ValueError(f"unknown relation {rt}#{rl}")
        key = (st, sid, srel or None)
        row = self.rows.setdefault((rt, rid, rl), {})
        if delete:
            row.pop(key, None)
        else:
            row[key] = expires_at

    def _live(self, rt, rid, rl, now):
        return [k for k, ex in self.rows.get((rt, rid, rl), {}).items() if ex == 0 or ex > now]

    def _check(self, rt, rid, name, subj, depth, now):
        st, sid, srel = subj
        if srel is not None and (rt, rid, name) == (st, sid, srel):
            return T
        d = self.defs[rt]
        if name in d["relations"]:
            subs = self._live(rt, rid, name, now)
            r = F
            if srel is None:
                if (st, sid, None) in subs or (st, "*", None) in subs:
                    return T
            for (ut, uid, urel) in subs:
                if urel is None:
                    continue
                v = E if depth + 1 > MAX_DEPTH else self._check(ut, uid, urel, subj, depth + 1, now)
                if v == T:
                    return T
                r = max(r, v)  # E(2) > F(0)
            return r
        return self._eval(d["permissions"][name], rt, rid, subj, depth, now)

    def _eval(self, e, rt, rid, subj, depth, now):
        k = e[0]
        if k == "nil":
            return F
        if k == "ref":
            return self._check(rt, rid, e[1], subj, depth, now)
        if k == "arrow":
            r = F
            for (xt, xid, _xrel) in self._live(rt, rid, e[1], now):
                if xid == "*":
                    continue
                xd = self.defs[xt]
                if e[2] not in xd["relations"] and e[2] not in xd["permissions"]:
                    continue
                v = E if depth + 1 > MAX_DEPTH else self._check(xt, xid, e[2], subj, depth + 1, now)
                if v == T:
                    return T
                r = max(r, v)
            return r
        a = self._eval(e[1], rt, rid, subj, depth, now)
        b = self._eval(e[2], rt, rid, subj, depth, now)
        if k == "+":
            return T if T in (a, b) else (E if E in (a, b) else F)
        if k == "-":
            b = {T: F, F: T, E: E}[b]
        return F if F in (a, b) else (E if E in (a, b) else T)

    def check(self, rt, rid, perm, st, sid, srel="", now=0):
        d = self.defs.get(rt)
        if d is None or st not in self.defs or (perm not in d["relations"] and perm not in d["permissions"]):
            return ERR
        if srel in ("", "..."):
            srel = None
        if srel is not None and srel not in self.defs[st]["relations"] and srel not in self.defs[st]["permissions"]:
            return ERR
        v = self._check(rt, rid, perm, (st, sid, srel), 0, now)
        return {F: NO, T: HAS, E: ERR}[v]

    def lookup_resources(self, rt, perm, st, sid, srel="", now=0):
        cands = {rid for (t, rid, rl) in self.rows if t == rt and self._live(t, rid, rl, now)}
        if srel not in ("", "...", None) and st == rt:
            cands.add(sid)  # a userset subject rt:x#r may be a member of rt:x#perm with no relationship at all
        return [r for r in sorted(cands) if self.check(rt, r, perm, st, sid, srel, now) == HAS]
