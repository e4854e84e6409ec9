"""Shared cases for the cooperative two-level meet (csrc/kernels.cuh coop_l2): object ids that collide in the filter
the kernel keeps over a range's children (same word, same two bits), so that "the filter lets it through" and "it is a
child of the range" can be told apart by a test."""
from collections import defaultdict

L2_SCHEMA = """definition user {}
definition group { relation member: user }
definition team { relation member: group#member }
definition namespace { relation viewer: team#member  permission view = viewer }"""
N_TEAMS = 4000


def filter_key(obj_id: int):
    """(word, mask) of an object id in the 64-word blocked filter: must restate kernels.cuh (ZG_L2_FILTER)."""
    hv = (obj_id * 0x9E3779B1) & 0xFFFFFFFF
    return hv >> 26, (1 << ((hv >> 21) & 31)) | (1 << ((hv >> 16) & 31))


def padding_rels():
    """Interns team:t0 .. t3999 (as resources of one relationship each)."""
    return [f"team:t{i}#member@group:gpad#member" for i in range(N_TEAMS)]


def collision_cases(team_id, max_pairs=24):
    """team_id(name) -> interned id. Returns (rels, [(check, expected v1 code)])."""
    by_key = defaultdict(list)
    for i in range(N_TEAMS):
        by_key[filter_key(team_id(f"t{i}"))].append(i)
    pairs = [v[:2] for v in by_key.values() if len(v) > 1][:max_pairs]
    assert len(pairs) >= 8, "no colliding ids: the filter's hash changed? update filter_key()"
    rels, checks = [], []
    for k, (a, b) in enumerate(pairs):
        others = [(a * 7 + j * 13 + 1) % N_TEAMS for j in range(k % 5)]  # a few more children around the colliding one
        others = [t for t in others if t not in (a, b)]
        for t in [a] + others:
            rels.append(f"namespace:n{k}#viewer@team:t{t}#member")   # the range holds a, not b
        rels.append(f"namespace:p{k}#viewer@team:t{b}#member")       # control: this range holds b
        rels.append(f"team:t{b}#member@group:g{k}#member")           # the subject reaches b only
        rels.append(f"group:g{k}#member@user:u{k}")
        checks.append((f"namespace:n{k}#view@user:u{k}", 1))         # filter says maybe, the search says no
        checks.append((f"namespace:p{k}#view@user:u{k}", 2))
    return rels, checks
