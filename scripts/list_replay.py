"""Filtered kube lists end to end at the C ABI (BASELINE config 5's post-filter shape):
concurrent clients each hand one List body to zg_list_postfilter -- scan, resolve, ONE bulk check on the GPU,
splice -- against a store whose objects carry real "namespace/name" ids. Reports filtered lists/s and body MB/s,
and checks a few outputs against the Python mirror of pkg/authz/postfilter.go.

    python scripts/list_replay.py [--pods 200000] [--items 10000] [--clients 32] [--rounds 4] [--devices 1] [--dry-run]

The clients are native threads (tests/cabi/loadgen.c): Python threads would measure the GIL. --devices N: ONE engine
handle owning N GPUs.

--dry-run builds the store and the bodies and runs the host-only stages (scan, resolve), no GPU call.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zgpu  # noqa: E402
from spicedb_kubeapi_proxy_b200 import _lib, postfilter as pf  # noqa: E402

SCHEMA = """
definition user {}
definition group { relation member: user | group#member }
definition namespace { relation viewer: user | group#member  permission view = viewer }
definition pod {
  relation namespace: namespace
  relation viewer: user | group#member | user:*
  relation banned: user
  permission view = (viewer + namespace->view) - banned
}
"""

ap = argparse.ArgumentParser()
ap.add_argument("--pods", type=int, default=200000)
ap.add_argument("--namespaces", type=int, default=200)
ap.add_argument("--users", type=int, default=2000)
ap.add_argument("--groups", type=int, default=200)
ap.add_argument("--items", type=int, default=10000)
ap.add_argument("--clients", type=int, default=32)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--dry-run", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(7)

e = zgpu.Engine(SCHEMA, host_only=a.dry_run, n_devices=1 if a.dry_run else a.devices)
t0 = time.perf_counter()
pod_names = [f"ns-{i % a.namespaces}/pod-{i}" for i in range(a.pods)]
pod_id = np.array([e.intern("pod", n) for n in pod_names], dtype=np.uint32)
ns_id = np.array([e.intern("namespace", f"ns-{i}") for i in range(a.namespaces)], dtype=np.uint32)
user_id = np.array([e.intern("user", f"user-{i}") for i in range(a.users)], dtype=np.uint32)
group_id = np.array([e.intern("group", f"group-{i}") for i in range(a.groups)], dtype=np.uint32)
e.add_bulk("pod", "namespace", "namespace", pod_id, ns_id[np.arange(a.pods) % a.namespaces])
k = a.pods // 2
e.add_bulk("pod", "viewer", "user", pod_id[rng.integers(0, a.pods, k)], user_id[rng.integers(0, a.users, k)])
e.add_bulk("pod", "viewer", "group", pod_id[rng.integers(0, a.pods, k // 4)], group_id[rng.integers(0, a.groups, k // 4)], srel="member")
e.add_bulk("pod", "banned", "user", pod_id[rng.integers(0, a.pods, k // 8)], user_id[rng.integers(0, a.users, k // 8)])
e.add_bulk("namespace", "viewer", "user", ns_id[rng.integers(0, a.namespaces, a.namespaces * 4)], user_id[rng.integers(0, a.users, a.namespaces * 4)])
e.add_bulk("namespace", "viewer", "group", ns_id[rng.integers(0, a.namespaces, a.namespaces)], group_id[rng.integers(0, a.groups, a.namespaces)], srel="member")
e.add_bulk("group", "member", "user", group_id[rng.integers(0, a.groups, a.users * 2)], user_id[rng.integers(0, a.users, a.users * 2)])
e.publish()
build_s = time.perf_counter() - t0


def pod_json(i):
    ns, name = pod_names[i].split("/")
    return {"apiVersion": "v1", "kind": "Pod",
            "metadata": {"name": name, "namespace": ns, "uid": f"{i:032x}", "resourceVersion": str(10**6 + i),
                         "labels": {"app": f"svc-{i % 50}", "tier": "backend"},
                         "annotations": {"checksum/config": "9f86d081884c7d659a2feaa0c55ad015a3bf4f1b2b0b822cd15d6c15b0f00a08"}},
            "spec": {"containers": [{"name": "main", "image": "registry.example/app:1.2.3", "args": ["--port=8080"] * 4,
                                     "env": [{"name": f"VAR_{k}", "value": "v" * 24} for k in range(12)],
                                     "resources": {"limits": {"cpu": "500m", "memory": "512Mi"}}}] * 2,
                     "nodeName": f"node-{i % 300}"},
            "status": {"phase": "Running", "podIP": "10.0.0.1",
                       "conditions": [{"type": t, "status": "True"} for t in ("Initialized", "Ready", "ContainersReady", "PodScheduled")]}}


def make_body():
    pick = rng.integers(0, a.pods, a.items)
    return json.dumps({"kind": "PodList", "apiVersion": "v1", "metadata": {"resourceVersion": "1"},
                       "items": [pod_json(int(i)) for i in pick]}, separators=(",", ":")).encode()


n_bodies = min(a.clients, 8)  # bodies are shared between clients; the subjects differ
bodies = [make_body() for _ in range(n_bodies)]
users = [f"user-{int(u)}" for u in rng.integers(0, a.users, a.clients)]
tpls = [e.list_template("pod", "view", "user", u) for u in users]
res = {"pods": a.pods, "items_per_list": a.items, "clients": a.clients, "body_mb": round(len(bodies[0]) / 1e6, 2),
       "tuples": int(e.num_tuples()), "store_build_s": round(build_s, 2)}

# host-only stages, one thread
t = time.perf_counter()
items, ib, ie = _lib.list_scan(bodies[0])
res["scan_ms"] = round((time.perf_counter() - t) * 1e3, 2)
t = time.perf_counter()
checks, checked = e.list_resolve(bodies[0], items, tpls[0])
res["resolve_ms"] = round((time.perf_counter() - t) * 1e3, 2)
assert checked.all() and (checks["res"] != 0xFFFFFFFF).all()
if a.dry_run:
    print(json.dumps(res))
    sys.exit(0)

# the mirror (Python loop over items, strings across the boundary) on a few lists = the checker
client = zgpu.client.PermissionsClient(SCHEMA, engine=e)
for i in range(min(3, a.clients)):
    want = pf.filter_list_response(bodies[i % n_bodies], ["pod:{{namespacedName}}#view@user:{{user.name}}"],
                                   pf.RequestInfo(), pf.UserInfo(name=users[i]), client)
    got = e.list_postfilter(bodies[i % n_bodies], [tpls[i]])
    assert got == want, f"client {i}: fused output differs from the mirror"
kept = len(json.loads(got)["items"] or [])
res["kept_of_last_checked_list"] = kept

# ---- native client threads
import ctypes as C  # noqa: E402
import subprocess  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spicedb-kubeapi-proxy_b200")
lg_so = os.path.join(ROOT, "tests", "cabi", "libloadgen.so")
subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"),
                os.path.join(ROOT, "tests", "cabi", "loadgen.c"), "-o", lg_so,
                # the library this process already loaded (ZGPU_LIB selects a tuning variant)
                os.environ.get("ZGPU_LIB") or os.path.join(PKG, "libzgpu.so"), "-Wl,-rpath," + PKG, "-lpthread"], check=True)
LG = C.CDLL(lg_so)


class ListClient(C.Structure):
    _fields_ = [("e", C.c_void_p), ("body", C.c_char_p), ("body_len", C.c_uint64), ("tpl", _lib._ListTemplate),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("out_len", C.c_uint64), ("rounds", C.c_int), ("rc", C.c_int)]


LG.loadgen_run_lists.restype = C.c_double
LG.loadgen_run_lists.argtypes = [C.POINTER(ListClient), C.c_int, C.c_int]
out_cap = max(len(b) for b in bodies) + 8
outs = np.empty((a.clients, out_cap), dtype=np.uint8)


def run_phase(mode, rounds):
    cl = (ListClient * a.clients)()
    for i in range(a.clients):
        b = bodies[i % n_bodies]
        cl[i].e, cl[i].body, cl[i].body_len, cl[i].tpl = e._h, b, len(b), tpls[i]
        cl[i].out, cl[i].out_cap, cl[i].rounds = outs[i].ctypes.data, out_cap, rounds
    s0 = e.stats()
    dt = LG.loadgen_run_lists(cl, a.clients, mode)
    bad = [(i, c.rc) for i, c in enumerate(cl) if c.rc]
    assert dt > 0 and not bad, bad[:5]
    return dt, s0, e.stats(), [int(c.out_len) for c in cl]


run_phase(0, 1)  # warm-up
dt, s0, s1, lens = run_phase(0, a.rounds)
for i in range(min(3, a.clients)):  # the concurrent answers are the single-caller answers
    assert outs[i, :lens[i]].tobytes() == e.list_postfilter(bodies[i % n_bodies], [tpls[i]]), f"client {i}: concurrent output differs"
n_lists = a.clients * a.rounds
res.update({"clients_are": "native threads (tests/cabi/loadgen.c)", "devices": int(s1["devices"]),
            "filtered_lists_per_s": round(n_lists / dt, 1), "body_MB_per_s": round(n_lists * len(bodies[0]) / 1e6 / dt, 1),
            "checks_per_s": round(n_lists * a.items / dt), "launches": int(s1["launches"] - s0["launches"]),
            "coalesced_requests": int(s1["coalesced_requests"] - s0["coalesced_requests"])})
# the pre-filter shape: LookupResources + scan + keep + splice per list (pkg/authz/lookups.go:65)
for i in range(min(2, a.clients)):
    r = pf.run_lookup_resources(client, ("pod", "$", "view", "user", users[i], ""), pf.RequestInfo())
    assert e.list_prefilter(bodies[i % n_bodies], tpls[i]) == pf.filter_list(bodies[i % n_bodies], r), f"client {i}: prefilter differs"
run_phase(1, 1)
dt, s0, s1, lens = run_phase(1, a.rounds)
res.update({"prefiltered_lists_per_s": round(n_lists / dt, 1), "prefilter_launches": int(s1["launches"] - s0["launches"]),
            "prefilter_lookup_batches": int(s1["lookup_batches"] - s0["lookup_batches"])})
print(json.dumps(res))
