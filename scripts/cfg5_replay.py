"""BASELINE config 5 replayed at the C ABI: many concurrent clients, each submitting one
list-sized CheckBulkPermissions (the post-filter shape, pkg/authz/postfilter.go:127-134) and one
LookupResources (the pre-filter shape, pkg/authz/lookups.go:65) against the cfg4 store.
The true end-to-end path (Go HTTP handlers) cannot run here (no Go toolchain); this drives the
same library entry points from threads (ctypes releases the GIL inside the calls).

    python scripts/cfg5_replay.py [--scale 1.0] [--clients 1000] [--items 10000] [--rounds 2] [--devices 8]

--devices N: ONE engine handle owning N GPUs (zg_config.n_devices: a replica per device, every bulk call and every
batch of lookups split across them) -- BASELINE config 5 names 8 GPUs behind the proxy's single client.
A third phase runs both shapes at once (the proxy's pre-filter goroutine runs concurrently with the list request,
pkg/authz/responsefilterer.go:165-183).
"""
import argparse, json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zgpu
from spicedb_kubeapi_proxy_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--clients", type=int, default=1000)
ap.add_argument("--items", type=int, default=10000)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--devices", type=int, default=1)
a = ap.parse_args()

w = workloads.cfg4(scale=a.scale)
e = zgpu.Engine(w.schema, device=0, n_devices=a.devices)
w.load_into(e)
e.publish()
base = w.check_items(e, zgpu.CHECK_DTYPE)
rng = np.random.default_rng(5)
n_users = int(max(g.subj.max() for g in w.groups if g.subj_type == "user" and not g.wildcard)) + 1
n_docs = int(max(g.res.max() for g in w.groups if g.res_type == "document")) + 1
view = e.slot_id("document", "view")

def make_list(user):  # one user lists 10k documents: one post-filter item per document
    it = np.zeros(a.items, dtype=zgpu.CHECK_DTYPE)
    it["res"] = rng.integers(0, n_docs, a.items)
    it["subj"] = user
    it["perm"], it["stype"], it["srel"] = view, e.type_id("user"), 0xFFFF
    return it

users = rng.integers(0, n_users, a.clients)
lists = [make_list(int(u)) for u in users]
ref = [e.check_bulk(l) for l in lists[:4]]  # single-caller answers for a spot check
errors, kept, found = [], [0] * a.clients, [0] * a.clients


def run_phase(fn):
    barrier = threading.Barrier(a.clients + 1)

    def body(i):
        barrier.wait()
        fn(i)

    th = [threading.Thread(target=body, args=(i,)) for i in range(a.clients)]
    [t.start() for t in th]
    s0 = e.stats()
    barrier.wait()
    t0 = time.perf_counter()
    [t.join() for t in th]
    return time.perf_counter() - t0, s0, e.stats()


def check_phase(i):  # post-filter shape
    out = np.empty(a.items, dtype=np.uint8)
    for _ in range(a.rounds):
        e.check_bulk(lists[i], out)
    kept[i] = int((out == 2).sum())
    if i < 4 and not np.array_equal(out, ref[i]):
        errors.append(i)


def lookup_phase(i):  # pre-filter shape
    found[i] = int(e.lookup_resources_ids("document", "view", "user", int(users[i])).size)


def mixed_phase(i):  # a list request: its pre-filter lookup and its post-filter bulk check
    found[i] = int(e.lookup_resources_ids("document", "view", "user", int(users[i])).size)
    out = np.empty(a.items, dtype=np.uint8)
    e.check_bulk(lists[i], out)


dt_c, c0, c1 = run_phase(check_phase)
users = rng.integers(0, n_users, a.clients)  # fresh subjects: nothing comes from the answer cache
dt_l, l0, l1 = run_phase(lookup_phase)
users = rng.integers(0, n_users, a.clients)
dt_m, m0, m1 = run_phase(mixed_phase)
lists_done = a.clients * a.rounds
print(json.dumps({
    "workload": f"cfg5 replay at the C ABI on {w.note}", "clients": a.clients, "items_per_list": a.items, "rounds": a.rounds,
    "postfilter": {"filtered_lists_per_s": lists_done / dt_c, "Mchecks_per_s": lists_done * a.items / dt_c / 1e6,
                   "wall_s": dt_c, "kept_per_list_mean": float(np.mean(kept)), "parity_spot_check_ok": not errors,
                   "kernel_launches": c1["launches"] - c0["launches"],
                   "coalesced": {"requests": c1["coalesced_requests"] - c0["coalesced_requests"],
                                 "launches": c1["coalesced_launches"] - c0["coalesced_launches"]}},
    "prefilter": {"lookups_per_s": a.clients / dt_l, "wall_s": dt_l, "results_per_lookup_mean": float(np.mean(found)),
                  "results_per_s_M": float(np.sum(found)) / dt_l / 1e6, "kernel_launches": l1["launches"] - l0["launches"],
                  "batches": l1["lookup_batches"] - l0["lookup_batches"],
                  "lookups_in_batches": l1["lookups_batched"] - l0["lookups_batched"]},
    "mixed": {"filtered_lists_per_s": a.clients / dt_m, "wall_s": dt_m,
              "what": "per client: one LookupResources + one 10k-item bulk check, all clients at once"},
    "devices": int(m1["devices"]), "store_tuples": int(m1["tuples"]),
    "note": "post-filter: one 10k-item zg_check_bulk per list, concurrent callers coalesced by the library's batcher; "
            "pre-filter: one zg_lookup_resources per client, concurrent calls answered up to 64 per launch sequence "
            "(multi-source reverse walk + one verification launch). Python threads drive the ABI (GIL released "
            "inside calls); the Go HTTP path cannot run here."}))
