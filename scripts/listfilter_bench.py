"""Host-side measurement of the list-response filter: bytes of list body per
second through zg_list_scan + zg_list_filter, beside a decode/encode round trip of the same body with
Python's C-accelerated json module (the reference's shape of work, postfilter.go:19-47: full
Unmarshal into maps, full Marshal back; Go's encoding/json cannot run here). CPU only, no GPU."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import zgpu  # noqa: F401,E402
from spicedb_kubeapi_proxy_b200 import _lib  # noqa: E402


def pod(i):
    return {"apiVersion": "v1", "kind": "Pod",
            "metadata": {"name": f"pod-{i}", "namespace": f"ns-{i % 200}", "uid": f"{i:032x}", "resourceVersion": str(10**6 + i),
                         "creationTimestamp": "2026-01-01T00:00:00Z", "labels": {"app": f"svc-{i % 50}", "tier": "backend"},
                         "annotations": {"checksum/config": "9f86d081884c7d659a2feaa0c55ad015a3bf4f1b2b0b822cd15d6c15b0f00a08"},
                         "ownerReferences": [{"apiVersion": "apps/v1", "kind": "ReplicaSet", "name": f"rs-{i % 500}", "uid": "x" * 36}]},
            "spec": {"containers": [{"name": "main", "image": "registry.example/app:1.2.3", "args": ["--port=8080"] * 4,
                                     "env": [{"name": f"VAR_{k}", "value": "v" * 24} for k in range(12)],
                                     "resources": {"limits": {"cpu": "500m", "memory": "512Mi"}},
                                     "volumeMounts": [{"name": "cfg", "mountPath": "/etc/cfg"}]}] * 2,
                     "nodeName": f"node-{i % 300}", "volumes": [{"name": "cfg", "configMap": {"name": "cfg"}}]},
            "status": {"phase": "Running", "podIP": "10.0.0.1", "conditions": [{"type": t, "status": "True"} for t in
                                                                                 ("Initialized", "Ready", "ContainersReady", "PodScheduled")]}}


def main(n_items=10000, reps=5):
    body = json.dumps({"kind": "PodList", "apiVersion": "v1", "metadata": {"resourceVersion": "1"},
                       "items": [pod(i) for i in range(n_items)]}, separators=(",", ":")).encode()
    keep = (np.arange(n_items) % 2).astype(np.uint8)
    best = {"scan+filter": 1e9, "json round trip": 1e9}
    for _ in range(reps):
        t = time.perf_counter()
        items, ib, ie = _lib.list_scan(body)
        out = _lib.list_filter(body, items, keep, ib, ie)
        best["scan+filter"] = min(best["scan+filter"], time.perf_counter() - t)
        t = time.perf_counter()
        d = json.loads(body)
        d["items"] = [it for it, k in zip(d["items"], keep) if k]
        ref = json.dumps(d, separators=(",", ":")).encode()
        best["json round trip"] = min(best["json round trip"], time.perf_counter() - t)
    assert json.loads(out) == json.loads(ref)
    # ingress: scanned items -> interned checks (zg_list_resolve), against the per-item string path
    from spicedb_kubeapi_proxy_b200 import workloads
    e = zgpu.Engine(workloads.BOOTSTRAP_SCHEMA, host_only=True)
    for i in range(0, n_items, 2):
        e.intern("pod", f"ns-{i % 200}/pod-{i}")
    e.intern("user", "alice")
    tpl = e.list_template("pod", "view", "user", "alice")
    rels = [("pod", f"ns-{i % 200}/pod-{i}", "view", "user", "alice", "") for i in range(n_items)]
    best.update({"zg_list_resolve": 1e9, "strings + zg_resolve_checks": 1e9})
    for _ in range(reps):
        t = time.perf_counter()
        a, checked = e.list_resolve(body, items, tpl)
        best["zg_list_resolve"] = min(best["zg_list_resolve"], time.perf_counter() - t)
        t = time.perf_counter()
        b = e.resolve_checks(rels)
        best["strings + zg_resolve_checks"] = min(best["strings + zg_resolve_checks"], time.perf_counter() - t)
    assert checked.all() and a.tobytes() == b.tobytes()
    res = {"items": n_items, "body_mb": round(len(body) / 1e6, 2),
           **{k: {"ms": round(v * 1e3, 2), "MB_per_s": round(len(body) / 1e6 / v, 1)} for k, v in best.items()}}
    print(json.dumps(res))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
