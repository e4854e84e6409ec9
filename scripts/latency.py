"""Latency of the small-store operations the proxy actually issues (BASELINE config 1 shape):
single CheckPermission, 1-item bulk check, one-relationship WriteRelationships (rebuild +
publish), LookupResources. Run on the GPU box."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import zgpu
from spicedb_kubeapi_proxy_b200 import workloads

w = workloads.cfg1()
e = zgpu.Engine(w.schema)
w.load_into(e)
e.publish()
items = w.check_items(e, zgpu.CHECK_DTYPE)
e.check_bulk(items[:1])
def bench(fn, n):
    ts = []
    for i in range(n):
        t = time.perf_counter(); fn(i); ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e6
    return f"p50 {np.percentile(ts,50):8.1f} us  p99 {np.percentile(ts,99):8.1f} us  mean {ts.mean():8.1f} us"
print("1-item zg_check_bulk        ", bench(lambda i: e.check_bulk(items[i % 1000:i % 1000 + 1]), 2000))
print("1000-item zg_check_bulk     ", bench(lambda i: e.check_bulk(items), 500))
print("1-item zg_check_bulk_str    ", bench(lambda i: e.check_bulk_str([("namespace", "0", "view", "user", "1", "")]), 1000))
print("LookupResources (flat)      ", bench(lambda i: e.lookup_resources_ids("namespace", "view", "user", i % 10), 1000))
C = zgpu.client
cl = C.PermissionsClient(w.schema)
for i in range(1000):
    pass
ups = [(zgpu._lib.OP_TOUCH, f"namespace:ns{i}#viewer@user:u{i % 10}", 0) for i in range(1000)]
cl.engine.write_relationships(ups)
print("WriteRelationships 1 rel, 1k-rel store ", bench(lambda i: cl.engine.write_relationships([(zgpu._lib.OP_TOUCH, f"namespace:w{i}#creator@user:u{i % 7}", 0)]), 300))
w3 = workloads.cfg3(scale=0.1)
e3 = zgpu.Engine(w3.schema); w3.load_into(e3); e3.publish()
t = time.perf_counter(); e3.publish(); print("publish 1M-rel store: %.2f ms" % ((time.perf_counter() - t) * 1e3))
print(e.stats())
