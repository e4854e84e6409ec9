"""Debug helper (GPU box): find GPU/oracle mismatches on a workload and classify them."""
import sys, numpy as np
sys.path.insert(0, '.')
import zgpu
from oracle.pyoracle import Oracle
from spicedb_kubeapi_proxy_b200 import workloads
name, scale = sys.argv[1], float(sys.argv[2])
w = workloads.by_name(name, scale)
e, o = zgpu.Engine(w.schema), Oracle(w.schema)
w.load_into(e); w.load_into(o); e.publish()
items = w.check_items(e, zgpu.CHECK_DTYPE)
got, want = e.check_bulk(items), o.check_bulk(items)
bad = np.nonzero(got != want)[0]
print("mismatches", bad.size, "of", items.size, "slots", e.slot_table())
for i in bad[:12]:
    single = e.check_bulk(items[i:i+1])[0]
    blk = e.check_bulk(items[(i//32)*32:(i//32)*32+32])[i % 32]
    print(i, items[i], "got", got[i], "want", want[i], "single", single, "block32", blk)
print("batch positions mod 32:", np.bincount(bad % 32, minlength=32))
print("got values", np.unique(got[bad], return_counts=True), "want", np.unique(want[bad], return_counts=True))
