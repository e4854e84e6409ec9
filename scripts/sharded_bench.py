"""Object-hash sharded store over real GPUs (one process per GPU, NCCL all-to-all between passes).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_bench.py [--workload cfg3] [--scale 0.1]

Every rank loads the same relationship stream and keeps the shard it owns; the batch is answered
by dist.ShardedStoreChecker and compared bit for bit with a replicated engine on the same GPU.
Prints one JSON line on rank 0."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import zgpu
from spicedb_kubeapi_proxy_b200 import dist as zdist, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--checks", type=int, default=200_000)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.by_name(a.workload, a.scale)
shard = zgpu.Engine(w.schema, device=local, shard_rank=rank, shard_count=world, subquery_capacity=1 << 24)
w.load_into(shard); shard.publish()
replica = zgpu.Engine(w.schema, device=local)
w.load_into(replica); replica.publish()
items = w.check_items(replica, zgpu.CHECK_DTYPE)[: a.checks]
want = replica.check_bulk(items)
ck = zdist.ShardedStoreChecker(shard, zdist.TorchTransport(), zgpu.CHECK_DTYPE)
got = ck.check_bulk(items)  # warm-up + parity
ok = bool(np.array_equal(got, want))
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    ck.check_bulk(items)
torch.cuda.synchronize(); dist.barrier()
dt = zdist.max_over_ranks(time.perf_counter() - t0, device="cuda")
t0 = time.perf_counter()
for _ in range(a.steps):
    replica.check_bulk(items)
rep = time.perf_counter() - t0
oks = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(oks, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"mode": "object-hash sharded store", "n_gpus": world, "workload": w.name, "tuples_total": w.n_tuples(),
                      "tuples_this_shard": shard.num_tuples(), "checks": int(items.size), "parity_with_replica": bool(oks.item()),
                      "sharded_Mchecks_s": items.size * a.steps / dt / 1e6, "replica_one_gpu_Mchecks_s": items.size * a.steps / rep / 1e6,
                      "levels": ck.stats["levels"], "subqueries_sent_rank0": ck.stats["subqueries_sent"] // (a.steps + 1),
                      "bytes_sent_rank0_per_step": ck.stats["bytes_sent"] // (a.steps + 1)}))
dist.destroy_process_group()
