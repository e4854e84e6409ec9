"""Turns the ncu captures of one GPU call (gpurun_out/<tag>_prof_{cfg3,cfg4}.ncu-rep) into the committed evidence:
profiles/<tag>_check_kernel.md (metrics table + hottest source lines), profiles/<tag>_sass_opcodes.txt (static opcode
histogram of the shipped check_kernel from cuobjdump), and the entries of profiles/ncu_traffic.json that bench.py reads
for roofline.traffic / dram_frac.      python scripts/profile_summary.py <tag>"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "launch__registers_per_thread", "launch__grid_size"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ms(v, unit):
    return float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)


out = [f"# {tag}: `zg::check_kernel<false,false>`, one launch each, `ncu --set full --clock-control none`\n",
       "Captured by `scripts/gpu_perf.sh` (`python bench.py --workload <wl> --configs '' --steps 2 --warmup 3`), read with",
       "`scripts/profile_summary.py`. ncu times are cold-cache and serialised: compare shares, not absolutes.\n"]
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
table = {}
for wl in ("cfg3", "cfg4"):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{wl}.ncu-rep")
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    table[wl] = m
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    traffic[wl] = {"dram_bytes": int(dram), "kernel_ms": to_ms(*m["gpu__time_duration.sum"]),
                   "source": f"profiles/{tag}_check_kernel.md"}
out.append("| metric | " + " | ".join(table) + " |")
out.append("|---|" + "---|" * len(table))
for k in WANT:
    out.append(f"| {k} | " + " | ".join(f"{table[w][k][0]} {table[w][k][1]}" if k in table[w] else "-" for w in table) + " |")
for wl in table:
    t = traffic[wl]
    inst = float(table[wl]["smsp__inst_executed.sum"][0])
    out.append(f"\n{wl}: DRAM traffic {t['dram_bytes'] / 1e6:.1f} MB per launch = {t['dram_bytes'] / (t['kernel_ms'] / 1e3) / 1e9:.0f} GB/s "
               f"under ncu; {inst / 1048576:.0f} warp-instructions per check (batch 1 048 576).")

# hottest source lines (needs the cubin of the same build for the line table)
so = os.environ.get("ZG_PROF_LIB") or os.path.join(ROOT, "spicedb-kubeapi-proxy_b200", "libzgpu.so")  # the build that was profiled
tmp = f"/tmp/zg_prof_{tag}"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, "device.sm_100a.cubin")], capture_output=True, text=True).stdout
sec = sass.split(".text._ZN2zg12check_kernelILb0ELb0EEEvNS_7KParamsE:")[1].split("\n.text.")[0] if "check_kernelILb0ELb0" in sass else ""
off2line, line = {}, None
ops = collections.Counter()
for l in sec.split("\n"):
    mm = re.search(r'//## File ".*kernels.cuh", line (\d+)', l)
    if mm:
        line = int(mm.group(1))
        continue
    mm = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if mm:
        off2line[int(mm.group(1), 16)] = line
        ops[mm.group(2).split(".")[0]] += 1
src = open(os.environ.get("ZG_PROF_SRC") or os.path.join(ROOT, "spicedb-kubeapi-proxy_b200", "csrc", "kernels.cuh")).read().split("\n")
for wl in table:
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{wl}.ncu-rep")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))[1:]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(rows[1][ix["Address"]], 16)
    inst, stall, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[1:]:
        ln = off2line.get(int(r[ix["Address"]], 16) - base)
        n = int(r[ix["Instructions Executed"]])
        inst[ln] += n
        thr[ln] += int(r[ix["Thread Instructions Executed"]])
        stall[ln] += int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    tot, ts = sum(inst.values()) or 1, sum(stall.values()) or 1
    out.append(f"\n## {wl}: hottest source lines (share of warp instructions, of stall samples, active lanes per instruction)\n")
    out.append("| inst | stalls | lanes | kernels.cuh |")
    out.append("|---|---|---|---|")
    for ln, n in inst.most_common(16):
        text = src[ln - 1].strip()[:110].replace("|", "\\|") if ln else "?"
        out.append(f"| {n / tot * 100:.1f} % | {stall[ln] / ts * 100:.1f} % | {thr[ln] / max(n, 1):.1f} | L{ln}: `{text}` |")
open(os.path.join(ROOT, "profiles", f"{tag}_check_kernel.md"), "w").write("\n".join(out) + "\n")
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.txt"), "w") as f:
    f.write(f"static opcode histogram of zg::check_kernel<false,false> in the shipped libzgpu.so ({sum(ops.values())} instructions; "
            "cuobjdump -xelf + nvdisasm -g). Integer / control code: no tensor (UTC*MMA, HMMA) and no TMA (UTMALDG, UBLKCP) opcodes "
            "-- by design, this is a pointer-chasing traversal.\n")
    for op, n in ops.most_common():
        f.write(f"{op:14s} {n}\n")
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("\n".join(out[:40]))
