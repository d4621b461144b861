"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per source line and
per region of nn2.cu (developer tool; the summaries land in profiles/)."""
import bisect, json, sys

path, src_path = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
cur, hdr_len, agg = None, None, {}
for line in open(path, errors="replace"):
    line = line.rstrip("\n")
    if line.startswith('"File Path"'):
        cur = line.split('","')[1].rstrip('"').split("/")[-1]
        continue
    if line.startswith('"Line No"'):
        hdr_len = len(line.split('","'))
        continue
    if hdr_len is None or line.startswith('"Function Name"') or line.startswith('"",'):
        continue
    f = line.strip('"').split('","')
    if not f[0].isdigit() or len(f) < hdr_len:
        continue
    ln = int(f[0])
    samples, inst = int(f[6 - hdr_len]), int(f[7 - hdr_len])
    text = '","'.join(f[1:len(f) - hdr_len + 2])
    a = agg.setdefault((cur, ln), [0, 0, text])
    a[0] += samples
    a[1] += inst
tot = sum(a[0] for a in agg.values())
toti = sum(a[1] for a in agg.values())
src = open(src_path).read().splitlines()
base = src_path.split("/")[-1]


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    return None


marks = [(1, "header"), (find("inline bool lane_rules_out"), "lane_rules_out"),
         (find("float filter_threshold"), "filter_threshold"), (find("void scan_rows("), "scan (filter FFMA2 loop)"),
         (find("void scan_subtile("), "scan: flags"), (find("// ---- resolve: rare"), "resolve: setup + threshold"),
         (find("u64 wmask = 0;"), "resolve: pass 1 (window mask)"), (find("// pass 2: exact FP64"), "resolve: pass 2 (FP64)"),
         (find("Dbest_l[r] = Db;"), "resolve: write-back + bound"), (find("// ---- exhaustive kernel"), "exhaustive"),
         (find("nn2_pruned_kernel(const NN2Params p, const __grid_constant__ IcpFuse f)"), "set-up (query load, state init)"),
         (find("// ---- correspondences -> per-row sums"), "fused ICP epilogue (gather, sums)"),
         (find("report in the caller's original indexing") or find("const int io = p.perm_q != nullptr ? p.perm_q[i] : i;"), "result write-back"),
         (find("// ---- starting bounds from"), "hints"), (find("// ---- query-row spheres"), "query-row spheres"),
         (find("auto refresh_bounds"), "refresh row bounds"), (find("auto coarse_rows ="), "coarse test (stage spheres)"),
         (find("auto coarse_rows_box"), "coarse test (sub-tile box)"), (find("auto exact_any ="), "exact test (stage)"),
         (find("auto exact_rows_box"), "exact test (sub-tile box)"), (find("// ---- seeds:"), "seeds"),
         (find("// ---- main loop"), "main loop control"), (find("if (produced_all && !sorted)"), "nearest-first sort"),
         (find("if (pending > 0 && (seeding"), "consume: tests, bulk copies, wait"),
         (find("seeding = false;"), "produce: stages -> FIFO"), (find("atomicAdd(p.evaluated + 0, (unsigned long long)nhalves"), "epilogue")]
marks = sorted(m for m in marks if m[0])
reg = {}
for (f, ln), (s, i, t) in agg.items():
    name = f if f != base else marks[bisect.bisect_right([m[0] for m in marks], ln) - 1][1]
    # the inline-PTX helpers of isr_common.cuh are inlined into nn2.cu's phases: the packed FFMA2 /
    # FMNMX3 (and their pack / unpack moves) are the filter scan itself (the resolve's pass 1 uses
    # scalar __fmaf_rn), the mbarrier / bulk-copy helpers belong to the consume phase
    if f == "isr_common.cuh":
        if any(k in t for k in ("fma.rn.f32x2", "min.f32", "mov.b64")):
            name = "scan (filter FFMA2 loop)"
        elif any(k in t for k in ("mbarrier", "cp.async.bulk", "smem_u32", "try_wait")):
            name = "consume: tests, bulk copies, wait"
    a = reg.setdefault(name, [0, 0])
    a[0] += s
    a[1] += i
print("total samples", tot, "instructions", toti)
regions = []
for n, (s, i) in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:36s} samples {100 * s / tot:5.1f}%  inst {100 * i / toti:5.1f}%")
    regions.append({"region": n, "samples_pct": round(100 * s / tot, 2), "instructions_pct": round(100 * i / toti, 2)})
print()
top = []
for (f, ln), (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{f}:{ln:5d} {100 * s / tot:5.1f}% inst {100 * i / toti:5.1f}%  {t.strip()[:100]}")
    top.append({"file": f, "line": ln, "samples_pct": round(100 * s / tot, 2), "instructions_pct": round(100 * i / toti, 2),
                "source": t.strip()[:120]})
if out_json:
    json.dump({"source_export": path, "total_samples": tot, "total_instructions": toti, "regions": regions,
               "top_lines": top}, open(out_json, "w"), indent=1)
