"""Developer tool: SASS instruction count of nn2_pruned_kernel per source region
(nvdisasm -g of the nn2 cubin extracted with cuobjdump -xelf)."""
import bisect, re, sys
lines = open(sys.argv[1]).read().splitlines()
src = open(sys.argv[2]).read().splitlines()
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." in l and "nn2_pruned_kernel" in l)
end = next(i for i in range(start + 1, len(lines)) if ".section" in lines[i])
cur, cnt = None, {}
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        cnt[cur] = cnt.get(cur, 0) + 1
tot = sum(cnt.values())
def find(s_):
    return next((i + 1 for i, l in enumerate(src) if s_ in l), None)
marks = [(1, "hdr"), (find("void scan_rows("), "scan_rows"), (find("bool resolve_flagged("), "resolve"),
         (find("void scan_subtile("), "scan_subtile(full)"), (find("void scan_subtile_pruned("), "scan_pruned wrapper"),
         (find("// ---- exhaustive kernel"), "exh"), (find("nn2_pruned_kernel(const NN2Params p)"), "setup"),
         (find("// ---- starting bounds from"), "hints"), (find("// ---- query-row spheres"), "rowspheres"),
         (find("auto refresh_bounds"), "refresh"), (find("auto coarse_rows ="), "coarse"),
         (find("auto coarse_rows_box"), "coarse_box"), (find("auto exact_any ="), "exact_any"),
         (find("auto exact_rows_box"), "exact_box"), (find("// ---- seeds:"), "seeds"), (find("// ---- main loop"), "mainctl"),
         (find("if (produced_all && !sorted)"), "sort"), (find("if (pending > 0 && (seeding"), "consume"),
         (find("seeding = false;"), "produce"), (find("atomicAdd(p.evaluated + 0, (unsigned long long)nhalves"), "epilogue")]
marks = sorted(m for m in marks if m[0])
reg = {}
for k, c in cnt.items():
    name = "?" if k is None else (k[0] if k[0] != "nn2.cu" else marks[bisect.bisect_right([m[0] for m in marks], k[1]) - 1][1])
    reg[name] = reg.get(name, 0) + c
print("instructions", tot)
for n, c in sorted(reg.items(), key=lambda kv: -kv[1]):
    print(f"{n:24s} {c:5d} {100 * c / tot:5.1f}%")
