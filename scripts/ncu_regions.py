"""Where a kernel's warp instructions are: the per-SASS-instruction execution counts of an `ncu --set full --import-source on`
report (source page), grouped into runs of consecutive instructions that execute equally often -- basic blocks that always run
together -- with each run's share of the kernel's warp instructions, its average active lanes and its first / last instruction.
    python scripts/ncu_regions.py gpurun_out/r2c20_prof.ncu-rep shade_kernel profiles/r2_final_shade_kernel_regions.json"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, kernel, out = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else kernel
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    inst = []
    for r in rows[2:]:
        if len(r) < len(hdr) - 5 or not r[idx["Instructions Executed"]].isdigit():
            break                                                    # (the page repeats the kernel in a second view)
        inst.append((r[idx["Source"]].strip(), int(r[idx["Instructions Executed"]]), int(r[idx["Thread Instructions Executed"]])))
    total = sum(i[1] for i in inst)
    groups, cur = [], None
    for n, (src, ie, te) in enumerate(inst):
        if cur and abs(ie - cur["executions"]) <= 0.03 * max(ie, cur["executions"], 1):
            cur["last"] = n; cur["instructions"] += 1; cur["warp_instructions"] += ie; cur["_te"] += te; cur["last_op"] = src
        else:
            cur = {"first": n, "last": n, "instructions": 1, "executions": ie, "warp_instructions": ie, "_te": te, "first_op": src, "last_op": src}
            groups.append(cur)
    regions = []
    for g in groups:
        if g["warp_instructions"] < 0.003 * total:
            continue
        g["share_pct"] = round(100.0 * g["warp_instructions"] / total, 2)
        g["avg_active_lanes"] = round(g.pop("_te") / max(1, g["warp_instructions"]), 1)
        regions.append(g)
    res = {"report": rep, "kernel": name, "sass_instructions": len(inst), "warp_instructions": total,
           "note": "runs of consecutive SASS instructions with (within 3 %) equal execution counts; runs below 0.3 % of the kernel omitted",
           "regions": regions}
    json.dump(res, open(out, "w"), indent=1)
    for g in regions:
        print("%4d-%4d n=%3d x %9d = %6.1f M  %5.1f %%  lanes %.1f  %s" % (g["first"], g["last"], g["instructions"], g["executions"],
                                                                         g["warp_instructions"] / 1e6, g["share_pct"], g["avg_active_lanes"], g["first_op"][:40]))
    print("total %.1f M warp instructions, %d SASS instructions" % (total / 1e6, len(inst)))


if __name__ == "__main__":
    main()
