#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) into the text kept under profiles/:
per kernel the duration, DRAM traffic, pipe utilisation, stall mix, opcode mix and the hottest SASS lines.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "sm__cycles_elapsed.max"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    print(f"# summary of {rep}\n")
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        print("## kernel:", d.get("Kernel Name", "")[:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
                         and d[k] not in ("", "n/a")), reverse=True)[:8]
        print("  stall reasons (warps stalled per issue-active cycle):")
        for v, k in stalls:
            print(f"    {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
        print()
    src = ncu(rep, "source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]):
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen or not b["rows"]:
            continue
        seen.add(b["name"])
        h = b["hdr"]
        iS, iSamp, iExec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        ops, samp, lines = collections.Counter(), collections.Counter(), []
        for r in b["rows"]:
            s = r[iS].strip()
            t = s.split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += int(r[iExec]); samp[op] += int(r[iSamp]); lines.append((int(r[iSamp]), int(r[iExec]), s))
        tot, ts = sum(ops.values()) or 1, sum(samp.values()) or 1
        print("## SASS of", b["name"][:100])
        print(f"  static instructions {len(lines)}, executed warp-instructions {tot}, samples {ts}")
        print("  opcode mix (share of executed / share of stall samples):")
        for op, c in ops.most_common(14):
            print(f"    {op:10s} {100 * c / tot:5.1f}%  {100 * samp[op] / ts:5.1f}%")
        print("  hottest lines (samples, executed, SASS):")
        for sm, ex, s in sorted(lines, reverse=True)[:12]:
            print(f"    {100 * sm / ts:5.2f}%  {ex:10d}  {s[:90]}")
        print()


if __name__ == "__main__":
    main()
