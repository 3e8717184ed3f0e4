#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions, stall samples and shared-memory wavefronts of one kernel launch of an
.ncu-rep (ncu --set full --import-source on), joined with the line table of the object file the kernel was built from
(nvdisasm -g).  The ncu CLI prints the SASS view only; this gives the source view without the GUI.
Usage: python tools/ncu_lines.py rep.ncu-rep csrc/rowown.o 'k_rowown<3, 2, 1' [launch_index] [top_n]"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def sass_lines(obj, mangled_hint):
    """offset -> (file, line) of the function whose mangled name contains every token of mangled_hint"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, cur, inside = {}, ("?", 0), False
    for l in txt:
        if l.startswith(".text."):
            inside = mangled_hint in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    mangled = sys.argv[6] if len(sys.argv) > 6 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.escape(kname.split("<")[0]), "--launch-skip", str(launch),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print("#", rows[0][1][:140])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix["Address"]] in seen or not re.fullmatch(r"(0x)?[0-9a-fA-F]+", r[ix["Address"]]):
            continue
        seen.add(r[ix["Address"]])
        data.append(r)
    base = min(int(r[ix["Address"]], 16) for r in data)
    if mangled is None:   # _ZN5nsgpu8k_rowownILi3ELi2ELb1ELb1ELb1ELb1EEE... from 'k_rowown<3, 2, 1, 1, 1, 1>'
        name, args = kname.split("<")[0], re.findall(r"\d+", kname.split("<", 1)[1]) if "<" in kname else []
        full = re.findall(r"\(int\)(\d+)|\(bool\)(\d+)", rows[0][1])
        mangled = name + "I" + "".join(("Li%sE" % a) if a else ("Lb%sE" % b) for a, b in full)
    table = sass_lines(obj, mangled)
    if not table:
        sys.exit("no function matching " + mangled)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        (fl, _s) = table.get(off, (("?", 0), ""))
        a = agg[fl]
        a["inst"] += f(r, "Instructions Executed")
        a["samples"] += f(r, "# Samples")
        for k in ("stall_short_sb", "stall_long_sb", "stall_wait", "stall_branch_resolving", "stall_mio", "stall_math", "stall_barrier", "stall_selected", "stall_not_selected"):
            a[k] += f(r, k)
        a["wave"] += f(r, "L1 Wavefronts Shared")
        a["wave_x"] += f(r, "L1 Wavefronts Shared Excessive")
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    print(f"total: {tot['inst']:.0f} warp-instructions, {tot['samples']:.0f} samples, shared wavefronts {tot['wave']:.0f} (excessive {tot['wave_x']:.0f})")
    print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s} {'ssb':>6s} {'lsb':>6s} {'wait':>6s} {'bar':>6s} {'issue':>6s} {'wave%':>6s} {'excess%':>7s}")
    for fl, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print(f"{fl[0] + ':' + str(fl[1]):28s} {100 * a['inst'] / tot['inst']:6.2f} {100 * a['samples'] / tot['samples']:6.2f} {a['stall_short_sb']:6.0f} {a['stall_long_sb']:6.0f} "
              f"{a['stall_wait']:6.0f} {a['stall_barrier']:6.0f} {a['stall_selected']:6.0f} {100 * a['wave'] / max(tot['wave'], 1):6.2f} {100 * a['wave_x'] / max(tot['wave_x'], 1):7.2f}")
    # by file
    byfile = collections.defaultdict(lambda: collections.Counter())
    for fl, a in agg.items():
        byfile[fl[0]].update(a)
    print("by file:")
    for fn, a in sorted(byfile.items(), key=lambda kv: -kv[1]["samples"]):
        print(f"  {fn:26s} inst {100 * a['inst'] / tot['inst']:6.2f}%  samples {100 * a['samples'] / tot['samples']:6.2f}%")


if __name__ == "__main__":
    main()
