"""Join an `ncu --page source --csv` dump (SASS rows with stall samples) with nvdisasm's line info of the same kernel and print
the share of samples per CUDA source line.   usage: ncu_by_line.py <source.csv> <cubin> <kernel-substring> [min_pct]"""
import csv
import re
import subprocess
import sys


def sass_lines(cubin, kern):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, on, line = [], False, None
    for l in txt:
        if l.startswith("\t.section") or l.startswith(".section"):
            on = kern in l and ".text." in l
            continue
        if not on:
            continue
        m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), line))
    return out


def main():
    src, cubin, kern = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows if r and re.fullmatch(r"0x[0-9a-f]+|[0-9]+", r[0] or "") and len(r) > si]
    sl = sass_lines(cubin, kern)
    if len(sl) != len(data):
        print(f"warning: {len(data)} profiled instructions vs {len(sl)} disassembled", file=sys.stderr)
    agg, tot, ins_tot = {}, 0.0, 0.0
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for k, r in enumerate(data):
        s = float(r[si] or 0)
        tot += s
        ins_tot += float(r[ii] or 0)
        key = sl[k][2] if k < len(sl) else None
        a = agg.setdefault(key, [0.0, 0.0, {}])
        a[0] += s
        a[1] += float(r[ii] or 0)
        for c in stall_cols:
            v = float(r[c] or 0)
            if v:
                a[2][hdr[c]] = a[2].get(hdr[c], 0.0) + v
    print(f"total samples {tot:.0f}, warp instructions {ins_tot:.3g}")
    for key, (s, n, st) in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if 100 * s / tot >= min_pct:
            top = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"{100 * s / tot:5.1f}%  inst {100 * n / ins_tot:5.1f}%  {key}   [{top}]")


if __name__ == "__main__":
    main()
