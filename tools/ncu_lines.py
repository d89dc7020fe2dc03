"""Join an `ncu --page source --csv` export (SASS view) with `nvdisasm -g` line info: executed warp instructions and
stall samples per source line / per opcode of one kernel instance in the report.
    python tools/ncu_lines.py <source.csv> <nvdisasm -g -c output> <kernel index> [top]
Profiling tooling."""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, dis, kidx = sys.argv[1], sys.argv[2], int(sys.argv[3])
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: offset -> (file, line)
    loc, cur = {}, ("?", 0)
    fn_started = False
    for ln in open(dis):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and "sepconv_tc_kernel" in dis or m:
            off = int(m.group(1), 16)
            if off not in loc:
                loc[off] = cur
    rows = list(csv.reader(open(src_csv)))
    # split per kernel ("Kernel Name" rows)
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[kidx]
    e = starts[kidx + 1] if kidx + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    by_line, by_op, line_stall = defaultdict(lambda: [0, 0]), defaultdict(int), defaultdict(lambda: defaultdict(int))
    tot = 0
    for r in rows[s + 2:e]:
        if len(r) <= iex:
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        off = a - base
        ex, smp = int(r[iex] or 0), int(r[ismp] or 0)
        l = loc.get(off, ("?", 0))
        by_line[l][0] += ex
        by_line[l][1] += smp
        op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
        by_op[op.split(".")[0]] += ex
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                line_stall[l][hdr[i]] += v
        tot += ex
    print("kernel %d: %d warp instructions executed" % (kidx, tot))
    print("-- by source line (instr, %% of instr, samples, top stalls)")
    for l, (ex, smp) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(line_stall[l].items(), key=lambda kv: -kv[1])[:3]
        print("  %-22s:%-4d %12d %5.1f%%  smp %6d  %s" % (l[0], l[1], ex, 100.0 * ex / tot, smp, " ".join("%s=%d" % (k[6:], v) for k, v in st)))
    print("-- by opcode")
    for op, ex in sorted(by_op.items(), key=lambda kv: -kv[1])[:30]:
        print("  %-12s %12d %5.1f%%" % (op, ex, 100.0 * ex / tot))


if __name__ == "__main__":
    main()
