"""Top source lines of a kernel in an .ncu-rep (needs -lineinfo at compile time and --import-source on at capture).
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys

def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    data = []
    fname = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
            d = {"file": fname, "line": int(r[0]), "src": r[1].strip(), "inst": int(r[ii]), "smp": int(r[si])}
            for name in ("stall_long_sb", "stall_short_sb", "stall_lg", "stall_mio", "stall_wait", "stall_math", "stall_barrier", "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected"):
                d[name] = int(r[hdr.index(name)])
            data.append(d)
        except (ValueError, IndexError):
            pass
    ti = sum(d["inst"] for d in data) or 1
    ts = sum(d["smp"] for d in data) or 1
    print("total warp-instructions %d, samples %d" % (ti, ts))
    agg = {}
    for name in ("stall_long_sb", "stall_short_sb", "stall_lg", "stall_mio", "stall_wait", "stall_math", "stall_barrier", "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_selected"):
        agg[name] = sum(d[name] for d in data)
    print("stall samples:", {k: "%.1f%%" % (100.0 * v / ts) for k, v in sorted(agg.items(), key=lambda x: -x[1])})
    for d in sorted(data, key=lambda x: -x["smp"])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * d["smp"] / ts, 100.0 * d["inst"] / ti, d["file"], d["line"], d["src"][:100]))

if __name__ == "__main__":
    main()
