"""Per-source-line stall samples / executed instructions of one kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_lines.py REPORT.ncu-rep DISASSEMBLY.dis MANGLED_KERNEL_NAME [SOURCE.cu]

DISASSEMBLY.dis = `nvdisasm -c -gi` of the cubin the report was taken from (cuobjdump -xelf <name> lib.so): the SASS page of
ncu carries no line numbers in CSV form, so lines are matched by instruction ORDER within the kernel."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, dis, kern = sys.argv[1:4]
    src = open(sys.argv[4]).read().split("\n") if len(sys.argv) > 4 else None
    text = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(text) if l.startswith(".text." + kern + ":"))
    seq, cur = [], None
    for l in text[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            seq.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    iss, ie, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
    if len(seq) != len(data):
        print("warning: %d instructions in the disassembly, %d in the report" % (len(seq), len(data)))
    agg = collections.defaultdict(lambda: [0, 0])
    for i in range(min(len(seq), len(data))):
        agg[seq[i]][0] += int(data[i][iss] or 0)
        agg[seq[i]][1] += int(data[i][ie] or 0)
    tot = sum(v[0] for v in agg.values()) or 1
    tote = sum(v[1] for v in agg.values()) or 1
    print("kernel %s: %d stall samples, %d warp instructions" % (kern, tot, tote))
    for k, v in sorted(agg.items(), key=lambda kv: kv[0] or ("", 0)):
        if v[0] > tot * 0.005 or v[1] > tote * 0.008:
            t = src[k[1] - 1].strip()[:100] if (src and k and k[0].endswith(".cu")) else ""
            print("%-22s %5.1f%% samples %5.1f%% inst  %s" % ("%s:%d" % k if k else "?", 100 * v[0] / tot, 100 * v[1] / tote, t))
    top = sorted(range(min(len(seq), len(data))), key=lambda i: -int(data[i][iss] or 0))[:14]
    print("hottest instructions:")
    for i in sorted(top):
        print("  %5d %-20s %5.1f%%  %s" % (i, "%s:%d" % seq[i] if seq[i] else "?", 100 * int(data[i][iss] or 0) / tot, data[i][isrc].strip()[:80]))


if __name__ == "__main__":
    main()
