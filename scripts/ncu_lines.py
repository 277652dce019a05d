"""Per-CUDA-source-line summary of an `ncu --set full --import-source on` capture (read on the CPU box):
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top]
prints the share of executed warp instructions and of stall samples by file and by source line."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, out = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and cur and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            try:
                ie, ns = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            stalls = {}
            for i, h in enumerate(hdr):
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        stalls[h[6:]] = int(r[i])
                    except ValueError:
                        pass
            out.append((cur, int(r[0]), r[1].strip()[:100], ie, ns, stalls))
    tot = float(sum(o[3] for o in out)) or 1.0
    tots = float(sum(o[4] for o in out)) or 1.0
    print("total warp instructions %d, stall samples %d" % (tot, tots))
    bf, bs = collections.Counter(), collections.Counter()
    allst = collections.Counter()
    for o in out:
        bf[o[0]] += o[3]
        bs[o[0]] += o[4]
        allst.update(o[5])
    for f in bf:
        print("%-22s inst %.3f  samples %.3f" % (f, bf[f] / tot, bs[f] / tots))
    print("stall mix:", "  ".join("%s %.3f" % (k, v / tots) for k, v in allst.most_common(10)))
    print("--- top lines by samples")
    for o in sorted(out, key=lambda x: -x[4])[:top]:
        st = " ".join("%s %.0f%%" % (k, 100.0 * v / max(o[4], 1)) for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3] if v)
        print("%-18s %4d inst %.4f samp %.4f [%s] | %s" % (o[0], o[1], o[3] / tot, o[4] / tots, st, o[2]))


if __name__ == "__main__":
    main()
