"""Top stall lines of one kernel from an ncu report's source page (SASS view).

    python tools/ncu_top.py <report.ncu-rep> <kernel regex> [launch index] [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    print(lines[0])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    total = 0
    for r in rows[1:]:
        try:
            n = int(r[ix["# Samples"]])
        except (ValueError, IndexError):
            continue
        total += n
        data.append((n, r))
    print(f"total samples {total}")
    agg = {}
    for n, r in data:
        for h in stall_cols:
            try:
                agg[h] = agg.get(h, 0) + int(r[ix[h]])
            except ValueError:
                pass
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for pos, (n, r) in enumerate(data):
        r.append(pos)
    for n, r in sorted(data, key=lambda x: -x[0])[:top]:
        st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols if r[ix[h]].isdigit()), reverse=True)[:2]
        print(f"{n:7d} {100*n/total:5.1f}%  #{r[-1]:5d} {r[ix['Source']].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main()
