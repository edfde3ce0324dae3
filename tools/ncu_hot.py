"""Top stall sites of one kernel of an ncu report, SASS level with the stall reason that dominates each site:
   python tools/ncu_hot.py report.ncu-rep [launch index] [top n]"""
import csv
import subprocess
import sys


def main():
    rep, k, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout.splitlines()
    print(out[0][:160])
    rows = list(csv.reader(out[1:]))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[1:]:
        try:
            n = int(r[ix["# Samples"]])
        except (ValueError, IndexError):
            continue
        data.append((n, r))
    tot = sum(n for n, _ in data)
    print("total samples", tot)
    for pos, (n, r) in enumerate(data):
        r.append(pos)
    for n, r in sorted(data, key=lambda t: -t[0])[:top]:
        why = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
        print(f"{100.0 * n / tot:5.1f}%  #{r[-1]:5d} {r[ix['Source']][:70]:70s} {why[0][1]}={why[0][0]} {why[1][1]}={why[1][0]} exec={r[ix['Instructions Executed']]}")


if __name__ == "__main__":
    main()
