"""Print slices of an ncu launch list (gpu__time_duration.sum CSV): the decoder tail and the encoder of the first block, plus the total."""
import csv
import sys


def load(path):
    rows = []
    lines = [l for l in open(path) if not l.startswith('==')]
    for x in csv.DictReader(lines):
        if x.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((x['Kernel Name'], x['Grid Size'], float(x['Metric Value'].replace(',', '')) / 1e3))
    return rows


if __name__ == "__main__":
    rows = load(sys.argv[1])
    print(len(rows), "launches, total %.3f ms" % (sum(r[2] for r in rows) / 1e3))
    i = [i for i, x in enumerate(rows) if 'head' in x[0]][0]
    for x in rows[i - 5:i + 1]:
        print("%-72s %-14s %8.1f" % (x[0][:70], x[1], x[2]))
    i = [i for i, x in enumerate(rows) if 'stem' in x[0]][0]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for x in rows[i:i + n]:
        print("%-72s %-14s %8.1f" % (x[0][:70], x[1], x[2]))
