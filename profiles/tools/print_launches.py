"""Print kernel name + duration (ns) from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys

for r in csv.reader(open(sys.argv[1])):
    if len(r) > 14 and r[0].isdigit():
        print(f"{r[4][:60]:60s} {r[-1]}")
