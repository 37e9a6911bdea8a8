"""Runs compute-sanitizer racecheck on scripts/sanitize_small.py and prints the hazards grouped by
(kind, writer source line, reader source line): which synchronisation each pair relies on is
annotated by hand in profiles/r02_racecheck_summary.txt."""
import collections, re, subprocess, sys
p = subprocess.run(["compute-sanitizer", "--tool", "racecheck", "--racecheck-report", "all", "--print-limit", "400000",
                    sys.executable, "scripts/sanitize_small.py"], capture_output=True, text=True)
txt = p.stdout + p.stderr
groups = collections.Counter()
kind = w = None
for line in txt.splitlines():
    m = re.search(r"Potential (\w+) hazard detected at __shared__", line)
    if m:
        kind, w = m.group(1), None
        continue
    m = re.search(r"(Write|Read) Thread .* in (\w+\.cuh?):(\d+)", line)
    if m and kind:
        if w is None:
            w = (m.group(1), m.group(2), m.group(3))
        else:
            groups[(kind, w, (m.group(1), m.group(2), m.group(3)))] += 1
            kind = None
for line in txt.splitlines():
    if "RACECHECK SUMMARY" in line or "ERROR SUMMARY" in line:
        print(line)
for (k, a, b), n in sorted(groups.items(), key=lambda kv: -kv[1]):
    print(f"{n:8d}  {k}  {a[0]} {a[1]}:{a[2]}  ->  {b[0]} {b[1]}:{b[2]}")
