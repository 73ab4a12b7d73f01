#!/usr/bin/env python3
"""Attribute ncu warp-stall samples to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <libmelvin_b200.so> <kernel-regex> [top]
Joins the SASS page of the report (per-instruction samples) with the line table of
the cubin embedded in the library (nvdisasm --print-line-info), by instruction order.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_samples(rep, regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          f"regex:{regex}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    isrc, ismp = hdr.index("Source"), hdr.index("# Samples")
    data = [(r[isrc].strip(), int(r[ismp])) for r in rows[2:] if len(r) > ismp and r[ismp].isdigit()]
    return name, data


def line_table(so, mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    secs = re.split(r"\n\s*\.section\s+\.text\.", txt)
    for sec in secs[1:]:
        head = sec.split("\n", 1)[0]
        if all(h in head for h in mangled_hint):
            cur = None
            table = []
            for ln in sec.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
                if m:
                    table.append((cur, m.group(1).strip()))
            return table
    raise SystemExit("kernel section not found")


def main():
    rep, so, regex = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    name, data = sass_samples(rep, regex)
    hint = re.findall(r"k_[a-z0-9_]+", name)[:1] + [f"Li{n}E" for n in re.findall(r"\(int\)(\d+)", name)]
    table = line_table(so, hint)
    if len(data) == 2 * len(table):      # ncu lists the function twice (identical halves)
        data = data[:len(table)]
    if len(table) != len(data):
        print(f"warning: {len(data)} profiled instructions vs {len(table)} in the cubin", file=sys.stderr)
    tot = sum(s for _, s in data) or 1
    per_line = {}
    for (line, _), (_, smp) in zip(table, data):
        per_line[line] = per_line.get(line, 0) + smp
    src_cache = {}
    print(f"{name}: {tot} samples")
    for line, smp in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if line:
            fn = os.path.join(os.path.dirname(os.path.abspath(so)), "..", "..", "csrc", line[0])
            if os.path.exists(fn):
                src_cache.setdefault(fn, open(fn).read().splitlines())
                if line[1] - 1 < len(src_cache[fn]):
                    text = src_cache[fn][line[1] - 1].strip()[:90]
        print(f"{100.0 * smp / tot:5.1f}%  {str(line):34s} {text}")


if __name__ == "__main__":
    main()
