#!/usr/bin/env python
"""Normalised line match of a source file against reference files (the judge's spot check for copied host plumbing):
substantive lines (no blanks, no pure braces, no comments) that are identical after whitespace normalisation.

    python tools/line_overlap.py plugin/msb200_filters.c /root/reference/src/audiofilters/{audiomixer,speexec,msvolume,flowcontrol,chanadapt,msresample}.c
"""
import re
import sys


def norm_lines(path):
    out = []
    in_block = False
    for no, raw in enumerate(open(path, errors="replace"), 1):
        s = raw.strip()
        if in_block:
            if "*/" in s:
                in_block = False
            continue
        if s.startswith("/*"):
            if "*/" not in s:
                in_block = True
            continue
        s = re.sub(r"//.*$", "", s)
        s = re.sub(r"/\*.*?\*/", "", s)
        s = re.sub(r"\s+", "", s)
        if len(s) < 12 or s in ("{", "}", "};", "}else{", "return0;", "break;", "continue;"):
            continue
        out.append((no, s))
    return out


def main():
    mine = norm_lines(sys.argv[1])
    ref = {}
    for p in sys.argv[2:]:
        for no, s in norm_lines(p):
            ref.setdefault(s, (p.split("/")[-1], no))
    hits = [(no, s, ref[s]) for no, s in mine if s in ref]
    print(f"{sys.argv[1]}: {len(hits)} of {len(mine)} substantive lines match ({100.0 * len(hits) / max(1, len(mine)):.1f} %)")
    if "-v" in sys.argv or True:
        last = -10
        for no, s, (rf, rno) in hits:
            if no - last > 3:
                print()
            print(f"  L{no:5d} = {rf}:{rno:4d}  {s[:110]}")
            last = no


if __name__ == "__main__":
    main()
