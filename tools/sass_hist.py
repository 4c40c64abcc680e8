"""Histogram the SASS opcodes of a cubin between two instruction addresses.

usage: python tools/sass_hist.py file.cubin [lo_hex hi_hex]   (addresses as printed by cuobjdump)
Used to count the issue slots per element of a generated kernel before spending GPU time.
"""
import collections
import re
import subprocess
import sys


def main():
    cubin = sys.argv[1]
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    hist = collections.Counter()
    full = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if lo <= addr < hi:
            op = m.group(2)
            hist[op.split(".")[0]] += 1
            full[op] += 1
    tot = sum(hist.values())
    for k, v in hist.most_common():
        print(f"{v:5d} {k}")
    print(f"{tot:5d} TOTAL")
    if "-v" in sys.argv:
        for k, v in full.most_common():
            print(f"   {v:5d} {k}")


if __name__ == "__main__":
    main()
