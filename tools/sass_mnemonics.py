"""SASS mnemonic counts per kernel of the built library (and of the L2 microbenchmark): which memory / tensor
instructions each kernel really contains.    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ("LDG.E.128", "LDG.E.64", "REDG.E.ADD.F32x4", "RED.E.ADD.F32", "ATOMG", "HMMA", "LDGSTS", "LDCU", "LDS.128", "STS.128",
         "SHFL", "MUFU", "UTMALDG", "UBLKRED", "UTMAREDG", "UTCHMMA", "LDTM", "SYS", "BAR.SYNC")


def scan(path, only=None):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("eslam::", "").replace("void ", "")
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            res[cur]["_total"] += 1
            for w in WATCH:
                if w in op:
                    res[cur][w] += 1
    for name, c in res.items():
        if only and not any(o in name for o in only):
            continue
        items = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{name:58s} {c['_total']:6d} instr  {items}")


if __name__ == "__main__":
    print("# cuobjdump -sass myslam_b200/libeslam_b200.so: instructions per kernel and counts of selected mnemonics")
    scan(os.path.join(ROOT, "myslam_b200", "libeslam_b200.so"))
    mb = os.path.join(ROOT, "tools", "microbench", "l2_gather_red")
    if os.path.exists(mb):
        print("# tools/microbench/l2_gather_red (the bulk-async variants measured against LDG / REDG)")
        scan(mb)
