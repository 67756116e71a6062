"""Static evidence from the shipped library (no GPU needed): per kernel, the registers / stack (spills) / shared
memory ptxas allotted (`cuobjdump -res-usage`) and the counts of the SASS mnemonics that show which hardware paths a
kernel uses — 128-bit global loads/stores, cp.async (LDGSTS), TMA bulk copies (UBLKCP), mbarrier (SYNCS), cluster /
distributed-shared-memory instructions, atomics / reductions, and FMAs. The library is built with -fmad=false, so
no a*b+c of the source is contracted; the FFMA / DFMA that remain are the inside of correctly rounded divisions
(-prec-div=true / -prec-sqrt=true) and of libdevice's log / lgamma / exp — replay parity with the
x86-64 reference (bit-exact tests) is the check that none of them changes a result.

  python tools/sass_evidence.py > profiles/<round>_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fba-pomdp_b200", "libfba_b200.so")

PATTERNS = [
    ("LDG.128", re.compile(r"\bLDG\.E\S*\.128")), ("STG.128", re.compile(r"\bSTG\.E\S*\.128")),
    ("LDGSTS", re.compile(r"\bLDGSTS")), ("UBLKCP", re.compile(r"\bUBLKCP")), ("SYNCS", re.compile(r"\bSYNCS")),
    ("cluster", re.compile(r"\bUCGABAR|\bCGAERRBAR|\bMAPA|\bST\.\S*\.SHARED::CLUSTER|\bLD\.\S*\.SHARED::CLUSTER|\bSTS\.\S*CLUSTER|\bCCTL\.\S*CLUSTER")),
    ("ATOM/RED", re.compile(r"\bATOMG?\b|\bATOMG?\.|\bRED\.|\bATOMS\.")),
    ("SHFL", re.compile(r"\bSHFL\.")), ("DADD", re.compile(r"\bDADD\b")), ("DFMA", re.compile(r"\bDFMA\b")),
    ("FFMA", re.compile(r"\bFFMA\b")),
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(sig):
    sig = sig.replace("fba::", "")
    m = re.match(r"(?:void )?([A-Za-z_0-9]+(?:<[^(]*>)?)\(", sig)
    return m.group(1) if m else sig[:60]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        line = line.strip()
        if line.startswith("Function "):
            fn = line[len("Function "):].rstrip(":")
        elif line.startswith("REG:") and fn:
            usage[fn] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    n_inst = collections.Counter()
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        n_inst[fn] += 1
        for tag, pat in PATTERNS:
            if pat.search(ins):
                counts[fn][tag] += 1
    names = demangle(sorted(usage))
    tags = [t for t, _ in PATTERNS]
    print("# %s — sm_100a, %d kernels; columns: registers, stack bytes (per-thread local arrays and spills), static shared bytes, "
          "SASS instructions, then mnemonic counts" % (os.path.relpath(LIB, ROOT), len(usage)))
    print("%-58s %4s %5s %6s %6s " % ("kernel", "reg", "stack", "smem", "inst") + " ".join("%8s" % t for t in tags))
    for fn in sorted(usage, key=lambda f: short(names[f])):
        u = usage[fn]
        print("%-58s %4s %5s %6s %6d " % (short(names[fn])[:58], u.get("REG"), u.get("STACK"), u.get("SHARED"), n_inst[fn])
              + " ".join("%8d" % counts[fn][t] for t in tags))
    spilled = [short(names[f]) for f in usage if int(usage[f].get("STACK", 0)) > 0]
    fma = [short(names[f]) for f in usage if counts[f]["FFMA"] + counts[f]["DFMA"] > 0]
    print("\nkernels with a stack frame: %s" % (", ".join(sorted(set(spilled))) or "none"))
    print("kernels containing FFMA/DFMA: %s" % (", ".join(sorted(set(fma))) or "none"))


if __name__ == "__main__":
    sys.exit(main())
