"""Instruction census of the shipped library's kernels (no GPU needed):

    python tools/sass_census.py > profiles/r02_sass_census.txt

cuobjdump -sass on horizonator_b200/lib/libhorizonator.so, per kernel: instructions, the depth-test reduction
(REDG.E.MIN.64), compare-and-swap loops (ATOMS.CAST), other atomics, local-memory loads/stores (register spills and
stack), MUFU; then ptxas -v's registers / stack / spill bytes for the same sources."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "horizonator_b200", "lib", "libhorizonator.so")
PATTERNS = [("REDG.E.MIN.64", r"\bREDG\.E\.MIN\.64"), ("ATOMS.CAST", r"\bATOMS\.CAST"), ("ATOM.E.ADD", r"\bATOM\.E\.ADD"),
            ("ATOMS.ADD", r"\bATOMS\.ADD"), ("ATOM.*MIN", r"\bATOMG?\.E\.MIN"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"),
            ("MUFU", r"\bMUFU"), ("CALL", r"\bCALL\.")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = rows.setdefault(m.group(1), {"instr": 0, **{k: 0 for k, _ in PATTERNS}})
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            continue
        cur["instr"] += 1
        for k, pat in PATTERNS:
            if re.search(pat, line):
                cur[k] += 1
    print("# cuobjdump -sass %s (sm_100a): instruction census per kernel" % os.path.relpath(LIB, ROOT))
    print("%-38s %7s " % ("kernel", "instr") + " ".join("%13s" % k for k, _ in PATTERNS))
    for name in sorted(rows):
        r = rows[name]
        print("%-38s %7d " % (name, r["instr"]) + " ".join("%13d" % r[k] for k, _ in PATTERNS))
    print()
    print("# ptxas -v for horizonator_b200/csrc/hz_kernels.cu (same flags as the build)")
    out = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
                          "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "horizonator_b200", "csrc"),
                          "-c", os.path.join(ROOT, "horizonator_b200", "csrc", "hz_kernels.cu"), "-o", os.devnull],
                         capture_output=True, text=True).stderr
    name = None
    for line in out.splitlines():
        line = line.replace("ptxas info    : ", "").strip()
        m = re.match(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            continue
        if name and line.startswith("Used"):
            print("%-38s %s; %s" % (name, line, frame))
            name = None
        elif name and "stack frame" in line:
            frame = line
    return 0


if __name__ == "__main__":
    sys.exit(main())
