"""Which kernels changed since a commit?  Compiles every csrc/*.cu of <commit> and of the working tree to PTX and compares
each kernel's text with register numbers, labels and mangled names normalised.  Used at the end of round 1 to show that the
experimental additions (robust pruned sweep, step-kernel variants) left every default-path kernel exactly as it was when
it was last verified on hardware:

    python scripts/ptx_same_as.py a57700b
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "torchdr_b200/csrc"
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-ptx"]


def kernels(ptx):
    out, name, buf = {}, None, []
    for line in ptx.splitlines():
        m = re.search(r"\.entry\s+(\S+?)\(", line)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
        if name:
            line = re.sub(r"_Z[A-Za-z0-9_]+", "SYM", line)
            line = re.sub(r"%(rd|rs|rh|fd|r|f|p)\d+", r"%\1N", line)
            line = re.sub(r"\$L__BB\d+_", "$L_", line)
            line = re.sub(r"\.b8 SYM\[\d+\]", ".b8 SYM[P]", line)
            buf.append(line)
    if name:
        out[name] = buf
    return out


def demangle(names):
    try:
        return subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    except OSError:
        return names


def build(tree):
    res = {}
    for f in sorted(os.listdir(os.path.join(tree, SRC))):
        if f.endswith(".cu"):
            out = os.path.join(tree, f + ".ptx")
            subprocess.run(["nvcc"] + FLAGS + [os.path.join(tree, SRC, f), "-o", out], check=True, capture_output=True)
            res.update(kernels(open(out).read()))
    return res


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(f"git -C {ROOT} archive {commit} {SRC} include | tar -x -C {tmp}", shell=True, check=True)
        old, new = build(tmp), build(ROOT)
    names = sorted(set(old) | set(new))
    pretty = dict(zip(names, demangle(names)))
    # a kernel that only gained template parameters keeps its text: pair it with its old self (the first line holds the
    # mangled name, already normalised away)
    gone = {n: old[n] for n in old if n not in new}
    for n in names:
        if n in gone:
            continue
        if n not in old:
            twin = next((o for o, body in gone.items() if body == new[n]), None)
            state = "new" if twin is None else "same*"
            if twin is not None:
                gone.pop(twin)
        else:
            state = "same" if old[n] == new[n] else "CHANGED"
        print(f"{state:8s} {pretty[n][:120]}")
    for n in gone:
        print(f"{'removed':8s} {pretty[n][:120]}")
    print("same* = same text under a new (templated) name")


if __name__ == "__main__":
    main()
