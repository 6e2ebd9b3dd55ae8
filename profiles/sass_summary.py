"""SASS evidence for the claims in DESIGN.md (run on the CPU box; needs only cuobjdump + the built libprt.so):
per kernel, the instruction count and the mnemonics that matter -- packed FP32 (FFMA2), 256-bit global loads
(LDG.E...256), byte->float conversions (I2F.U8 / PRMT), FP64 (DFMA/DMUL/DADD), warp votes / REDUX, shared-memory
stack traffic (LDS / STS), local-memory traffic (LDL / STL), and that the cubins are sm_100a only.
usage: python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyrenderer_b200.kernel_fingerprint import fingerprint  # noqa: E402

LIB = os.path.join(ROOT, "pyrenderer_b200", "libprt.so")
PATTERNS = [("FFMA2", r"\bFFMA2\b"), ("FFMA", r"\bFFMA\b"), ("LDG.256", r"\bLDG\.E[.\w]*\.256"), ("LDG.128", r"\bLDG\.E[.\w]*\.128"),
            ("LDG other", r"\bLDG\.E(?![.\w]*\.(256|128))"), ("I2F.U8", r"\bI2F[.\w]*U8"), ("PRMT", r"\bPRMT\b"), ("FMNMX", r"\bFMNMX3?\b"),
            ("DFMA/DMUL/DADD", r"\b(DFMA|DMUL|DADD)\b"), ("VOTE", r"\bVOTEU?\b"), ("REDUX", r"\b(REDUX|CREDUX)\b"), ("SHFL", r"\bSHFL\b"),
            ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("ATOMS", r"\bATOMS\b"), ("LDL", r"\bLDL\b"), ("STL", r"\bSTL\b"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"),
            ("BAR/grid sync", r"\bBAR\b")]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return out if len(out) == len(names) else names


def main():
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line)
    names = list(funcs)
    pretty = demangle(names)
    print(f"# SASS summary of pyrenderer_b200/libprt.so (cuobjdump -sass), source fingerprint {fingerprint()}")
    print("# cubins: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
    hdr = f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>9s}" for k, _ in PATTERNS)
    print(hdr)
    rows = []
    for n, p in zip(names, pretty):
        body = "\n".join(funcs[n])
        short = p.replace("(int)", "").replace("(bool)", "").replace("prt::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("<unnamed>::", "")
        rows.append((short, len(funcs[n]), [len(re.findall(rx, body)) for _, rx in PATTERNS]))
    for short, n, counts in sorted(rows):
        print(f"{short[:78]:78s} {n:6d} " + " ".join(f"{c:9d}" for c in counts))


if __name__ == "__main__":
    main()
