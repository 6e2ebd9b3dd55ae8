"""Content hash of the CUDA sources behind libprt.so.

profiles/ncu_*.json (hardware counters captured under ncu, which cannot run inside a timed bench)
carry the fingerprint of the sources they were captured from; bench.py quotes such a counter only
when the fingerprint still matches and says "stale" otherwise."""
import hashlib
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def fingerprint():
    h = hashlib.sha256()
    csrc = os.path.join(HERE, "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh")))
    files.append(os.path.join(HERE, "..", "include", "prt.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]
