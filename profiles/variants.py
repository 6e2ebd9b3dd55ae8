"""Build (on the CPU box) or run (on the GPU box) tuning variants of libprt.so.
  python profiles/variants.py build     -> pyrenderer_b200/variants/<name>.so
  python profiles/variants.py run       -> soup-1M Mrays/s + Cornell render per variant (subprocess each)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "pyrenderer_b200", "variants")
# name -> extra nvcc defines.  Round-1 findings (profiles/r1_sweeps.txt): speculative leaf postponement
# -6 %, PRMT+FADD byte->float instead of I2F.U8 -8 %, L1 prefetch of the far child 0 %, smem stack
# depth 8/12/16 no effect -- none of them is in the tree any more; add -D switches here to try new ones.
VARIANTS = {
    "base": (),
    "smin": ("-DPRT_STACK_MIN=1",),
}
if sys.argv[1] == "build":
    from pyrenderer_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        print(name, build.build(force=True, defines=defs, out=os.path.join(VDIR, name + ".so")))
elif sys.argv[1] == "exact":
    for name in VARIANTS:
        env = dict(os.environ, PRT_LIB=os.path.join(VDIR, name + ".so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "prof_exact.py")] + sys.argv[2:], env=env, capture_output=True, text=True)
        print(f"[{name}] " + " | ".join(l for l in r.stdout.splitlines() if l.startswith("{")) + ("" if r.returncode == 0 else " ERR " + r.stderr[-300:]), flush=True)
else:
    for name in VARIANTS:
        env = dict(os.environ, PRT_LIB=os.path.join(VDIR, name + ".so"))
        for mode in ("util", "render1"):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sweep.py"), mode], env=env, capture_output=True, text=True)
            lines = [l for l in r.stdout.splitlines() if "util:" in l or l.startswith("(") or l.startswith("render")]
            print(f"[{name}] {mode}: " + " | ".join(lines) + ("" if r.returncode == 0 else " ERR " + r.stderr[-300:]), flush=True)
