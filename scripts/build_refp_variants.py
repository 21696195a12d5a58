"""Tuning builds of the persistent Mode R kernel: variants/lib_<name>.so, reusing the main build's objects
for every other file.  usage: build_refp_variants.py name:-Dflag,-Dflag ..."""
import glob, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from teeline_b200 import build as tb
tb.build()
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    od = os.path.join(ROOT, "build", f"obj_{name}")
    os.makedirs(od, exist_ok=True)
    for o in glob.glob(os.path.join(ROOT, "build", "obj", "*.o")):
        if not o.endswith("k2_two_opt_ref.o"):
            shutil.copy2(o, od)
            os.utime(os.path.join(od, os.path.basename(o)))
    if os.path.exists(os.path.join(od, "k2_two_opt_ref.o")):
        os.remove(os.path.join(od, "k2_two_opt_ref.o"))
    tb.build(extra=[f for f in flags.split(",") if f], lib=os.path.join(ROOT, "variants", f"lib_{name}.so"), obj_dir=od)
    print("built", name)
