"""Build tuning variants of libb200dr.so (see nvdiffrast_b200/build.py:build_variant).  Usage: python scripts/build_variants.py name=-DX=1,-DY=2 ..."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nvdiffrast_b200 import build as b
for a in sys.argv[1:]:
    name, defs = a.split("=", 1)
    out = b.build_variant(name, [d for d in defs.split(",") if d])
    log = open(os.path.join(os.path.dirname(b.SO_PATH), "csrc", "_build", "var_" + name, "build.log")).read()
    regs = re.findall(r"Compiling entry function '(\S*?(?:texture_grad_fastILi3ELi3ELi1|texture_fwd_fastILi3ELi3ELi1|raster_grad_kernelILb1|raster_bins_kernelILb0ELb0)\S*)'.*?\n.*?\n.*?spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log)
    print(name, [(r[0][-40:], "spill", r[1], "regs", r[2]) for r in regs])
