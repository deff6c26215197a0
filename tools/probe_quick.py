import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
r = pkg.vrt.Renderer(0)
for pairs in (10, 20):
    for ctas in (2, 4):
        v = r.term_peak(pairs, ctas)
        print(f"pairs {pairs} ctas/SM {ctas}: {v:.3e} terms/s = {v*15/74.45e12:.3f} of roofline")
