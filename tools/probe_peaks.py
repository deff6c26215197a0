import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
r = pkg.vrt.Renderer(0)
print("ffma TF", r.fp32_peak(False), "ffma2 TF", r.fp32_peak(True))
for pairs in (5, 10, 20):
    for ctas in (1, 2, 3, 4):
        try:
            v = r.term_peak(pairs, ctas)
            print(f"pairs {pairs} ctas/SM {ctas}: {v:.3e} terms/s = {v*15/74.45e12:.3f} of roofline")
        except Exception as e:
            print(pairs, ctas, e)

clk = 1.965e9
for nf, nm, nl in ((9, 0, 0), (9, 0, 2), (9, 1, 0), (9, 2, 0), (9, 2, 2), (0, 2, 0), (1, 2, 0), (4, 2, 0), (18, 2, 0)):
    v = r.mix_peak(nf, nm, nl)
    cyc = 148 * 4 * 32 * clk / v  # SMSP-cycles per warp-step
    print(f"mix nf={nf} nm={nm} nl={nl}: {v:.3e} steps/s -> {cyc:.2f} cycles per warp-step per SMSP")
