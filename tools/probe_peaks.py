"""Roofline probes on the GPU itself: FFMA/FFMA2 throughput, ceiling of K2's loop body, pipe-mix costs.
Run on the GPU box:  python tools/probe_peaks.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
r = pkg.vrt.Renderer(0)
print("ffma TFLOP/s", r.fp32_peak(False), "ffma2 TFLOP/s", r.fp32_peak(True))
for pairs in (10, 20, -10, -20):
    for ctas in (1, 2, 4):
        v = r.term_peak(pairs, ctas)
        kind = "signed body" if pairs > 0 else "sign-uniform body"
        print(f"{kind}: {abs(pairs)} pairs/thread, {ctas} CTAs/SM: {v:.3e} terms/s = {v * 15 / 74.45e12:.3f} of the roofline")
clk = 1.965e9
for nf, nm, nl in ((9, 0, 0), (9, 0, 2), (9, 1, 0), (9, 2, 0), (9, 2, 2), (1, 2, 0), (4, 2, 0), (18, 2, 0)):
    v = r.mix_peak(nf, nm, nl)
    cyc = 148 * 4 * 32 * clk / v  # SMSP-cycles per warp-step
    print(f"mix: {nf} FFMA2 + {nm} MUFU.RCP + {nl} LOP3 per step: {v:.3e} steps/s -> {cyc:.2f} cycles per warp-step per SMSP")
