"""Stage times of a large batch whose two spectra are on DIFFERENT axes (the reference tests' shape: 1000
emission + 500 absorption channels), i.e. of k_moments_em + k_moments_abs instead of the fused kernel
(development aid).  usage: [CHI_LIB=...] python tools/two_axes_bench.py [draws]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_bench as qb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
qb.run("emission_absorption_physical", 4, 1000, 500, B, False)
qb.run("emission_absorption_physical", 4, 1000, 500, B, True)
