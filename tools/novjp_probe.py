"""A/B of the fused kernel's instantiation without the VJP plumbing (CHI_EA_NOVJP=0|1|2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_bench as q
kind = "emission_absorption_physical"
for pv in (False, True):
    for N, B in ((2, 100000), (3, 100000), (4, 100000), (5, 50000), (8, 25000)):
        q.run(kind, N, 1000, 1000, B, pv)
