import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_bench as q
kind = "emission_absorption_physical"
for N in (1, 2, 3, 4):
    q.run(kind, N, 1000, 1000, 100000, False)
q.run(kind, 2, 334, 166, 100000, False)
