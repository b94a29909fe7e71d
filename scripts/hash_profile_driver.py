import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.hash_microbench import run
for mode in ("reference", "trilinear"):
    print(run(1 << 22, 19, mode))
