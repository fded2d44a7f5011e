"""K6 probe: gnofix on chr1, n individuals with 20 planted switch errors each (python scripts/gnofix_probe.py [n])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts import bench_configs as bc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
print(bc.cfg5(2 * n))
