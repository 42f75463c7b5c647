import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from pypercolate_b200 import _native, lowering
g = lowering.lowered_spanning_2d_grid(256)
ctx = _native.Context(0); ctx.set_graph(g); ctx.set_ps(np.linspace(0.45, 0.55, 100))
R = 296
seeds = np.arange(R, dtype=np.uint32) * 7 + 1
ctx.run_fused(R, _native.PERM_FEISTEL, seeds, _native.FUSE_MICRO)
ctx.synchronize()
