import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
os.environ.setdefault("PYTEST_CURRENT_TEST", "x")
import test_gpu_round2 as t
ref = t._l256_statistics("PERM_MT19937")
for rounds in (20, 8, 6, 4, 3, 2):
    os.environ["PZ_FEISTEL_ROUNDS"] = str(rounds)
    t._STAT_CACHE.pop(("PERM_FEISTEL", 10000, False), None)
    got = t._l256_statistics("PERM_FEISTEL")
    z = t._z_scores(ref, got)
    print("rounds %2d: max|z| %.2f  mean z^2 %.2f  KS %.2f" % (rounds, np.abs(z).max(), np.mean(z * z), t._ks_first_spanning(ref, got)), flush=True)
os.environ.pop("PZ_FEISTEL_ROUNDS", None)
