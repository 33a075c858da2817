import sys, time
sys.path.insert(0, '.')
import numpy as np
import indelminer_b200
from indelminer_b200 import synth
R = indelminer_b200.Realigner()
for n in (30, 3000, 8000, 16000, 32000, 65536):
    t = synth.make_support_tasks(n, seed=5)
    packed = (t["targets"], t["target_off"], t["queries"], t["query_off"])
    for _ in range(20): R.indel_support_batch(None, None, packed=packed)
    t0 = time.perf_counter()
    K = 100
    for _ in range(K): R.indel_support_batch(None, None, packed=packed)
    dt = (time.perf_counter() - t0) / K
    print(f"pairs={n} per call {dt*1e3:.3f} ms")
