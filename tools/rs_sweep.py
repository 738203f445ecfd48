"""Ring-depth sweep of the streaming row kernel (csrc/rls_rowstream.cu) at the benchmark shapes, after a correctness
pass on ragged shapes.  usage: python tools/rs_sweep.py [NS ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.rowstream_probe import timeit, setenv, check
shapes = [(16384, 65536, np.float32), (8192, 65536, np.complex64), (16384, 16384, np.complex64), (65536, 16384, np.float32), (131072, 8192, np.float32)]
ok = True
for dt in (np.float32, np.complex64):
    for (m, n) in [(5, 7), (130, 4099), (67, 20000), (41, 65536), (500, 65536), (35, 40000), (777, 16384)]:
        ok &= check(m, n, dt)
print("ALL OK" if ok else "FAILURES", flush=True)
for ns in [int(a) for a in sys.argv[1:]] or [0, 3, 4, 5, 6]:
    setenv(RLS_ROWSTREAM_NS=ns or None)
    for (m, n, dt) in shapes:
        timeit(m, n, dt, label=f"NS={ns or 'default'}")
