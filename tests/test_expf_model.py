"""The reference's shaded-sample count hangs on `throughput < 0.1` (example_tracer.cpp:131) where throughput is a
product of libm expf results (:175,189).  The CUDA path therefore carries a restatement of glibc's expf algorithm
(lrf_device.cuh expf_libm; table from tools/gen_exp2f_table.py) instead of CUDA's expf (<= 2 ulp, different
roundings).  This test pins the restatement (oracle twin orc_expf_model, same constants) against the libm of the
machine the tests run on — the GPU box runs it too, so a host whose libm rounds differently is detected."""
import ctypes
import ctypes.util
import subprocess
import sys

import numpy as np


def test_expf_model_equals_host_libm(orc):
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(42)
    xs = np.concatenate([
        -np.abs(rng.standard_normal(60000)).astype(np.float32) * np.float32(2.5),      # the range -sigma*delta lives in
        -rng.random(30000).astype(np.float32) * np.float32(0.05),
        -np.exp(rng.uniform(-30, 5, 20000)).astype(np.float32),
        np.float32([0.0, -0.0, -1e-45, -87.0, -88.5, -103.0, -103.97, -104.0, -200.0, -np.inf]),
    ])
    bad = 0
    for x in xs:
        a, b = libm.expf(float(x)), orc.orc_expf_model(float(x), 1)
        if np.float32(a).view(np.uint32) != np.float32(b).view(np.uint32):
            bad += 1
    assert bad == 0, "%d of %d arguments round differently from this host's libm expf" % (bad, xs.size)


def test_exp2f_table_is_reproducible():
    """the 32-entry table in lrf_device.cuh / lrf_oracle.cpp is what tools/gen_exp2f_table.py derives from first principles"""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_exp2f_table.py")], capture_output=True, text=True, check=True).stdout
    want = [w.strip().rstrip(",") for w in out.replace("\n", " ").split() if w.startswith("0x")]
    assert len(want) == 32
    for f in ("literf-radiancefields_b200/csrc/lrf_device.cuh", "oracle/lrf_oracle.cpp"):
        src = open(os.path.join(root, f)).read()
        for w in want:
            assert w in src, (f, w)
