#!/usr/bin/env python
"""Developer probe: FP64 vector and tensor (mma.sync f64, by shape) peaks of the device."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exp_b200 import _lib
L = _lib.lib()
tf, hint = C.c_double(), C.c_double()
_lib.check(L.expb200_fp64_peak(0, C.byref(tf), C.byref(hint)))
print(f"DFMA peak {tf.value:.2f} TFLOP/s")
for k, nm in enumerate(("m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16")):
    _lib.check(L.expb200_dmma_peak(0, k, C.byref(tf)))
    print(f"DMMA {nm:9s} {tf.value:.2f} TFLOP/s")
