"""us per grid-wide barrier of a co-resident cooperative kernel (acb_debug_grid_barrier): the floor per dependent phase
of the persistent decode step.  Run on the GPU box:  python profiles/perf_grid_barrier.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from audiocraft_b200 import _lib  # noqa: E402

L = _lib.lib()
torch.cuda.init()
sms = L.acb_device_sm_count(0)
us = C.c_float(0)
for variant, name in ((0, 'red.release + ld.acquire poll'), (1, 'red.release + relaxed poll + fence'), (2, 'fence + atomicAdd + volatile spin')):
    for threads in (128, 288, 544):
        for work in (0, 200):
            _lib.check(L.acb_debug_grid_barrier(sms, threads, 2000, variant, work, 5, C.byref(us)))
            print(f'variant {variant} ({name}), {sms} CTAs x {threads} threads, {work} dependent FMAs between barriers: {us.value:.3f} us per barrier')
for ctas in (16, 74):
    _lib.check(L.acb_debug_grid_barrier(ctas, 288, 2000, 0, 0, 5, C.byref(us)))
    print(f'variant 0, {ctas} CTAs x 288 threads: {us.value:.3f} us per barrier')
