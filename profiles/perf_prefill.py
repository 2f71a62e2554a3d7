"""Prompt prefill vs feeding the prompt one decode step at a time (MusicGen-medium, batch 8 = CFG rows 16, 12 s prompt = 600
frames): CUDA-event time of acb_lm_prefill against the same positions through the captured step graph.
    python profiles/perf_prefill.py [--scale medium] [--batch 8] [--frames 600]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiocraft_b200 import _lib  # noqa: E402
from audiocraft_b200.loaders import load_lm_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--scale', default='medium')
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--frames', type=int, default=600)
a = ap.parse_args()
lm = load_lm_model(f'synthetic/{a.scale}')
B, S = a.batch, 1504
cross = torch.randn(2 * B, 16, lm.dim, device='cuda') * 0.1
cross[B:] = 0
lm._ensure(2 * B, S, 16, B)
samp = _lib.LMSampling(1, 1.0, 250, 0.0, 3.0, 1, 0)
lm._bufs['seq'][:B].random_(0, lm.card)
lm._bufs['seq_mask'].fill_(1)
_lib.check(lm._lib.acb_lm_begin(lm._handle, _lib.ptr(cross), B, 2 * B, 16, S, C.byref(samp), _lib.stream()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for what in ('prefill', 'steps'):
    for it in range(2):
        lm._bufs['pos'][0] = 0
        torch.cuda.synchronize()
        e0.record()
        if what == 'prefill':
            _lib.check(lm._lib.acb_lm_prefill(lm._handle, 0, a.frames, _lib.stream()))
        else:
            _lib.check(lm._lib.acb_lm_steps(lm._handle, a.frames, _lib.stream()))
        e1.record()
        torch.cuda.synchronize()
    print(f'{what:8s}: {a.frames} prompt positions, rows {2 * B}: {e0.elapsed_time(e1):8.1f} ms  '
          f'({e0.elapsed_time(e1) / a.frames * 1e3:.1f} us per position)')
