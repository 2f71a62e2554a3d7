"""Per-layer CUDA-event timing of EnCodec-32k encode + decode (32 x 10 s, the per-GPU slice of BASELINE configs[3]).
   python profiles/perf_encodec.py [--batch 32] [--seconds 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiocraft_b200.loaders import load_compression_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--seconds', type=float, default=10.0)
ap.add_argument('--enc', default='fp32_tc')
ap.add_argument('--dec', default='tf32x3')
a = ap.parse_args()
from audiocraft_b200 import synth  # noqa: E402
from audiocraft_b200.encodec import EncodecModel  # noqa: E402
_cfg = synth.ENCODEC_CONFIGS['encodec_32k']
cm = EncodecModel(synth.synth_encodec_state_dict(_cfg, 0), _cfg, 'cuda', encoder_precision=a.enc, decoder_precision=a.dec)
x = torch.randn(a.batch, 1, int(a.seconds * 32000), device='cuda') * 0.1
for _ in range(2):
    codes, _ = cm.encode(x)
    cm.decode(codes)
torch.cuda.synchronize()
cm._profile = []
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
codes, _ = cm.encode(x)
e1.record()
y = cm.decode(codes)
e2.record()
torch.cuda.synchronize()
tot = 0.0
print(f'{"layer":46s} {"in":>22s} {"out":>22s} {"ms":>8s} {"GFLOP":>8s} {"TFLOP/s":>8s} {"GB/s":>8s}')
for L, si, so, a0, a1 in cm._profile:
    ms = a0.elapsed_time(a1)
    tot += ms
    if L['kind'] == 'conv':
        fl = 2.0 * so[0] * so[2] * L['cout'] * L['cin'] * L['k']
    elif L['kind'] == 'convtr':
        fl = 2.0 * si[0] * si[2] * L['cout'] * L['cin'] * L['k']
    else:
        fl = 2.0 * si[0] * si[2] * L['layers'] * 2 * 4 * L['dim'] * L['dim']
    byt = 4.0 * (si[0] * si[1] * si[2] + so[0] * so[1] * so[2])
    name = f"{L['kind']} {L['prefix'][:-1]} k={L.get('k', '-')} s={L.get('stride', '-')}"
    print(f'{name:46s} {str(si):>22s} {str(so):>22s} {ms:8.3f} {fl / 1e9:8.1f} {fl / ms / 1e9:8.2f} {byt / ms / 1e6:8.0f}')
n = x.numel()
print(f'layers total {tot:.2f} ms; encode {e0.elapsed_time(e1):.2f} ms, decode {e1.elapsed_time(e2):.2f} ms, '
      f'{n / (e0.elapsed_time(e2) / 1e3) / 1e6:.1f} MSamples/s enc+dec')
