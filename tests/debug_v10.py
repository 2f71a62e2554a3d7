"""Bring-up of the default decode step (cluster split-K GEMMs, LayerNorm on load: csrc/lm.cu lm_gemm2_kernel) against the round-1
11-kernel layer (ACB_LM_STEP=v9).  Not a pytest file.

  python tests/debug_v10.py [arch] [batch] [n_steps]     teacher-forced logits, v10 vs v9, per step; then greedy generation
"""
import faulthandler
import os
import sys
import time

faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from audiocraft_b200 import synth  # noqa: E402
from audiocraft_b200.lm import LMModel  # noqa: E402
from tests import helpers as H  # noqa: E402


def log(*a):
    print(f'[{time.time() % 1000:7.2f}]', *a, flush=True)


arch = sys.argv[1] if len(sys.argv) > 1 else 'lm_mini'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cfg = synth.lm_config(arch)
sd = synth.synth_lm_state_dict(cfg, seed=3)
_, _, cross = H.lm_condition(cfg, sd, B, 5, 1)
m = LMModel(sd, cfg, None, None)
g = torch.Generator().manual_seed(5)
seq = torch.randint(0, m.card, (B, m.n_q, T + 4), generator=g)
seq[:, :, 0] = m.card


def run(mode):
    os.environ['ACB_LM_STEP'] = mode
    out = m.teacher_forced_logits(seq, cross, 3.0, n_steps=T)
    torch.cuda.synchronize()
    return out


lv = run('v9')
log(f'{arch} B={B}: v9 done, launches/step {m.launches_per_step}')
lf = run('v10')
log(f'v10 done, launches/step {m.launches_per_step}')
bad = False
for i in range(T):
    e = (lf[i] - lv[i]).abs().max().item()
    log(f'step {i}: v10 vs v9 {e:.3e}   |logit| max {lv[i].abs().max():.2f}')
    bad |= not (e <= 2e-2)
os.environ['ACB_LM_STEP'] = 'v10'
a = m.generate(None, [], num_samples=B, max_gen_len=T + 8, use_sampling=False, cross_attention_src=cross)
os.environ['ACB_LM_STEP'] = 'v9'
b = m.generate(None, [], num_samples=B, max_gen_len=T + 8, use_sampling=False, cross_attention_src=cross)
same = (a == b).float().mean().item()
log('greedy token agreement v10 vs v9', same)
log('V10', 'FAILED' if bad or same < 0.95 else 'OK')
sys.exit(1 if bad or same < 0.95 else 0)
