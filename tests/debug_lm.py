"""Stage-by-stage LM bring-up with per-launch synchronisation (ACB_DEBUG=1). Not a pytest file."""
import faulthandler
import os
import sys
import time

os.environ['ACB_DEBUG'] = '1'
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from audiocraft_b200 import synth  # noqa: E402
from audiocraft_b200.lm import LMModel  # noqa: E402
from tests import helpers as H  # noqa: E402
from oracle import lm_oracle as LO  # noqa: E402


def log(*a):
    print(f'[{time.time() % 1000:.2f}]', *a, flush=True)


cfg = synth.lm_config('lm_mini')
sd = synth.synth_lm_state_dict(cfg, seed=3)
B, T = 2, 10
_, _, cross = H.lm_condition(cfg, sd, B, 5, 1)
log('building model')
m = LMModel(sd, cfg, None, None)
torch.cuda.synchronize()
log('model built')
seq = torch.full((B, 4, T + 4), cfg['card'], dtype=torch.long)
lg = m.teacher_forced_logits(seq, cross, 3.0, n_steps=2)
torch.cuda.synchronize()
log('teacher forced ok', lg.shape, float(lg.abs().max()))
o = LO.LMOracle(sd, cfg, half_gemm=True)
logs = []
o.generate(None, cross, B, T, use_sampling=False, record_logits=logs, teacher=seq)
log('max diff vs oracle', float((lg.cpu() - torch.stack(logs[:2])).abs().max()))
out = m.generate(None, [], num_samples=B, max_gen_len=T, use_sampling=False, cross_attention_src=cross)
torch.cuda.synchronize()
log('generate ok', out[0, 0].tolist())
