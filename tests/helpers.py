"""Shared seeded-input builders for the parity tests and the golden-vector generator."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from audiocraft_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def audio_input(cfg: dict, batch: int, length: int, seed: int) -> torch.Tensor:
    """Seeded audio-like input: a few sinusoids + noise, amplitude ~0.3."""
    g = torch.Generator()
    g.manual_seed(seed)
    t = torch.arange(length, dtype=torch.float32) / cfg['sample_rate']
    x = torch.zeros(batch, cfg['channels'], length)
    for b in range(batch):
        for c in range(cfg['channels']):
            f = 110.0 * (1 + torch.rand(3, generator=g) * 8)
            a = torch.rand(3, generator=g) * 0.2
            ph = torch.rand(3, generator=g) * 6.28
            x[b, c] = (a.view(-1, 1) * torch.sin(6.2831853 * f.view(-1, 1) * t.view(1, -1) + ph.view(-1, 1))).sum(0)
    return x + 0.05 * torch.randn(x.shape, generator=g)


def lm_condition(cfg: dict, sd: dict, batch: int, t_text: int, seed: int):
    """(hidden [B,T,cond_dim], mask [B,T], cross_cfg [2B,T,d]) -- the rows the reference's condition provider
    returns for conditions + null conditions: output_proj(hidden) * mask, then exact zeros for the null rows
    (audiocraft/modules/conditioners.py:509-515, audiocraft/models/lm.py:497-509)."""
    hid, mask = synth.synth_text_condition(cfg, batch, t_text, seed)
    w = sd['condition_provider.conditioners.description.output_proj.weight'].float().cpu()
    b = sd['condition_provider.conditioners.description.output_proj.bias'].float().cpu()
    emb = (hid @ w.t() + b) * mask.unsqueeze(-1)
    return hid, mask, torch.cat([emb, torch.zeros_like(emb)], dim=0)


def exp_noise(seed: int, step: int, rows: int, card: int) -> torch.Tensor:
    """Exponential(1) noise for the injected-noise multinomial, one independent stream per step."""
    g = torch.Generator()
    g.manual_seed(seed * 100003 + step)
    return torch.empty(rows, card).exponential_(1, generator=g)


def fullsize_sequence(cfg: dict, batch: int, T: int, seed: int) -> torch.Tensor:
    """Seeded codes [B, K, T] laid out in the delay pattern (codebook k shifted by delays[k] + 1, `card` = special token
    elsewhere): the [B, K, T + max_delay + 1] sequence LMModel.generate builds with build_pattern_sequence
    (audiocraft/modules/codebooks_patterns.py:154-179, 339-356)."""
    g = torch.Generator()
    g.manual_seed(seed)
    K, card, delays = cfg['n_q'], cfg['card'], cfg['delays']
    codes = torch.randint(0, card, (batch, K, T), generator=g)
    S = T + max(delays) + 1
    seq = torch.full((batch, K, S), card, dtype=torch.long)
    for k in range(K):
        seq[:, k, 1 + delays[k]:1 + delays[k] + T] = codes[:, k]
    return seq
