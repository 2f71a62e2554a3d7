"""Model assembly (mirror of the parts of audiocraft/models/loaders.py + builders.py the path needs).

The reference resolves names through the HF hub and re-instantiates modules from an omegaconf ``xp.cfg``
(loaders.py:78-126, builders.py:70-175).  Here: a checkpoint in the reference's export format
(``{'best_state': state_dict, 'xp.cfg': yaml}``, utils/export.py:20-79) found on local disk is loaded with the
hyper-parameters read from its yaml (PyYAML; omegaconf is not in the image), and ``synthetic/<arch>`` builds seeded
random weights of a released architecture (no pretrained weights exist offline).
"""
import os
import typing as tp

import torch

from . import synth
from .conditioners import ConditionFuser, ConditioningProvider, PrecomputedTextConditioner
from .encodec import EncodecModel
from .lm import LMModel

_SCALES = {'small': 'musicgen_small', 'medium': 'musicgen_medium', 'large': 'musicgen_large'}


def _cfg_from_yaml(text: str) -> dict:
    import yaml
    return yaml.safe_load(text)


def load_compression_model(name: str, device='cuda', seed: int = 0) -> EncodecModel:
    if name.startswith('synthetic/'):
        arch = name.split('/', 1)[1]
        cfg = synth.ENCODEC_CONFIGS[arch]
        return EncodecModel(synth.synth_encodec_state_dict(cfg, seed), cfg, device)
    if os.path.isfile(name):
        pkg = torch.load(name, map_location='cpu', weights_only=False)
        if 'pretrained' in pkg:
            raise RuntimeError(f"{name} redirects to '{pkg['pretrained']}' (HF hub); no network in this image")
        xp = _cfg_from_yaml(pkg['xp.cfg']) if isinstance(pkg['xp.cfg'], str) else pkg['xp.cfg']
        sea, rvq = xp['seanet'], xp['rvq']
        cfg = dict(channels=xp['channels'], dimension=sea['dimension'], n_filters=sea['n_filters'],
                   n_residual_layers=sea['n_residual_layers'], ratios=list(sea['ratios']),
                   kernel_size=sea['kernel_size'], last_kernel_size=sea['last_kernel_size'],
                   residual_kernel_size=sea['residual_kernel_size'], dilation_base=sea['dilation_base'],
                   causal=xp['encodec']['causal'], pad_mode=sea['pad_mode'], compress=sea['compress'],
                   lstm=sea['lstm'], norm=sea['norm'], trim_right_ratio=sea.get('trim_right_ratio', 1.0),
                   sample_rate=xp['sample_rate'], n_q=rvq['n_q'], bins=rvq['bins'],
                   renormalize=xp['encodec']['renormalize'])
        return EncodecModel(pkg['best_state'], cfg, device)
    raise FileNotFoundError(f"compression model '{name}': not a local checkpoint and not 'synthetic/<arch>' "
                            "(the HF hub is unreachable from this image)")


def synthetic_text_encoder(cfg: dict, t_text: int = 16, seed: int = 0):
    """Deterministic stand-in for the frozen T5 encoder (weights unavailable offline): hidden states are a seeded
    function of the prompt string, lengths follow the word count."""
    def encode(entries: tp.List[str]):
        hs, ms = [], []
        for e in entries:
            g = torch.Generator()
            g.manual_seed((hash_str(e) + seed) % (2 ** 31))
            hs.append(torch.randn((t_text, cfg['cond_dim']), generator=g))
            n = max(1, min(t_text, len(e.split()) + 1))
            ms.append((torch.arange(t_text) < n).long())
        return torch.stack(hs), torch.stack(ms)
    return encode


def hash_str(s: str) -> int:
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xffffffff
    return h


def load_lm_model(name: str, device='cuda', seed: int = 0, text_encoder=None) -> LMModel:
    if name.startswith('synthetic/'):
        arch = name.split('/', 1)[1]
        cfg = synth.lm_config(_SCALES.get(arch, arch))
        sd = synth.synth_lm_state_dict(cfg, seed, device=device, dtype=torch.float16)
        enc = text_encoder or synthetic_text_encoder(cfg)
        provider = ConditioningProvider({'description': PrecomputedTextConditioner(cfg['cond_dim'], cfg['dim'], enc)})
        return LMModel(sd, cfg, provider, ConditionFuser({'cross': ['description']}), device)
    if os.path.isfile(name):
        pkg = torch.load(name, map_location='cpu', weights_only=False)
        xp = _cfg_from_yaml(pkg['xp.cfg']) if isinstance(pkg['xp.cfg'], str) else pkg['xp.cfg']
        tl = xp['transformer_lm']
        delays = list(xp['codebooks_pattern']['delay']['delays'])
        cfg = dict(dim=tl['dim'], num_heads=tl['num_heads'], num_layers=tl['num_layers'],
                   hidden_scale=tl.get('hidden_scale', 4), n_q=tl['n_q'], card=tl['card'], delays=delays,
                   max_period=10000.0, positional_scale=1.0, cross_attention=True,
                   cfg_coef=xp['classifier_free_guidance']['inference_coef'], cond_dim=768,
                   two_step_cfg=tl.get('two_step_cfg', False))
        if text_encoder is None:
            raise RuntimeError("a text_encoder callable (T5 hidden states) is required to drive a real checkpoint")
        provider = ConditioningProvider({'description': PrecomputedTextConditioner(768, cfg['dim'], text_encoder)})
        return LMModel(pkg['best_state'], cfg, provider, ConditionFuser({'cross': ['description']}), device)
    raise FileNotFoundError(f"LM '{name}': not a local checkpoint and not 'synthetic/<scale>'")


def load_musicgen(name: str, device=None, seed: int = 0):
    from .musicgen import MusicGen
    device = 'cuda' if device is None else device
    if name.startswith('synthetic/'):
        lm = load_lm_model(name, device, seed)
        cm = load_compression_model('synthetic/encodec_32k', device, seed + 1)
        return MusicGen(name, cm, lm, max_duration=30)
    if os.path.isdir(name):
        lm = load_lm_model(os.path.join(name, 'state_dict.bin'), device)
        cm = load_compression_model(os.path.join(name, 'compression_state_dict.bin'), device)
        return MusicGen(name, cm, lm, max_duration=30)
    raise FileNotFoundError(f"MusicGen '{name}': pass a local checkpoint directory or 'synthetic/<small|medium|large>'")


def load_audiogen(name: str, device=None, seed: int = 0):
    """`synthetic/audiogen-medium` (seeded random weights of the released architecture) or a local checkpoint directory."""
    from .musicgen import AudioGen
    device = 'cuda' if device is None else device
    if name.startswith('synthetic/'):
        scale = name.split('/', 1)[1].replace('audiogen-', '')
        lm = load_lm_model(f'synthetic/{scale}', device, seed)
        cm = load_compression_model('synthetic/encodec_16k', device, seed + 1)
        return AudioGen(name, cm, lm, max_duration=10)
    if os.path.isdir(name):
        lm = load_lm_model(os.path.join(name, 'state_dict.bin'), device)
        cm = load_compression_model(os.path.join(name, 'compression_state_dict.bin'), device)
        return AudioGen(name, cm, lm, max_duration=10)
    raise FileNotFoundError(f"AudioGen '{name}': pass a local checkpoint directory or 'synthetic/audiogen-medium'")
