"""Instantiate the REAL reference modules (oracle/ref_import.py) at this repo's architecture tables and load the seeded
synthetic state dicts into them.  TEST / BASELINE INFRASTRUCTURE ONLY.

Nothing here re-implements the reference: the constructors below are the reference's own classes
(audiocraft/models/encodec.py:125-183 EncodecModel, audiocraft/modules/seanet.py:63,156, audiocraft/quantization/vq.py:16,
audiocraft/models/lm.py:96-175 LMModel) called with the hyper-parameters of its config tree (restated in
audiocraft_b200/synth.py).  On CUDA the LM is built the way audiocraft/models/loaders.py:115-118 + builders.py:136-175 do
it: dtype float16 for the transformer only (embeddings, heads and the condition provider stay fp32), and generation runs under
``torch.autocast('cuda', float16)`` like audiocraft/models/genmodel.py:74-78.
"""
import typing as tp

import torch

from . import ref_import as R
from audiocraft_b200 import synth


def build_ref_encodec(cfg: dict, sd: tp.Dict[str, torch.Tensor], device='cpu'):
    seanet, qt, enc = R.mod('modules.seanet'), R.mod('quantization'), R.mod('models.encodec')
    kw = dict(channels=cfg['channels'], dimension=cfg['dimension'], n_filters=cfg['n_filters'],
              n_residual_layers=cfg['n_residual_layers'], ratios=cfg['ratios'], norm=cfg['norm'],
              kernel_size=cfg['kernel_size'], last_kernel_size=cfg['last_kernel_size'],
              residual_kernel_size=cfg['residual_kernel_size'], dilation_base=cfg['dilation_base'],
              causal=cfg['causal'], pad_mode=cfg['pad_mode'], compress=cfg['compress'], lstm=cfg['lstm'])
    m = enc.EncodecModel(seanet.SEANetEncoder(**kw), seanet.SEANetDecoder(**kw, trim_right_ratio=cfg['trim_right_ratio']),
                         qt.ResidualVectorQuantizer(dimension=cfg['dimension'], n_q=cfg['n_q'], bins=cfg['bins'],
                                                    kmeans_init=False),
                         frame_rate=cfg['sample_rate'] // synth.encodec_hop(cfg), sample_rate=cfg['sample_rate'],
                         channels=cfg['channels'], causal=cfg['causal'], renormalize=cfg['renormalize'])
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def build_ref_lm(cfg: dict, sd: tp.Dict[str, torch.Tensor], table: tp.Dict[str, tp.Tuple[torch.Tensor, torch.Tensor]],
                 device='cpu', dtype=torch.float32, **lm_kwargs):
    """Reference LMModel + a stub TextConditioner that stands in for T5Conditioner (no T5 weights offline): same
    contract (audiocraft/modules/conditioners.py:345-380, 509-515), hidden states looked up in ``table``
    (``table['__null__']`` for a dropped description).  Returns (model, ConditioningAttributes)."""
    lmm, cond, pat = R.mod('models.lm'), R.mod('modules.conditioners'), R.mod('modules.codebooks_patterns')

    class StubText(cond.TextConditioner):
        def __init__(self, dim, output_dim):
            super().__init__(dim, output_dim)

        def tokenize(self, x):
            hs, ms = zip(*[table['__null__'] if xi is None else table[xi] for xi in x])
            dev = self.output_proj.weight.device
            return {'hid': torch.stack(hs).to(dev), 'mask': torch.stack(ms).to(dev)}

        def forward(self, inputs):
            mask = inputs['mask']
            return self.output_proj(inputs['hid']) * mask.unsqueeze(-1), mask

    fuser = cond.ConditionFuser({'cross': ['description'], 'sum': [], 'prepend': [], 'input_interpolate': []})
    kw = dict(n_q=cfg['n_q'], card=cfg['card'], dim=cfg['dim'], num_heads=cfg['num_heads'],
              hidden_scale=cfg['hidden_scale'], norm='layer_norm', norm_first=True, bias_proj=False,
              cfg_coef=cfg['cfg_coef'], num_layers=cfg['num_layers'], bias_ff=False, bias_attn=False, causal=True,
              memory_efficient=True, cross_attention=True, activation='gelu', positional_embedding='sin', dropout=0.0)
    kw.update(lm_kwargs)
    # builders.py:136-175: device / dtype reach the transformer only; embeddings, heads and the condition provider stay
    # fp32 and the whole model is moved with .to(device).  The modules are CONSTRUCTED on the meta device (no random init
    # of 1.8-3.3 B parameters just to overwrite them) and the seeded state dict is assigned in with the dtypes the
    # reference would hold: `dtype` for transformer.* matrices and norms, fp32 for everything else.
    with torch.device('meta'):
        prov = cond.ConditioningProvider({'description': StubText(cfg['cond_dim'], cfg['dim'])}, device=device)
        m = lmm.LMModel(pat.DelayedPatternProvider(cfg['n_q'], delays=cfg['delays']), prov, fuser, dtype=dtype, **kw)
    want = {k: v.dtype for k, v in m.state_dict().items()}
    cast = {}
    for k, v in sd.items():
        assert k in want, k
        cast[k] = v.detach().to(device=device, dtype=want[k])
    missing, unexpected = m.load_state_dict(cast, strict=False, assign=True)
    assert not unexpected and all(k.endswith('rope.frequencies') for k in missing), (missing, unexpected)
    for mod in m.modules():                      # buffers created on meta (rope frequencies): recompute like rope.py:68-69
        for name, buf in list(mod.named_buffers(recurse=False)):
            if buf.is_meta:
                if name == 'frequencies':
                    hd = cfg['dim'] // cfg['num_heads']
                    adim = torch.arange(0, hd, 2, device=device, dtype=torch.float32)[: hd // 2]
                    mod.register_buffer(name, 1.0 / (float(kw.get('max_period', 10000.0)) ** (adim / hd)))
                else:
                    raise RuntimeError(f'unmaterialised buffer {name}')
    m.condition_provider.device = device
    return m.eval(), cond.ConditioningAttributes


def text_table(hid: torch.Tensor, mask: torch.Tensor, cond_dim: int):
    """description key -> (hidden [T,cond_dim], mask [T]); '__null__' = what a dropped description tokenizes to."""
    t_text = hid.shape[1]
    table = {f'd{i}': (hid[i], mask[i]) for i in range(hid.shape[0])}
    table['__null__'] = (torch.zeros(t_text, cond_dim, dtype=hid.dtype), torch.zeros(t_text, dtype=mask.dtype))
    return table


@torch.no_grad()
def ref_cfg_conditions(m, CA, n: int):
    """What LMModel.generate prepares for batched CFG (audiocraft/models/lm.py:488-511): conditions + nullified
    conditions -> tokenize -> condition provider -> {'description': ([2n,T,d], mask)}."""
    cond = R.mod('modules.conditioners')
    conds = [CA(text={'description': f'd{i}'}) for i in range(n)]
    null = cond.ClassifierFreeGuidanceDropout(p=1.0)(conds)
    tokenized = m.condition_provider.tokenize(conds + null)
    return m.condition_provider(tokenized)


@torch.no_grad()
def ref_decode_steps(m, cfg_conditions, seq: torch.Tensor, n_steps: int, **sample_kw):
    """Run ``n_steps`` iterations of the body of LMModel.generate's loop (audiocraft/models/lm.py:540-565: one
    _sample_next_token call per step, inside the streaming context the caller holds).  ``seq`` [B,K,1] is the current
    delay-pattern column; returns the last sampled column."""
    for _ in range(n_steps):
        nxt = m._sample_next_token(seq, cfg_conditions, {}, **sample_kw)
        seq = nxt
    return seq
