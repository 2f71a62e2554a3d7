"""Generate tests/golden/*.pt by running the REAL reference modules (read-only tree at /root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
The reference's own tests hold no numeric golden vectors for this path, so these
fixtures are what pins the oracle (tests/test_oracle_vs_golden.py) and, on the GPU box where /root/reference
does not exist, the CUDA path (tests/test_gpu_*.py).  Inputs and weights are regenerated from seeds by
tests/helpers.py + audiocraft_b200/synth.py; only outputs are stored.
"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_import as R  # noqa: E402
from tests import helpers as H  # noqa: E402
from audiocraft_b200 import synth  # noqa: E402

warnings.filterwarnings('ignore')
torch.set_grad_enabled(False)


def build_ref_encodec(cfg, sd):
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
    return m.eval()


def build_ref_lm(cfg, sd, table):
    lmm, cond, pat = R.mod('models.lm'), R.mod('modules.conditioners'), R.mod('modules.codebooks_patterns')

    class StubText(cond.TextConditioner):
        """Stands in for T5Conditioner (no T5 weights offline): same contract, hidden states from a table."""
        def __init__(self, dim, output_dim):
            super().__init__(dim, output_dim)

        def tokenize(self, x):
            hs, ms = zip(*[table['__null__'] if xi is None else table[xi] for xi in x])
            return {'hid': torch.stack(hs), 'mask': torch.stack(ms)}

        def forward(self, inputs):
            mask = inputs['mask']
            return self.output_proj(inputs['hid']) * mask.unsqueeze(-1), mask

    prov = cond.ConditioningProvider({'description': StubText(cfg['cond_dim'], cfg['dim'])})
    fuser = cond.ConditionFuser({'cross': ['description'], 'sum': [], 'prepend': [], 'input_interpolate': []})
    m = lmm.LMModel(pat.DelayedPatternProvider(cfg['n_q'], delays=cfg['delays']), prov, fuser, n_q=cfg['n_q'],
                    card=cfg['card'], dim=cfg['dim'], num_heads=cfg['num_heads'], hidden_scale=cfg['hidden_scale'],
                    norm='layer_norm', norm_first=True, bias_proj=False, cfg_coef=cfg['cfg_coef'],
                    num_layers=cfg['num_layers'], bias_ff=False, bias_attn=False, causal=True, memory_efficient=True,
                    cross_attention=True, activation='gelu', positional_embedding='sin', dropout=0.0)
    m.load_state_dict(sd, strict=True)
    return m.eval(), cond.ConditioningAttributes


def golden_encodec(name, batch, length, wseed, xseed, full):
    cfg = synth.ENCODEC_CONFIGS[name]
    sd = synth.synth_encodec_state_dict(cfg, seed=wseed)
    m = build_ref_encodec(cfg, sd)
    x = H.audio_input(cfg, batch, length, xseed)
    codes, scale = m.encode(x)
    xin = m.preprocess(x)[0]
    latent = m.encoder(xin)
    wav = m.decode(codes, scale)
    qlat = m.decode_latent(codes)
    out = dict(name=name, batch=batch, length=length, wseed=wseed, xseed=xseed, codes=codes, scale=scale,
               x_head=x[..., :64].clone())
    if full:
        out.update(latent=latent, wav=wav, qlat=qlat)
    else:  # keep the fixture small: full codes, strided views of the float tensors
        out.update(latent=latent, wav_head=wav[..., :512].clone(), wav_stride=97,
                   wav_strided=wav[..., ::97].clone(), wav_len=wav.shape[-1])
    torch.save(out, os.path.join(H.GOLDEN_DIR, f'{name}.pt'))
    print(name, 'codes', tuple(codes.shape), 'wav', tuple(wav.shape))


def golden_lm(name, batch, t_text, T, wseed, cseed, steps_logits, topn=None):
    cfg = synth.lm_config(name)
    sd = synth.synth_lm_state_dict(cfg, seed=wseed)
    hid, mask, _ = H.lm_condition(cfg, sd, batch, t_text, cseed)
    table = {f'd{i}': (hid[i], mask[i]) for i in range(batch)}
    table['__null__'] = (torch.zeros(t_text, cfg['cond_dim']), torch.zeros(t_text, dtype=torch.long))
    m, CA = build_ref_lm(cfg, sd, table)
    conds = [CA(text={'description': f'd{i}'}) for i in range(batch)]
    out = dict(name=name, batch=batch, t_text=t_text, T=T, wseed=wseed, cseed=cseed)
    greedy = m.generate(None, conds, max_gen_len=T, use_sampling=False)
    out['greedy'] = greedy
    # teacher-forced logits along the greedy path: re-run with a hook on _sample_next_token's CFG-mixed logits
    logits = []
    orig = m._sample_next_token

    def spy(sequence, cfg_conditions, unconditional_state, use_sampling=False, temp=1.0, top_k=0, top_p=0.0,
            cfg_coef=None, cfg_coef_beta=None, two_step_cfg=None):
        B = sequence.shape[0]
        state = {k: v.clone() for k, v in m.get_streaming_state().items()}
        al = m(torch.cat([sequence, sequence], 0), conditions=[], condition_tensors=cfg_conditions)
        c, u = al.split(B, dim=0)
        coef = m.cfg_coef if cfg_coef is None else cfg_coef
        logits.append((u + (c - u) * coef)[:, :, -1, :].clone())
        m.set_streaming_state(state)
        return orig(sequence, cfg_conditions, unconditional_state, use_sampling, temp, top_k, top_p,
                    cfg_coef=cfg_coef, cfg_coef_beta=cfg_coef_beta, two_step_cfg=two_step_cfg)

    m._sample_next_token = spy
    greedy2 = m.generate(None, conds, max_gen_len=T, use_sampling=False)
    m._sample_next_token = orig
    assert (greedy2 == greedy).all()
    lg = torch.stack(logits[:steps_logits], 0)  # [steps,B,K,card]
    if topn is None:
        out['logits'] = lg
    else:
        tv, ti = lg.topk(topn, dim=-1)
        out['logits_top_v'], out['logits_top_i'] = tv, ti
    if topn is None:
        # sampled generations with torch's own generator (what the reference does, utils/utils.py:103)
        torch.manual_seed(11)
        out['sampled_topk'] = m.generate(None, conds, max_gen_len=T, use_sampling=True, top_k=10, temp=0.9)
        torch.manual_seed(12)
        out['sampled_topp'] = m.generate(None, conds, max_gen_len=T, use_sampling=True, top_k=0, top_p=0.8)
        torch.manual_seed(13)
        out['sampled_plain'] = m.generate(None, conds, max_gen_len=T, use_sampling=True, top_k=0, top_p=0.0, temp=1.3)
        out['continuation'] = m.generate(greedy[..., :5].clone(), conds, max_gen_len=T, use_sampling=False)
        # generate_unconditional: descriptions are None -> null condition rows, still doubled by CFG (genmodel.py:135-149)
        out['unconditional'] = m.generate(None, [CA(text={'description': None}) for _ in range(batch)],
                                          max_gen_len=T, use_sampling=False)
    torch.save(out, os.path.join(H.GOLDEN_DIR, f'{name}.pt'))
    print(name, 'greedy', tuple(greedy.shape))


def golden_patterns():
    pat = R.mod('modules.codebooks_patterns')
    out = {}
    for (K, T, delays) in [(4, 7, [0, 1, 2, 3]), (4, 1, [0, 1, 2, 3]), (8, 12, [0, 1, 2, 3, 4, 5, 6, 7]),
                           (3, 5, [0, 0, 2])]:
        p = pat.DelayedPatternProvider(K, delays=delays).get_pattern(T)
        g = torch.Generator()
        g.manual_seed(K * 100 + T)
        codes = torch.randint(0, 50, (2, K, T), generator=g)
        seq, idx, mask = p.build_pattern_sequence(codes, 99)
        back, idx2, mask2 = p.revert_pattern_sequence(seq, special_token=-1)
        out[(K, T, tuple(delays))] = dict(codes=codes, seq=seq, mask=mask, back=back, back_mask=mask2,
                                          first_step_T0=[p.get_first_step_with_timesteps(t) for t in range(T)])
    torch.save(out, os.path.join(H.GOLDEN_DIR, 'patterns.pt'))
    print('patterns', len(out))


def golden_sampling():
    utils = R.mod('utils.utils')
    g = torch.Generator()
    g.manual_seed(21)
    logits = torch.randn(6, 4, 2048, generator=g) * 2.0
    probs = torch.softmax(logits, -1)
    out = dict(seed=21)
    torch.manual_seed(31)
    out['top_k_250'] = utils.sample_top_k(probs.clone(), 250)
    torch.manual_seed(32)
    out['top_p_0.9'] = utils.sample_top_p(probs.clone(), 0.9)
    torch.manual_seed(33)
    out['plain'] = utils.multinomial(probs.clone(), 1)
    torch.save(out, os.path.join(H.GOLDEN_DIR, 'sampling.pt'))
    print('sampling ok')


def golden_stereo():
    """InterleaveStereoCompressionModel over the tiny mono codec (renormalize off), both interleavings."""
    enc = R.mod('models.encodec')
    cfg = dict(synth.ENCODEC_CONFIGS['encodec_tiny'])
    cfg['renormalize'] = False
    m = build_ref_encodec(cfg, synth.synth_encodec_state_dict(cfg, seed=1))
    cfg2 = dict(cfg)
    cfg2['channels'] = 2
    x = H.audio_input(cfg2, 2, 900, 9)
    out = dict(wseed=1, xseed=9, length=900, batch=2)
    for pt in (False, True):
        w = enc.InterleaveStereoCompressionModel(m, per_timestep=pt)
        codes, _ = w.encode(x)
        out[f'codes_pt{int(pt)}'] = codes
        out[f'wav_pt{int(pt)}'] = w.decode(codes, None)
        out[f'props_pt{int(pt)}'] = (w.num_codebooks, w.frame_rate, w.channels, w.cardinality, w.total_codebooks)
    torch.save(out, os.path.join(H.GOLDEN_DIR, 'encodec_tiny_stereo.pt'))


if __name__ == '__main__':
    os.makedirs(H.GOLDEN_DIR, exist_ok=True)
    golden_patterns()
    golden_sampling()
    golden_encodec('encodec_tiny', 2, 1234, 1, 2, full=True)
    golden_encodec('encodec_tiny_causal', 2, 777, 3, 4, full=True)
    golden_encodec('encodec_24k', 1, 24000, 5, 6, full=False)
    golden_encodec('encodec_32k', 1, 32000, 7, 8, full=False)
    golden_lm('lm_mini', 2, 5, 12, 3, 1, steps_logits=15)
    golden_lm('lm_tiny', 3, 4, 9, 4, 2, steps_logits=12)
    golden_lm('musicgen_small', 1, 6, 3, 9, 5, steps_logits=6, topn=32)
    golden_stereo()
