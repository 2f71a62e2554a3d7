"""The reference's own hot-path tests pin PROPERTIES, not numbers.  This file re-runs those
properties on CPU against the oracle (the checker the GPU parity tests trust) and against the host-side mirrors, so a
regression in either shows up without a GPU.  Each test names the reference test it follows."""
import math

import numpy as np
import pytest
import torch

from audiocraft_b200 import synth
from audiocraft_b200.encodec import conv_geometry, convtr_geometry
from audiocraft_b200.patterns import DelayedPatternProvider
from oracle import encodec_oracle as EO
from oracle import lm_oracle as LO
from tests import helpers as H


# ---------------------------------------------------------------------------------------------- transformer / streaming
@pytest.mark.parametrize('with_cross', [False, True])
def test_streaming_equals_batch(with_cross):
    """tests/modules/test_transformer.py:39-49, 133-161: feeding the sequence step by step through the KV cache equals
    one causal pass (||delta|| / ||y|| < 1e-6 in fp32) -- the contract the on-device decode step is built on."""
    cfg = synth.lm_config('lm_tiny')
    sd = synth.synth_lm_state_dict(cfg, seed=2)
    B, S = 2, 11
    tokens = torch.randint(0, cfg['card'], (B, cfg['n_q'], S), generator=torch.Generator().manual_seed(0))
    cross = None
    if with_cross:
        _, _, c = H.lm_condition(cfg, sd, B, 5, 1)
        cross = c[:B]
    o = LO.LMOracle(sd, cfg)
    whole = o.forward(tokens, cross)                                   # [B, K, S, card], one multi-token causal call
    o.reset()
    steps = torch.cat([o.forward(tokens[:, :, s:s + 1], cross) for s in range(S)], dim=2)
    assert steps.shape == whole.shape
    assert (steps - whole).norm() / whole.norm() < 1e-6
    # a prompt of 4 tokens in one call followed by single steps (LMModel.generate's first iteration, lm.py:540-548)
    o.reset()
    mixed = torch.cat([o.forward(tokens[:, :, :4], cross)] + [o.forward(tokens[:, :, s:s + 1], cross) for s in range(4, S)], dim=2)
    assert (mixed - whole).norm() / whole.norm() < 1e-6


def test_causality_of_the_oracle():
    """tests/modules/test_transformer.py:22-37 (there via gradients): logits at step s do not depend on tokens > s."""
    cfg = synth.lm_config('lm_tiny')
    sd = synth.synth_lm_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(1)
    a = torch.randint(0, cfg['card'], (1, cfg['n_q'], 9), generator=g)
    b = a.clone()
    b[:, :, 6:] = torch.randint(0, cfg['card'], (1, cfg['n_q'], 3), generator=g)
    o = LO.LMOracle(sd, cfg)
    ya = o.forward(a, None)
    o.reset()
    yb = o.forward(b, None)
    assert torch.equal(ya[:, :, :6], yb[:, :, :6]) and not torch.equal(ya[:, :, 6:], yb[:, :, 6:])


# ---------------------------------------------------------------------------------------------- delay patterns
def _slow_delay_layout(timesteps, n_q, delays):
    """The layout DelayedPatternProvider.get_pattern describes (codebooks_patterns.py:339-356): step 0 is empty, then
    code (t, q) sits at sequence step t + 1 + delays[q]; written as plain loops like the reference test's slow
    implementation (tests/modules/test_codebooks_patterns.py:107-150)."""
    S = timesteps + max(delays) + 1
    layout = [[None] * S for _ in range(n_q)]
    for q in range(n_q):
        for t in range(timesteps):
            layout[q][t + 1 + delays[q]] = t
    return layout


@pytest.mark.parametrize('n_q,timesteps,delays', [(4, 10, [0, 1, 2, 3]), (4, 1, [0, 1, 2, 3]), (2, 7, [0, 3]),
                                                   (8, 12, [0, 0, 1, 1, 2, 2, 3, 3]), (3, 5, [0, 1, 1])])
def test_delay_pattern_build_and_revert_against_slow_reference(n_q, timesteps, delays):
    """tests/modules/test_codebooks_patterns.py:107-246: build / revert against a slow loop implementation, bit exact."""
    special = 999
    B = 2
    z = torch.arange(B * n_q * timesteps).view(B, n_q, timesteps) % 97
    layout = _slow_delay_layout(timesteps, n_q, delays)
    S = len(layout[0])
    want = torch.full((B, n_q, S), special, dtype=z.dtype)
    want_mask = torch.zeros(n_q, S, dtype=torch.bool)
    for q in range(n_q):
        for s in range(S):
            if layout[q][s] is not None:
                want[:, q, s] = z[:, q, layout[q][s]]
                want_mask[q, s] = True
    pattern = DelayedPatternProvider(n_q, delays=delays).get_pattern(timesteps)
    assert pattern.num_sequence_steps == timesteps + max(delays)          # layout length without the special step 0
    values, _, mask = pattern.build_pattern_sequence(z, special)
    assert torch.equal(values, want) and torch.equal(mask, want_mask)
    assert torch.equal(LO.build_delay_sequence(z, delays, special)[0], want)
    back, _, back_mask = pattern.revert_pattern_sequence(values, special)
    assert torch.equal(back, z) and bool(back_mask.all())
    assert torch.equal(LO.revert_delay_sequence(values, delays, special, timesteps)[0], z)
    # a sequence cut short (generation in progress): the missing codes come back as the special token
    cut = values[:, :, :S - 2]
    back_cut, _, m_cut = pattern.revert_pattern_sequence(cut, special)
    for q in range(n_q):
        for t in range(timesteps):
            present = t + 1 + delays[q] < S - 2
            assert bool(m_cut[q, t]) == present
            assert all(int(back_cut[b, q, t]) == (int(z[b, q, t]) if present else special) for b in range(B))
    for t in range(timesteps):
        assert pattern.get_first_step_with_timesteps(t) == t + 1 + min(delays)


# ---------------------------------------------------------------------------------------------- conv / SEANet shapes
@pytest.mark.parametrize('k,s,d', [(4, 1, 1), (4, 2, 1), (3, 1, 3), (10, 5, 1), (3, 2, 3)])
@pytest.mark.parametrize('causal', [False, True])
def test_streamable_conv_lengths(k, s, d, causal):
    """tests/modules/test_conv.py:160-173: output length ceil(L / stride) for every (kernel, stride, dilation), causal
    or not -- for the oracle's conv and for the geometry the kernels are launched with."""
    for L in (1, 7, 50, 51, 203):
        x = torch.randn(1, 2, L)
        w, b = torch.randn(3, 2, k), torch.randn(3)
        # (inputs shorter than the reflect padding go through pad1d's zero-extension rule, conv.py:71-88)
        y = EO.sconv1d(x, w, b, stride=s, dilation=d, causal=causal, pad_mode='reflect')
        assert y.shape == (1, 3, math.ceil(L / s)), (L, y.shape)
        assert conv_geometry(L, k, s, d, causal, True)[2] == math.ceil(L / s)


@pytest.mark.parametrize('causal,ratio', [(False, 1.0), (True, 1.0), (True, 0.5), (True, 0.0)])
def test_streamable_convtr_lengths(causal, ratio):
    """tests/modules/test_conv.py:182-203: a transposed conv with K = 2 * stride gives L * stride steps after trimming,
    whatever the trim ratio."""
    for s in (2, 4, 5, 8):
        for L in (1, 3, 50):
            x = torch.randn(1, 3, L)
            w, b = torch.randn(3, 2, 2 * s), torch.randn(2)
            y = EO.sconvtr1d(x, w, b, stride=s, causal=causal, trim_right_ratio=ratio)
            assert y.shape == (1, 2, L * s)
            assert convtr_geometry(L, 2 * s, s, causal, ratio)[1] == L * s


@pytest.mark.parametrize('name,frames', [('encodec_24k', 75), ('encodec_tiny', None), ('encodec_tiny_causal', None)])
def test_seanet_shapes_and_length_round_trip(name, frames):
    """tests/modules/test_seanet.py:18-56 (24000 samples -> [1, 128, 75] -> 24000) and
    tests/models/test_encodec_model.py:37-46 (decode(encode(x)) covers x for random lengths)."""
    cfg = synth.ENCODEC_CONFIGS[name]
    sd = synth.synth_encodec_state_dict(cfg, seed=1)
    o = EO.EncodecOracle(sd, cfg)
    hop = int(np.prod(cfg['ratios']))
    if frames is not None:
        x = H.audio_input(cfg, 1, 24000, 0)
        z = o.encode_latent(x)
        assert z.shape == (1, cfg['dimension'], frames)
        codes, _ = o.encode(x)
        assert o.decode(codes).shape == (1, cfg['channels'], 24000)
    else:
        g = torch.Generator().manual_seed(4)
        for _ in range(6):
            L = int(torch.randint(hop, 3000, (1,), generator=g))
            x = H.audio_input(cfg, 2, L, L)
            codes, scale = o.encode(x)
            assert codes.shape == (2, cfg['n_q'], math.ceil(L / hop)) and codes.dtype == torch.int64
            assert (scale is not None) == bool(cfg.get('renormalize'))
            y = o.decode(codes, scale)
            assert y.shape[:2] == x.shape[:2] and y.shape[-1] == math.ceil(L / hop) * hop >= L
