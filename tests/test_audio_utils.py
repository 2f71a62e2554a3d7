"""convert_audio (host-side prompt preprocessing, audiocraft/data/audio_utils.py:20-59).  The resampler restates
julius.resample_frac, which is absent here: parity is unpinned, so these tests hold it to the properties its
construction guarantees."""
import math

import pytest
import torch

from audiocraft_b200.audio_utils import convert_audio, convert_audio_channels, resample_frac


@pytest.mark.parametrize('old,new', [(44100, 32000), (16000, 32000), (48000, 32000), (24000, 32000), (32000, 16000)])
def test_resample_length_constant_and_sine(old, new):
    n = old // 2                                    # half a second
    t_old = torch.arange(n, dtype=torch.float64) / old
    f0 = 440.0
    x = torch.sin(2 * math.pi * f0 * t_old).float()[None, None]
    y = resample_frac(x, old, new)
    assert y.shape == (1, 1, math.floor(new * n / old))
    t_new = torch.arange(y.shape[-1], dtype=torch.float64) / new
    ref = torch.sin(2 * math.pi * f0 * t_new).float()
    edge = new // 50                                # replicate padding distorts the first / last ~20 ms
    err = (y[0, 0, edge:-edge] - ref[edge:-edge]).abs().max()
    assert err < 2e-3, err
    c = resample_frac(torch.full((2, 1, n), 0.37), old, new)
    assert (c - 0.37).abs().max() < 1e-5            # unit-sum phase kernels + replicate padding keep a constant exactly


def test_resample_identity_linearity_and_batch_shape():
    x = torch.randn(3, 2, 1000)
    assert resample_frac(x, 32000, 32000) is x
    assert torch.equal(resample_frac(x, 48000 * 3, 48000 * 3), x)
    a, b = torch.randn(1, 1, 4410), torch.randn(1, 1, 4410)
    lhs = resample_frac(2.0 * a - 0.5 * b, 44100, 32000)
    rhs = 2.0 * resample_frac(a, 44100, 32000) - 0.5 * resample_frac(b, 44100, 32000)
    torch.testing.assert_close(lhs, rhs, rtol=0, atol=1e-5)
    y = resample_frac(x, 44100, 32000)              # leading dimensions are kept, rows are independent
    assert y.shape[:2] == (3, 2)
    torch.testing.assert_close(y[1, 0], resample_frac(x[1, 0][None], 44100, 32000)[0], rtol=0, atol=1e-6)
    with pytest.raises(ValueError):
        resample_frac(x, 0, 32000)


def test_resample_rejects_above_nyquist_content():
    old, new, n = 48000, 16000, 48000
    t = torch.arange(n, dtype=torch.float64) / old
    hi = torch.sin(2 * math.pi * 11000.0 * t).float()[None, None]     # above the 8 kHz Nyquist of the target
    y = resample_frac(hi, old, new)
    assert y[..., 400:-400].abs().max() < 2e-2


def test_convert_audio_channels_rules():
    x = torch.randn(2, 2, 50)
    assert convert_audio_channels(x, 2) is x
    torch.testing.assert_close(convert_audio_channels(x, 1), x.mean(dim=1, keepdim=True))
    m = torch.randn(2, 1, 50)
    up = convert_audio_channels(m, 2)
    assert up.shape == (2, 2, 50) and torch.equal(up[:, 0], up[:, 1])
    six = torch.randn(1, 6, 10)
    assert torch.equal(convert_audio_channels(six, 2), six[:, :2])
    with pytest.raises(ValueError):
        convert_audio_channels(torch.randn(1, 2, 10), 3)
    y = convert_audio(torch.randn(1, 2, 44100), 44100, 32000, 1)
    assert y.shape == (1, 1, 32000)
