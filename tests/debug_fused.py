"""Bring-up of the persistent fused decode step (csrc/lm_step.cu), phase by phase.  Not a pytest file.

  python tests/debug_fused.py phases [arch]   stop the kernel after n = 1, 2, ... grid barriers (ACB_LM_STEP_STOP) and check what
                                              phase n wrote against a torch restatement of that phase applied to the buffers
                                              phase n-1 left behind (errors do not accumulate: the first bad phase is named)
  python tests/debug_fused.py e2e [arch]      teacher-forced logits, fused vs the per-phase kernels (ACB_LM_STEP=v5) vs the oracle
  python tests/debug_fused.py gen [arch]      greedy generation, fused vs per-phase
"""
import ctypes as C
import faulthandler
import os
import sys
import time

faulthandler.enable()
os.environ['ACB_LM_STEP'] = 'fused'   # opt-in path: must be set before the model is built (weights are packed at load)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from audiocraft_b200 import _lib, synth  # noqa: E402
from audiocraft_b200.lm import LMModel  # noqa: E402
from tests import helpers as H  # noqa: E402


def log(*a):
    print(f'[{time.time() % 1000:7.2f}]', *a, flush=True)


mode = sys.argv[1] if len(sys.argv) > 1 else 'phases'
arch = sys.argv[2] if len(sys.argv) > 2 else 'lm_mini'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = synth.lm_config(arch)
sd = synth.synth_lm_state_dict(cfg, seed=3)
T_TEXT, T = 5, 10
_, _, cross = H.lm_condition(cfg, sd, B, T_TEXT, 1)
m = LMModel(sd, cfg, None, None)
torch.cuda.synchronize()
log(f'{arch}: model built, fused_ok={m.fused_ok}')
d, Hh, L, ffn, card, K = m.dim, m.num_heads, m.num_layers, m.ffn_dim, m.card, m.n_q
rows = 2 * B
g = torch.Generator().manual_seed(5)
seq = torch.randint(0, card, (B, K, T + 4), generator=g)
seq[:, :, 0] = card


def plan():
    out = (C.c_int * 32)()
    _lib.check(m._lib.acb_lm_debug_step_plan(m._handle, out))
    v = list(out)
    names = ['qkv', 'o', 'cq', 'co', 'ff1', 'ff2', 'heads']
    return {n: dict(N=v[4 * i], K=v[4 * i + 1], ks=v[4 * i + 2], kb=v[4 * i + 3]) for i, n in enumerate(names)}, v[28:32]


def run(stop, n_steps=1):
    if stop is None:
        os.environ.pop('ACB_LM_STEP_STOP', None)
    else:
        os.environ['ACB_LM_STEP_STOP'] = str(stop)
    os.environ['ACB_LM_STEP'] = 'fused'
    out = m.teacher_forced_logits(seq, cross, 3.0, n_steps=n_steps)
    torch.cuda.synchronize()
    return out


def snap():
    b = m._bufs
    return {k: b[k].clone() for k in ('x', 'part', 'stats', 'a16', 'f16', 'logits', 'k_cache', 'v_cache', 'ck_cache', 'cv_cache')}


def psum(s, name, pl, width):
    """sum of the split-K partials of GEMM `name`: part is [slots][R][ldmax] flat per GEMM as [ks][R][N]"""
    ks, R = pl[name]['ks'], RPAD[0]
    flat = s['part'].reshape(-1)[:ks * R * width].reshape(ks, R, width)
    return flat[:, :rows].sum(0)


RPAD = [16]


def stats_view(s):
    """stats is [8][R][2] with the kernel's R (padded live rows), whatever the buffer was allocated for"""
    R = RPAD[0]
    return s['stats'].reshape(-1)[:8 * R * 2].reshape(8, R, 2)


def ln(x, gb):
    return F.layer_norm(x, (d,), gb[0], gb[1], 1e-5).half()


def report(what, got, want, tol):
    err = (got.float() - want.float()).abs().max().item()
    scale = want.float().abs().max().item()
    ok = err <= tol * max(1.0, scale)
    log(f'  {"ok " if ok else "BAD"} {what}: max err {err:.3e} (|ref| max {scale:.3e})')
    return ok


if mode == 'phases':
    run(1)
    pl, misc = plan()
    log('plan', pl, 'stages/R/phases/smem', misc)
    RPAD[0] = misc[1]
    w = m._w
    pos = 0
    prev = None
    n_check = 1 + 12 * min(L, 2)   # embed + two layers
    all_ok = True
    for n in range(1, n_check + 1):
        run(n)
        cur = snap()
        k = (n - 2) % 12 if n >= 2 else -1
        l = (n - 2) // 12 if n >= 2 else 0
        lnw = w['ln'][l] if l < L else None
        if n == 1:
            tok = seq[:, :, pos].cuda()
            x = sum(w['emb'][kk][tok[:, kk]].float() for kk in range(K))
            half = d // 2
            ph = pos / w['inv_freq']
            x = x + torch.cat([torch.cos(ph), torch.sin(ph)])[None] * cfg['positional_scale']
            x = torch.cat([x, x], 0)[:rows]
            ok = report('embed x', cur['x'][:rows], x, 1e-5)
            xc = cur['x'][:rows].reshape(rows, 8, d // 8)
            ok &= report('embed stats mean', stats_view(cur)[:, :rows, 0].t(), xc.mean(-1), 1e-5)
            ok &= report('embed stats M2', stats_view(cur)[:, :rows, 1].t(), ((xc - xc.mean(-1, keepdim=True)) ** 2).sum(-1), 1e-4)
        elif k == 0:
            want = ln(prev['x'][:rows], lnw[0:2]).float() @ w['w_qkv'][l].float().t()
            ok = report(f'L{l} qkv gemm', psum(cur, 'qkv', pl, 3 * d), want, 2e-3)
        elif k == 1:
            qkv = psum(prev, 'qkv', pl, 3 * d)
            v = qkv[:, 2 * d:].half()
            ok = report(f'L{l} self-attn out (pos 0: = v)', cur['a16'][:rows], v, 1e-3)
            kk = qkv[:, d:2 * d].half().reshape(rows, Hh, 64)
            ok &= report(f'L{l} k cache', cur['k_cache'][l, :rows, :, pos], kk, 0)
            ok &= report(f'L{l} v cache', cur['v_cache'][l, :rows, :, pos], v.reshape(rows, Hh, 64), 0)
        elif k == 2:
            want = prev['a16'][:rows].float() @ w['w_o'][l].float().t()
            ok = report(f'L{l} o gemm', psum(cur, 'o', pl, d), want, 2e-3)
        elif k in (3, 7, 11):
            name = {3: 'o', 7: 'co', 11: 'ff2'}[k]
            want = prev['x'][:rows] + psum(prev, name, pl, d)
            ok = report(f'L{l} residual after {name}', cur['x'][:rows], want, 1e-5)
            xc = cur['x'][:rows].reshape(rows, 8, d // 8)
            ok &= report(f'L{l} stats mean', stats_view(cur)[:, :rows, 0].t(), xc.mean(-1), 1e-5)
            ok &= report(f'L{l} stats M2', stats_view(cur)[:, :rows, 1].t(), ((xc - xc.mean(-1, keepdim=True)) ** 2).sum(-1), 1e-4)
        elif k == 4:
            want = ln(prev['x'][:rows], lnw[2:4]).float() @ w['w_cq'][l].float().t()
            ok = report(f'L{l} cq gemm', psum(cur, 'cq', pl, d), want, 2e-3)
        elif k == 5:
            q = psum(prev, 'cq', pl, d).half().float().reshape(rows, Hh, 1, 64) / 8.0
            kc = cur['ck_cache'][l, :rows, :, :T_TEXT].float()
            vc = cur['cv_cache'][l, :rows, :, :T_TEXT].float()
            att = torch.softmax(q @ kc.transpose(-1, -2), -1) @ vc
            ok = report(f'L{l} cross-attn', cur['a16'][:rows], att.reshape(rows, d).half(), 2e-3)
        elif k == 6:
            want = prev['a16'][:rows].float() @ w['w_co'][l].float().t()
            ok = report(f'L{l} co gemm', psum(cur, 'co', pl, d), want, 2e-3)
        elif k == 8:
            want = ln(prev['x'][:rows], lnw[4:6]).float() @ w['w_ff1'][l].float().t()
            ok = report(f'L{l} ff1 gemm', psum(cur, 'ff1', pl, ffn), want, 2e-3)
        elif k == 9:
            want = F.gelu(psum(prev, 'ff1', pl, ffn).half().float()).half()
            ok = report(f'L{l} gelu', cur['f16'][:rows], want, 1e-3)
        elif k == 10:
            want = prev['f16'][:rows].float() @ w['w_ff2'][l].float().t()
            ok = report(f'L{l} ff2 gemm', psum(cur, 'ff2', pl, d), want, 2e-3)
        all_ok &= ok
        prev = cur
    log('PHASES', 'ALL OK' if all_ok else 'FAILED')
    sys.exit(0 if all_ok else 1)

if mode == 'e2e':
    n = 8
    lf = run(None, n)
    os.environ['ACB_LM_STEP'] = 'v5'
    lv = m.teacher_forced_logits(seq, cross, 3.0, n_steps=n)
    torch.cuda.synchronize()
    os.environ['ACB_LM_STEP'] = 'fused'
    for i in range(n):
        log(f'step {i}: fused vs per-phase {(lf[i] - lv[i]).abs().max():.3e}   |logit| max {lv[i].abs().max():.2f}')
    bad = (lf - lv).abs().max().item() > 4e-2
    log('E2E', 'FAILED' if bad else 'OK')
    sys.exit(1 if bad else 0)

if mode == 'gen':
    os.environ['ACB_LM_STEP'] = 'fused'
    a = m.generate(None, [], num_samples=B, max_gen_len=T, use_sampling=False, cross_attention_src=cross)
    torch.cuda.synchronize()
    log('fused launches/step', m.launches_per_step)
    os.environ['ACB_LM_STEP'] = 'v5'
    b = m.generate(None, [], num_samples=B, max_gen_len=T, use_sampling=False, cross_attention_src=cross)
    torch.cuda.synchronize()
    log('per-phase launches/step', m.launches_per_step)
    same = (a == b).float().mean().item()
    log('greedy token agreement fused vs per-phase', same, a[0, 0].tolist())
    sys.exit(0 if same > 0.95 else 1)
