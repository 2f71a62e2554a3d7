"""The reference's own implementation of the benchmarked path, timed on the box (bench.py's `--impl reference` arm,
the `cpu_baseline` field and the `reference_gpu` block).  BASELINE INFRASTRUCTURE: nothing here is product code, and none
of the audiocraft_b200 kernels or host classes are on these paths -- only `audiocraft_b200.synth` (architecture tables and
seeded weights in the reference's state_dict layout) is shared so both arms run the same model.

What runs is the UNMODIFIED reference (baseline/_ref, see baseline/install_ref.sh) through oracle/ref_import.py's
third-party stubs:
  * LM: `LMModel._sample_next_token` inside `LMModel.streaming()` -- the body of `LMModel.generate`'s loop
    (audiocraft/models/lm.py:540-565) -- for the CPU windows; the whole `LMModel.generate` for the GPU pass;
  * codec: `EncodecModel.decode` (audiocraft/models/encodec.py:236-259).
CPU (fp32, like audiocraft/models/loaders.py:115-118 picks on cpu): one bounded SAMPLE = `n_steps` decode steps with an
empty KV cache + `n_steps` with a KV cache of `ctx` positions (streaming state injected: a legitimately shaped cache
without paying a 1500-token CPU prefill) + the codec decode of 1 s of tokens; the full 30 s pass is integrated from those:
step time is linear in the cache length, so  T_pass = S * (t_0 + t_ctx) / 2 + 30 * t_dec  with S = T + n_q - 1 steps.
GPU (fp16 transformer under autocast, SDPA, eager -- the path the reference runs on CUDA, genmodel.py:74-78): the full
pass, no extrapolation.
"""
import os
import sys
import time

import torch


def _log(*a):
    print('[reference_arm]', *a, file=sys.stderr, flush=True)


def _prefer_ref():
    """bench.py must time baseline/_ref (the pip-installed copy that travels to the GPU box), not the source tree."""
    root = os.path.dirname(os.path.abspath(__file__))
    cand = os.path.join(root, '_ref')
    if os.path.isdir(os.path.join(cand, 'audiocraft', 'modules')):
        os.environ.setdefault('AUDIOCRAFT_REFERENCE', cand)


_prefer_ref()

_SCALES = {'small': 'musicgen_small', 'medium': 'musicgen_medium', 'large': 'musicgen_large'}


def available() -> bool:
    from oracle import ref_import as R
    return R.available()


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)


def _models(scale: str, batch: int, t_text: int, device, lm_dtype):
    from oracle import ref_models as RM
    from audiocraft_b200 import synth
    cfg = synth.lm_config(_SCALES.get(scale, scale))
    gen_dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    sd = synth.synth_lm_state_dict(cfg, 0, device=gen_dev, dtype=torch.float16 if gen_dev == 'cuda' else torch.float32)
    hid, mask = synth.synth_text_condition(cfg, batch, t_text, seed=0)
    lm, CA = RM.build_ref_lm(cfg, sd, RM.text_table(hid, mask, cfg['cond_dim']), device=device, dtype=lm_dtype)
    del sd
    ecfg = synth.ENCODEC_CONFIGS['encodec_32k']
    cm = RM.build_ref_encodec(ecfg, synth.synth_encodec_state_dict(ecfg, 1), device=device)
    return cfg, lm, CA, cm


class CpuReference:
    """Reference modules on the host cores; `sample()` = one bounded sample (see module docstring)."""

    def __init__(self, scale='medium', batch=8, duration=30.0, t_text=16, threads=None):
        self.avail = host_threads()
        self.threads = threads or self.avail
        torch.set_num_threads(self.threads)
        self.scale, self.batch, self.duration = scale, batch, duration
        t0 = time.perf_counter()
        self.cfg, self.lm, self.CA, self.cm = _models(scale, batch, t_text, 'cpu', torch.float32)
        _log(f'reference {scale} LM + EnCodec-32k built on cpu in {time.perf_counter() - t0:.1f} s; {self.avail} host threads available')
        from oracle import ref_models as RM
        self.RM = RM
        with torch.no_grad():
            self.cfg_conditions = RM.ref_cfg_conditions(self.lm, self.CA, batch)
        self.frame_rate = 50
        self.T = int(duration * self.frame_rate)
        self.S = self.T + self.cfg['n_q'] - 1          # decode steps of one generate() (lm.py:540)
        self.kw = dict(use_sampling=True, temp=1.0, top_k=250, top_p=0.0, cfg_coef=3.0)
        if threads is None:
            self._calibrate_threads()

    def _calibrate_threads(self):
        """"All the host threads it can use": rows = 2B GEMVs stop scaling (and oversubscribed OpenMP teams collapse) well
        below the core count of a big box, so the thread count is the fastest of a short sweep, timed on one decode step."""
        cands = sorted({n for n in (8, 16, 32, 64, 128, self.avail) if n <= self.avail})
        best, best_t = None, None
        for n in cands:
            torch.set_num_threads(n)
            t = self._window(1, 1, budget_s=20.0)
            _log(f'  {n} threads: {t * 1e3:.0f} ms per decode step')
            if best_t is None or t < best_t:
                best, best_t = n, t
            if t > 2.5 * best_t:
                break
        self.threads = best
        torch.set_num_threads(best)
        _log(f'using {best} threads')

    @torch.no_grad()
    def _window(self, ctx: int, n_steps: int, budget_s: float = 8.0) -> float:
        """seconds per decode step with `ctx` cached positions (at most n_steps steps, at least one, about budget_s seconds)"""
        lm, B, K = self.lm, self.batch, self.cfg['n_q']
        seq = torch.full((B, K, 1), lm.special_token_id, dtype=torch.long)
        with lm.streaming():
            seq = self.RM.ref_decode_steps(lm, self.cfg_conditions, seq, 1, **self.kw)   # creates the streaming state
            if ctx > 1:
                state = lm.get_streaming_state()
                g = torch.Generator().manual_seed(1)
                for k in list(state.keys()):
                    v = state[k]
                    if k.endswith('past_keys') or k.endswith('past_values'):   # [rows, H, t, 64] (transformer.py:266-298)
                        state[k] = torch.randn(v.shape[:2] + (ctx,) + v.shape[3:], generator=g, dtype=v.dtype) * 0.5
                    elif k.endswith('offsets'):
                        state[k] = torch.full_like(v, ctx)
                lm.set_streaming_state(state)
            t0 = time.perf_counter()
            done = 0
            while done < n_steps and (done == 0 or time.perf_counter() - t0 < budget_s):
                seq = self.RM.ref_decode_steps(lm, self.cfg_conditions, seq, 1, **self.kw)
                done += 1
            dt = time.perf_counter() - t0
        return dt / done

    @torch.no_grad()
    def _decode_1s(self) -> float:
        codes = torch.randint(0, 2048, (self.batch, 4, self.frame_rate))
        t0 = time.perf_counter()
        self.cm.decode(codes, None)
        return time.perf_counter() - t0

    def sample(self, n_steps=4, ctx=None):
        ctx = self.S - 1 if ctx is None else ctx
        t0 = time.perf_counter()
        s0 = self._window(1, n_steps)
        s1 = self._window(ctx, n_steps)
        sd = self._decode_1s()
        wall = time.perf_counter() - t0
        t_pass = self.S * (s0 + (s0 + (s1 - s0) * (self.S - 1) / max(1, ctx - 1))) / 2 + self.duration * sd
        _log(f'sample: {s0 * 1e3:.0f} ms/step at KV 1, {s1 * 1e3:.0f} ms/step at KV {ctx}, decode 1 s {sd * 1e3:.0f} ms -> '
             f'{self.batch * self.duration / t_pass:.4f} audio-s/s ({wall:.1f} s wall)')
        return dict(wall_s=wall, step_ctx0_s=s0, step_ctx_s=s1, ctx=ctx, decode_1s_s=sd, pass_s=t_pass,
                    value=self.batch * self.duration / t_pass, n_steps=n_steps)

    def describe(self, smp):
        return (f"reference modules (baseline/_ref) fp32 on {self.threads} of {self.avail} host threads (fastest of a sweep): <= {smp['n_steps']} decode steps at KV 1 "
                f"({smp['step_ctx0_s'] * 1e3:.0f} ms/step) + {smp['n_steps']} at KV {smp['ctx']} ({smp['step_ctx_s'] * 1e3:.0f} ms/step, "
                f"streaming state injected) + EnCodec decode of 1 s ({smp['decode_1s_s'] * 1e3:.0f} ms) of the batch={self.batch} "
                f"{self.scale} workload; integrated over {self.S} steps + {self.duration:g} s decode = {smp['pass_s']:.0f} s per pass")


def cpu_baseline(scale='medium', batch=8, duration=30.0, n_samples=1, n_steps=4, warm=True):
    """cpu_baseline object for bench.py's b200 line (median of n_samples bounded samples)."""
    from oracle import ref_import as R
    ref = CpuReference(scale, batch, duration)
    if warm:
        ref.sample(n_steps=1, ctx=64)
    smps = [ref.sample(n_steps=n_steps) for _ in range(n_samples)]
    smps.sort(key=lambda s: s['value'])
    med = smps[len(smps) // 2]
    return dict(value=round(med['value'], 4), unit='audio-s/s', cores=ref.threads, kind='reference' if R.kind() == '_ref' else 'reference-tree',
                sample=ref.describe(med), spread=[round(s['value'], 4) for s in smps])


@torch.no_grad()
def gpu_reference(scale='medium', batch=8, duration=30.0, passes=2, t_text=16, encodec_items=32):
    """The reference's CUDA path on this box: fp16 transformer under autocast, eager PyTorch + SDPA.  Full passes."""
    from oracle import ref_models as RM
    dev = torch.device('cuda', torch.cuda.current_device())
    cfg, lm, CA, cm = _models(scale, batch, t_text, dev, torch.float16)
    conds = [CA(text={'description': f'd{i}'}) for i in range(batch)]
    T = int(duration * 50)
    kw = dict(use_sampling=True, temp=1.0, top_k=250, top_p=0.0, cfg_coef=3.0)

    def one(max_len):
        with torch.autocast('cuda', dtype=torch.float16):
            tokens = lm.generate(None, conds, max_gen_len=max_len, **kw)
        return cm.decode(tokens, None)

    t0 = time.perf_counter()
    one(48)                                   # warm-up (cuBLAS / cuDNN heuristics, allocator)
    torch.cuda.synchronize()
    per_step = (time.perf_counter() - t0) / 51
    _log(f'reference {scale} on cuda: warm-up {per_step * 1e3:.1f} ms per decode step (upper bound)')
    if per_step * (T + 3) > 90:               # keep the default bench run bounded
        passes = 1
    times = []
    for _ in range(passes):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wav = one(T)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
        _log(f'full pass {times[-1]:.2f} s')
    best = min(times)
    out = dict(value=round(batch * duration / best, 3), unit='audio-s/s', pass_s=[round(t, 3) for t in times],
               what=f"reference LMModel.generate (fp16 autocast, SDPA, eager; {T + cfg['n_q'] - 1} steps, CFG rows={2 * batch}) + "
                    f"EncodecModel.decode on this GPU, best of {passes} full {duration:g} s passes, same synthetic weights and shapes",
               wav_shape=list(wav.shape))
    del lm
    # EnCodec 32 kHz encode + decode, 32 x 10 s, the reference's own modules (cuDNN convs / cuDNN LSTM, fp32 with TF32 allowed
    # as torch defaults for cuDNN convolutions)
    if encodec_items:
        x = torch.randn(encodec_items, 1, 320000, device=dev) * 0.1
        for _ in range(2):
            c, s = cm.encode(x)
            cm.decode(c, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            c, s = cm.encode(x)
            y = cm.decode(c, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out['encodec'] = dict(value=round(x.numel() / (ms / 1e3) / 1e6, 2), unit='MSamples/s', ms=round(ms, 2),
                              what=f'reference EncodecModel.encode + decode, {encodec_items} x 10 s at 32 kHz, fp32 cuDNN (TF32 convs allowed)')
        # how exact is that path?  RVQ indices of the reference on CUDA vs the reference on the host (fp32) for 1 x 2 s
        try:
            from audiocraft_b200 import synth
            torch.set_num_threads(min(16, host_threads()))
            ecfg = synth.ENCODEC_CONFIGS['encodec_32k']
            cpu_cm = RM.build_ref_encodec(ecfg, synth.synth_encodec_state_dict(ecfg, 1), device='cpu')
            xs = x[:1, :, :64000]
            c_gpu, _ = cm.encode(xs)
            c_cpu, _ = cpu_cm.encode(xs.cpu())
            out['encodec']['code_agreement_with_its_own_fp32_cpu_path'] = round(float((c_gpu.cpu() == c_cpu).float().mean()), 4)
        except Exception as ex:   # evidence only
            out['encodec']['code_agreement_with_its_own_fp32_cpu_path'] = 'unavailable: ' + type(ex).__name__
    return out


@torch.no_grad()
def cpu_encodec_baseline(items=1, seconds=10.0):
    """EnCodec-32k encode+decode of the reference modules on the host cores, MSamples/s (thread count = fastest of a short
    sweep on 1 s of audio: oversubscribed OpenMP teams collapse on a many-core box, 128 threads measured 14x slower than 16)."""
    from oracle import ref_models as RM
    from audiocraft_b200 import synth
    avail = host_threads()
    ecfg = synth.ENCODEC_CONFIGS['encodec_32k']
    cm = RM.build_ref_encodec(ecfg, synth.synth_encodec_state_dict(ecfg, 1), device='cpu')
    x = torch.randn(items, 1, int(seconds * 32000)) * 0.1

    def run(xx):
        t0 = time.perf_counter()
        c, s = cm.encode(xx)
        cm.decode(c, s)
        return time.perf_counter() - t0

    best, best_t = None, None
    for n in sorted({t for t in (8, 16, 32, 64, avail) if t <= avail}):
        torch.set_num_threads(n)
        run(x[..., :16000])
        t = run(x[..., :32000])
        _log(f'  EnCodec 1 s on {n} threads: {t * 1e3:.0f} ms')
        if best_t is None or t < best_t:
            best, best_t = n, t
        if t > 2.5 * best_t:
            break
    torch.set_num_threads(best)
    dt = run(x)
    return dict(value=round(x.numel() / dt / 1e6, 4), unit='MSamples/s', cores=best, kind='reference',
                sample=f'reference EncodecModel.encode + decode of {items} x {seconds:g} s at 32 kHz, fp32, {best} of {avail} host threads '
                       f'(fastest of a sweep), {dt:.1f} s')
