"""MusicGen generation API on B200: mirror of ``audiocraft.models.genmodel.BaseGenModel`` and
``audiocraft.models.musicgen.MusicGen`` (host code, kept verbatim in behaviour: parameter plumbing, duration ->
token count, > max_duration sliding window, tokens -> audio).  Text-to-music only: melody / style conditioning
(`generate_with_chroma`, `set_style_conditioner_params`) is a conditioner front-end outside the hot path
 and raises.
"""
import typing as tp

import torch

from .conditioners import ConditioningAttributes
from .encodec import CompressionModel
from .lm import LMModel


class BaseGenModel:
    """audiocraft/models/genmodel.py:28-267."""

    def __init__(self, name: str, compression_model: CompressionModel, lm: LMModel,
                 max_duration: tp.Optional[float] = None):
        self.name = name
        self.compression_model = compression_model
        self.lm = lm
        self.cfg = None
        self.compression_model.eval()
        self.lm.eval()
        if max_duration is None:
            raise ValueError("You must provide max_duration when building directly your GenModel")
        self.max_duration: float = max_duration
        self.duration = self.max_duration
        self.extend_stride: tp.Optional[float] = None
        self.device = lm.device
        self.generation_params: dict = {}
        self._progress_callback: tp.Optional[tp.Callable[[int, int], None]] = None

    @property
    def frame_rate(self) -> float:
        return self.compression_model.frame_rate

    @property
    def sample_rate(self) -> int:
        return self.compression_model.sample_rate

    @property
    def audio_channels(self) -> int:
        return self.compression_model.channels

    def set_custom_progress_callback(self, progress_callback: tp.Optional[tp.Callable[[int, int], None]] = None):
        self._progress_callback = progress_callback

    @torch.no_grad()
    def _prepare_tokens_and_attributes(self, descriptions: tp.Sequence[tp.Optional[str]],
                                       prompt: tp.Optional[torch.Tensor]):
        attributes = [ConditioningAttributes(text={'description': description}) for description in descriptions]
        if prompt is not None:
            if descriptions is not None:
                assert len(descriptions) == len(prompt), "Prompt and nb. descriptions doesn't match"
            prompt = prompt.to(self.device)
            prompt_tokens, scale = self.compression_model.encode(prompt)
            assert scale is None
        else:
            prompt_tokens = None
        return attributes, prompt_tokens

    def generate_unconditional(self, num_samples: int, progress: bool = False, return_tokens: bool = False):
        descriptions: tp.List[tp.Optional[str]] = [None] * num_samples
        attributes, prompt_tokens = self._prepare_tokens_and_attributes(descriptions, None)
        tokens = self._generate_tokens(attributes, prompt_tokens, progress)
        if return_tokens:
            return self.generate_audio(tokens), tokens
        return self.generate_audio(tokens)

    def generate(self, descriptions: tp.List[str], progress: bool = False, return_tokens: bool = False):
        attributes, prompt_tokens = self._prepare_tokens_and_attributes(descriptions, None)
        assert prompt_tokens is None
        tokens = self._generate_tokens(attributes, prompt_tokens, progress)
        if return_tokens:
            return self.generate_audio(tokens), tokens
        return self.generate_audio(tokens)

    def generate_continuation(self, prompt: torch.Tensor, prompt_sample_rate: int,
                              descriptions: tp.Optional[tp.List[tp.Optional[str]]] = None,
                              progress: bool = False, return_tokens: bool = False):
        if prompt.dim() == 2:
            prompt = prompt[None]
        if prompt.dim() != 3:
            raise ValueError("prompt should have 3 dimensions: [B, C, T] (C = 1).")
        if prompt_sample_rate != self.sample_rate or prompt.shape[1] != self.audio_channels:
            # the reference resamples with julius here (data/audio_utils.py:54-59); host IO, out of scope (8f.2)
            raise NotImplementedError("convert_audio (resample / remix) is not built: pass the prompt at "
                                      f"{self.sample_rate} Hz with {self.audio_channels} channel(s)")
        if descriptions is None:
            descriptions = [None] * len(prompt)
        attributes, prompt_tokens = self._prepare_tokens_and_attributes(descriptions, prompt)
        assert prompt_tokens is not None
        tokens = self._generate_tokens(attributes, prompt_tokens, progress)
        if return_tokens:
            return self.generate_audio(tokens), tokens
        return self.generate_audio(tokens)

    def _generate_tokens(self, attributes: tp.List[ConditioningAttributes],
                         prompt_tokens: tp.Optional[torch.Tensor], progress: bool = False) -> torch.Tensor:
        """genmodel.py:193-260."""
        total_gen_len = int(self.duration * self.frame_rate)
        max_prompt_len = int(min(self.duration, self.max_duration) * self.frame_rate)
        current_gen_offset: int = 0

        def _progress_callback(generated_tokens: int, tokens_to_generate: int):
            generated_tokens += current_gen_offset
            if self._progress_callback is not None:
                self._progress_callback(generated_tokens, tokens_to_generate)
            else:
                print(f'{generated_tokens: 6d} / {tokens_to_generate: 6d}', end='\r')

        if prompt_tokens is not None:
            assert max_prompt_len >= prompt_tokens.shape[-1], "Prompt is longer than audio to generate"
        callback = _progress_callback if progress else None

        if self.duration <= self.max_duration:
            gen_tokens = self.lm.generate(prompt_tokens, attributes, callback=callback, max_gen_len=total_gen_len,
                                          **self.generation_params)
        else:
            assert self.extend_stride is not None, "Stride should be defined to generate beyond max_duration"
            assert self.extend_stride < self.max_duration, "Cannot stride by more than max generation duration."
            all_tokens = []
            if prompt_tokens is None:
                prompt_length = 0
            else:
                all_tokens.append(prompt_tokens)
                prompt_length = prompt_tokens.shape[-1]
            stride_tokens = int(self.frame_rate * self.extend_stride)
            while current_gen_offset + prompt_length < total_gen_len:
                time_offset = current_gen_offset / self.frame_rate
                chunk_duration = min(self.duration - time_offset, self.max_duration)
                max_gen_len = int(chunk_duration * self.frame_rate)
                gen_tokens = self.lm.generate(prompt_tokens, attributes, callback=callback, max_gen_len=max_gen_len,
                                              **self.generation_params)
                if prompt_tokens is None:
                    all_tokens.append(gen_tokens)
                else:
                    all_tokens.append(gen_tokens[:, :, prompt_tokens.shape[-1]:])
                prompt_tokens = gen_tokens[:, :, stride_tokens:]
                prompt_length = prompt_tokens.shape[-1]
                current_gen_offset += stride_tokens
            gen_tokens = torch.cat(all_tokens, dim=-1)
        return gen_tokens

    def generate_audio(self, gen_tokens: torch.Tensor) -> torch.Tensor:
        assert gen_tokens.dim() == 3
        with torch.no_grad():
            return self.compression_model.decode(gen_tokens, None)


class MusicGen(BaseGenModel):
    """audiocraft/models/musicgen.py:41-338."""

    def __init__(self, name: str, compression_model: CompressionModel, lm: LMModel,
                 max_duration: tp.Optional[float] = None):
        super().__init__(name, compression_model, lm, max_duration)
        self.set_generation_params(duration=15)  # default duration

    @staticmethod
    def get_pretrained(name: str = 'facebook/musicgen-medium', device=None):
        """The reference pulls checkpoints from the HF hub (musicgen.py:56-94); offline this accepts a directory
        holding the reference's exported `state_dict.bin` + `compression_state_dict.bin`, or `synthetic/<scale>`
        (small | medium | large) for seeded random weights of the released architectures."""
        from .loaders import load_musicgen
        return load_musicgen(name, device=device)

    def set_generation_params(self, use_sampling: bool = True, top_k: int = 250, top_p: float = 0.0,
                              temperature: float = 1.0, duration: float = 30.0, cfg_coef: float = 3.0,
                              cfg_coef_beta: tp.Optional[float] = None, two_step_cfg: bool = False,
                              extend_stride: float = 18):
        assert extend_stride < self.max_duration, "Cannot stride by more than max generation duration."
        self.extend_stride = extend_stride
        self.duration = duration
        self.generation_params = {
            'use_sampling': use_sampling,
            'temp': temperature,
            'top_k': top_k,
            'top_p': top_p,
            'cfg_coef': cfg_coef,
            'two_step_cfg': two_step_cfg,
            'cfg_coef_beta': cfg_coef_beta,
        }

    def set_style_conditioner_params(self, *args, **kwargs):
        raise NotImplementedError("MusicGen-Style conditioning is not built on the B200 path (SURVEY.md 8f.3)")

    def generate_with_chroma(self, *args, **kwargs):
        raise NotImplementedError("melody (chroma) conditioning is not built on the B200 path (SURVEY.md 8f.3)")


class AudioGen(BaseGenModel):
    """audiocraft/models/audiogen.py:23-93: the same LM decode + EnCodec decode path at 16 kHz (codec
    `encodec_large_nq4_s320`, 4 codebooks, delays [0,1,2,3]); only the defaults differ from MusicGen."""

    def __init__(self, name: str, compression_model: CompressionModel, lm: LMModel,
                 max_duration: tp.Optional[float] = None):
        super().__init__(name, compression_model, lm, max_duration)
        self.set_generation_params(duration=5)  # default duration

    @staticmethod
    def get_pretrained(name: str = 'facebook/audiogen-medium', device=None):
        from .loaders import load_audiogen
        return load_audiogen(name, device=device)

    def set_generation_params(self, use_sampling: bool = True, top_k: int = 250, top_p: float = 0.0,
                              temperature: float = 1.0, duration: float = 10.0, cfg_coef: float = 3.0,
                              two_step_cfg: bool = False, extend_stride: float = 2):
        assert extend_stride < self.max_duration, "Cannot stride by more than max generation duration."
        self.extend_stride = extend_stride
        self.duration = duration
        self.generation_params = {
            'use_sampling': use_sampling,
            'temp': temperature,
            'top_k': top_k,
            'top_p': top_p,
            'cfg_coef': cfg_coef,
            'two_step_cfg': two_step_cfg,
        }
