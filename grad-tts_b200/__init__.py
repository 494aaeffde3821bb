"""B200-native Grad-TTS hot path: reverse-diffusion mel decoder + Monotonic Alignment Search.

Drop-in Python surface (same names and signatures as the reference's `model/` package):
    GradTTS.forward(x, x_lengths, n_timesteps, temperature, stoc, spk, length_scale)
    Diffusion.forward / reverse_diffusion(z, mask, mu, n_timesteps, stoc, spk)
    GradLogPEstimator2d.forward(x, mask, mu, t, spk)
    monotonic_align.maximum_path(value, mask)
    TextEncoder.forward(x, x_lengths, spk)            (model/text_encoder.py)
    hifigan.Generator(h).forward(mel)                 (hifi-gan/models.py)
    inference.synthesize(generator, vocoder, ...)     (the loop body of inference.py)
All arithmetic of those calls runs in hand-written sm_100a kernels behind the C ABI declared in
include/gradtts_b200.h (libgradtts_b200.so); this package is only the host-side mirror.
"""
from . import _lib, synth  # noqa: F401
from .model import GradTTS  # noqa: F401
from .model.diffusion import Diffusion, GradLogPEstimator2d  # noqa: F401
from .model.monotonic_align import maximum_path  # noqa: F401
from . import dist  # noqa: F401
from . import likelihood  # noqa: F401
from . import hifigan  # noqa: F401
from . import inference  # noqa: F401
from .model.text_encoder import TextEncoder  # noqa: F401

__all__ = ["GradTTS", "Diffusion", "GradLogPEstimator2d", "maximum_path", "synth", "dist", "likelihood", "hifigan", "inference", "TextEncoder"]
