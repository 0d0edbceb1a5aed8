"""Synthetic 48 kHz PCM in the shape BASELINE.json / SURVEY.md 8(d) prescribe.

Stream s (seed 1234+s): a harmonic source (f0 ~ U[80,400] Hz, slow vibrato, 20 harmonics
with 1/h roll-off) plus white noise at an SNR ~ U[0,20] dB, peak-normalised to 0.25 at the
"/32768" scale the reference CLI feeds (/root/reference/src/main.cpp:34).  ``scale=32768``
gives the int16-scale floats that the reference's ``train()`` feeds the same C API with
(/root/reference/src/denoise.cpp:697-698) and that make the comb-filter branch execute
(SURVEY.md 0.6).  numpy only -- used by tests and as the host-side generator of bench.py.
"""
from __future__ import annotations

import numpy as np

from .weights import uniform_pm

FRAME = 480
SR = 48000.0


def _u01(n, seed, tag):
    return (uniform_pm((n,), seed, tag, 1.0).astype(np.float64) + 1.0) * 0.5


def _voice_and_noise(sd: int, T: int):
    """One stream's harmonic 'speech' component and its noise component (already at the stream's SNR)."""
    t = np.arange(T, dtype=np.float64) / SR
    p = _u01(8, sd, 1)
    f0 = 80.0 + 320.0 * p[0]
    vib_rate = 3.0 + 4.0 * p[1]
    vib_depth = 0.01 + 0.03 * p[2]
    snr_db = 20.0 * p[3]
    phase = 2 * np.pi * (f0 * t - f0 * vib_depth / (2 * np.pi * vib_rate) * np.cos(2 * np.pi * vib_rate * t))
    sig = np.zeros(T)
    for h in range(1, 21):
        if h * f0 * (1 + vib_depth) < 0.45 * SR:
            sig += np.sin(h * phase + 2 * np.pi * p[4] * h) / h
    # amplitude envelope: syllable-like on/off so that silence and onsets occur
    env = 0.5 * (1 + np.sin(2 * np.pi * (1.5 + 2 * p[5]) * t + 2 * np.pi * p[6]))
    sig *= env ** 2
    # gaussian noise from two uniform draws (Box-Muller)
    u1 = np.maximum(_u01(T, sd, 2), 2.0 ** -24)
    u2 = _u01(T, sd, 3)
    noise = np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)
    ps, pn = np.mean(sig ** 2) + 1e-12, np.mean(noise ** 2)
    return sig, noise * np.sqrt(ps / pn / (10 ** (snr_db / 10)))


def synth_pcm(n_streams: int, n_frames: int, seed: int = 1234, scale: float = 1.0,
              first_stream: int = 0) -> np.ndarray:
    """float32 [n_streams, n_frames*480]; stream k is independent of n_streams."""
    T = n_frames * FRAME
    out = np.empty((n_streams, T), np.float32)
    for k in range(n_streams):
        sig, noise = _voice_and_noise(seed + first_stream + k, T)
        mix = sig + noise
        mix *= 0.25 / (np.max(np.abs(mix)) + 1e-12)
        out[k] = (mix * scale).astype(np.float32)
    return out


def synth_pairs(n_pairs: int, n_frames: int, seed: int = 1234, first_pair: int = 0):
    """(clean, noisy) int16 [n_pairs, n_frames*480] each: what the reference's training-data generator reads
    as its <speech> and <noisy> files (denoise.cpp:600-690; the noisy file is the finished mixture)."""
    T = n_frames * FRAME
    clean = np.empty((n_pairs, T), np.int16)
    noisy = np.empty((n_pairs, T), np.int16)
    for k in range(n_pairs):
        sig, noise = _voice_and_noise(seed + first_pair + k, T)
        a = 0.25 / (np.max(np.abs(sig + noise)) + 1e-12)
        clean[k] = to_int16(sig * a)
        noisy[k] = to_int16((sig + noise) * a)
    return clean, noisy


def to_int16(x: np.ndarray) -> np.ndarray:
    """float (+-1 scale) -> int16 the way a WAV writer would (round, clip)."""
    return np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)
