"""CPU ORACLE (test infrastructure, NOT a product path) -- mel-spectrogram front end and inversion back end.

numpy restatement of what the reference computes through librosa (not installed here; pinned in prose only,
README.md:7-14 -- the 0.7 / 0.8 series) at
  datasets/data_loader.py:113-180        get_song_extract: librosa.stft(n_fft=2048, hop_length=512, window='hann',
                                         center=True, pad_mode='reflect') -> librosa.feature.melspectrogram(S=|stft|^2,
                                         sr=16000, fmin=125, fmax=7600, n_mels=96, power=2) -> librosa.power_to_db -> clip
  melspec_inversion_basis.py:42-119      stft_inversion: librosa.db_to_power -> librosa.feature.inverse.mel_to_stft
                                         (non-negative least squares against the mel basis, then ^(1/power)) -> mixture
                                         phase re-use or single_channel_wiener_filter -> librosa.istft(hop_length=512)
librosa semantics restated from its published algorithm: periodic Hann window (scipy get_window fftbins=True), reflect
padding by n_fft/2, rfft per frame in float64 cast to complex64, Slaney mel scale + Slaney area normalisation
(filters.mel htk=False, norm='slaney'), power_to_db(ref=1, amin=1e-10, top_db=80), istft = windowed overlap-add of irfft
frames divided by the window sum-of-squares, trimmed by n_fft/2.
Deviation, stated: librosa solves the NNLS with scipy's L-BFGS-B from the clipped least-squares solution; the system is
under-determined (96 equations, 1025 unknowns per frame), so the minimiser is not unique and depends on the optimiser's
path.  ``nnls_pg`` below is the algorithm the CUDA kernel runs (same initial point, FISTA projected gradient, fixed
iteration count) -- the device path is checked against THIS, and against the residual it reaches, not against L-BFGS-B.
Pins: librosa is absent and the reference tree holds no STFT / mel golden vector, so the exact outputs are "parity unpinned";
what IS available is checked (tests/test_oracle_mel.py, tests/golden/real_inversion.npz, cut from
basis_sep_results/beethoven_sonata_1_sep_1min): re-analysing the reference's own inverted mixture audio reproduces the mel
spectrograms it was inverted from (median 2.2 dB, mean -1.5 dB over the top 40 dB -- the inconsistency of an inverted STFT;
a wrong mel scale / normalisation / dB reference would be 10+ dB off), and this module's inversion of the same spectrogram
with the same phase agrees with the reference's librosa output to 10.6-11.8 dB SDR (the L-BFGS-B vs FISTA difference).
"""
from __future__ import annotations

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y: np.ndarray, n_fft: int = 2048, hop: int = 512) -> np.ndarray:
    """[L] -> complex64 [1 + n_fft/2, 1 + L // hop]."""
    y = np.asarray(y)
    w = hann_periodic(n_fft)
    yp = np.pad(y.astype(np.float64), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    frames = np.stack([yp[t * hop: t * hop + n_fft] for t in range(n_frames)], axis=1)
    return np.fft.rfft(frames * w[:, None], axis=0).astype(np.complex64)


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filters(sr=16000, n_fft=2048, n_mels=96, fmin=125.0, fmax=7600.0) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> float32 [n_mels, 1 + n_fft/2]."""
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def power_to_db(S, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    return np.maximum(log_spec, log_spec.max() - top_db)


def db_to_power(S_db):
    return np.power(10.0, 0.1 * np.asarray(S_db))


def melspectrogram_db(y, sr=16000, n_fft=2048, hop=512, n_mels=96, fmin=125.0, fmax=7600.0, dbmin=-100.0, dbmax=20.0):
    """One extract: (mel dB [n_mels, T] float32, stft complex64 [F, T]) -- data_loader.py:144-162."""
    S = stft(y, n_fft, hop)
    mel = np.dot(mel_filters(sr, n_fft, n_mels, fmin, fmax), (np.abs(S) ** 2).astype(np.float32))
    return np.clip(power_to_db(mel), dbmin, dbmax).astype(np.float32), S


def nnls_pg(A: np.ndarray, B: np.ndarray, iters: int = 300) -> np.ndarray:
    """min 1/2 ||A X - B||^2, X >= 0: X0 = clip(pinv(A) B, 0), then FISTA with step 1 / lambda_max(A^T A)."""
    A = A.astype(np.float64)
    B = B.astype(np.float64)
    X = np.clip(np.linalg.pinv(A) @ B, 0, None)
    step = 1.0 / np.linalg.norm(A, 2) ** 2
    Y, Xp = X.copy(), X.copy()
    for k in range(1, iters + 1):
        G = A.T @ (A @ Y - B)
        X = np.maximum(0.0, Y - step * G)
        Y = X + ((k - 1.0) / (k + 2.0)) * (X - Xp)
        Xp = X
    return X


def mel_to_stft(M_power, sr=16000, n_fft=2048, fmin=125.0, fmax=7600.0, power=2.0, iters=300):
    A = mel_filters(sr, n_fft, M_power.shape[0], fmin, fmax)
    return np.power(nnls_pg(A, M_power, iters), 1.0 / power)


def istft(S: np.ndarray, hop: int = 512) -> np.ndarray:
    n_fft = 2 * (S.shape[0] - 1)
    T = S.shape[1]
    w = hann_periodic(n_fft)
    y = np.zeros(n_fft + hop * (T - 1))
    wss = np.zeros_like(y)
    frames = np.fft.irfft(S.astype(np.complex128), n=n_fft, axis=0)
    for t in range(T):
        y[t * hop: t * hop + n_fft] += w * frames[:, t]
        wss[t * hop: t * hop + n_fft] += w ** 2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: len(y) - n_fft // 2].astype(np.float32)


def single_channel_wiener_filter(psd_sources, stft_mixture):
    """melspec_inversion_basis.py:93-119."""
    psd_sources = np.asarray(psd_sources)
    return (psd_sources / (np.sum(psd_sources, axis=0) + 1e-10)) * stft_mixture


def stft_inversion(melspecs_db, stft_mixture, wiener_filter=False, iters=300, hop=512):
    """melspec_inversion_basis.py:42-90 for one extract: list of dB mel spectrograms [n_mels, T] -> list of waveforms."""
    mags = np.array([mel_to_stft(db_to_power(m), n_fft=2 * (stft_mixture.shape[0] - 1), iters=iters) for m in melspecs_db])
    use_w = wiener_filter and len(melspecs_db) > 1
    if use_w:
        cs = single_channel_wiener_filter(mags ** 2, stft_mixture)
    else:
        cs = mags * np.exp(1j * np.angle(stft_mixture))[None]
    return [istft(c, hop) for c in cs]


def griffinlim(mag: np.ndarray, phase0: np.ndarray, n_iter: int = 32, hop: int = 512, momentum: float = 0.99) -> np.ndarray:
    """librosa.griffinlim restated (fast Griffin-Lim, Perraudin et al.): ``phase0`` replaces librosa's random initial phases."""
    n_fft = 2 * (mag.shape[0] - 1)
    angles = np.exp(1j * phase0)
    rebuilt = np.zeros_like(angles)
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft((mag * angles).astype(np.complex64), hop)
        rebuilt = stft(inverse, n_fft, hop).astype(np.complex128)
        angles = rebuilt - (momentum / (1 + momentum)) * tprev
        angles = angles / (np.abs(angles) + 1e-16)
    return istft((mag * angles).astype(np.complex64), hop)
