"""``get_song_extract`` -- mirror of the reference's ``datasets/data_loader.py:113-180`` (+ ``load_wav``,
datasets/preprocessing.py:9-26): cut a mixture and its two sources into 2.04 s extracts, STFT them, map to 96 mel bands in
dB.  The wav files are read with the standard library (PCM, down-mixed to mono like librosa.load(mono=True)); the STFT, the
mel projection and the dB conversion run on the GPU (audiosourcesep_b200.melspec -> csrc/mel_kernels.cu).
No resampling: the file's rate must equal ``sr`` (the reference relies on librosa.load to resample)."""
from __future__ import annotations

import wave

import numpy as np
import torch

from .. import melspec


def read_wav_mono(path: str):
    """PCM wav -> (float32 mono in [-1, 1), rate)."""
    with wave.open(path, "rb") as w:
        rate, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    if nch > 1:
        a = a.reshape(-1, nch).mean(axis=1)
    return a, rate


def load_wav(path, length_sec, sr=None):
    """reference: datasets/preprocessing.py:9-26 -- the song cut into windows of int(rate * length_sec) samples
    (drop_remainder=True).  Returns (float32 [n_windows, LENGTH], rate)."""
    song, rate = read_wav_mono(path)
    if sr is not None and rate != sr:
        raise ValueError(f"{path}: sampling rate {rate}, expected {sr} (no resampler here; convert the file first)")
    LENGTH = int(rate * length_sec)
    n = len(song) // LENGTH
    return song[: n * LENGTH].reshape(n, LENGTH), rate


def get_song_extract(mix_path, piano_path, violin_path, duration, **kwargs):
    """reference: datasets/data_loader.py:113-180.  Returns (mel_spec, raw_audio, stft_mixture):
    mel_spec = [mix, piano, violin], each float32 [n_extract, n_mels, T, 1]; raw_audio = the three concatenated
    waveforms; stft_mixture complex64 [n_extract, 1 + n_fft/2, T]."""
    length_sec = kwargs["length_sec"]
    fmin, fmax, sr = kwargs["fmin"], kwargs["fmax"], kwargs["sr"]
    dbmin, dbmax = kwargs["dbmin"], kwargs["dbmax"]
    n_fft, hop_length, n_mels = kwargs["n_fft"], kwargs["hop_length"], kwargs["n_mels"]
    use_dB = kwargs["use_dB"]
    if not use_dB:
        raise NotImplementedError("power-scale mel spectrograms (use_dB=False) are not on the separation path of the configs")
    n_extract = int(round(duration / length_sec, 0))
    tracks = []
    for p in (mix_path, piano_path, violin_path):
        frames, _ = load_wav(p, length_sec, sr=sr)
        if frames.shape[0] < 2 + n_extract:
            raise ValueError(f"{p}: {frames.shape[0]} windows of {length_sec} s, {2 + n_extract} needed (the first two are skipped)")
        tracks.append(frames[2: 2 + n_extract])                       # "skip 2 first frames" (:131-134)
    raw_audio = [np.concatenate(list(t)) for t in tracks]
    mel_spec, stft_mixture = [], None
    for i, t in enumerate(tracks):
        S = melspec.stft(t, n_fft=n_fft, hop_length=hop_length)
        if i == 0:
            stft_mixture = S.cpu().numpy()
        mel = melspec.melspectrogram_db(S, sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax, dbmin=dbmin, dbmax=dbmax)
        mel_spec.append(mel.unsqueeze(-1).contiguous())
    return mel_spec, raw_audio, stft_mixture
