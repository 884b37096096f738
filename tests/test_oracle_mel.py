"""CPU tests of the mel front / back end oracle (oracle/mel_oracle.py) and of the host-side filter bank."""
import numpy as np

from oracle import mel_oracle as mo


def test_mel_filter_bank_known_properties():
    from audiosourcesep_b200 import melspec
    A = mo.mel_filters()
    assert A.shape == (96, 1025) and A.dtype == np.float32
    assert np.array_equal(A, melspec.mel_filters())               # product and oracle restate the same published formula
    assert (A > 0).sum(axis=0).max() <= 2                         # triangular filters overlap pairwise only
    # Slaney normalisation: every filter has unit area in Hz (up to the FFT-bin discretisation)
    df = 16000.0 / 2048
    area = A.sum(axis=1) * df
    assert np.all(np.abs(area - 1.0) < 0.35) and abs(float(np.mean(area)) - 1.0) < 0.02
    # band edges: nothing below fmin = 125 Hz or above fmax = 7600 Hz
    freqs = np.linspace(0, 8000, 1025)
    assert A[:, freqs < 125].sum() == 0 and A[:, freqs > 7600].sum() == 0
    # Slaney scale: linear below 1 kHz (200/3 Hz per mel), 1 kHz = 15 mel
    assert abs(float(mo.hz_to_mel(1000.0)) - 15.0) < 1e-12 and abs(float(mo.mel_to_hz(mo.hz_to_mel(4321.0))) - 4321.0) < 1e-9


def test_stft_round_trip_and_power_to_db():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(32640).astype(np.float32) * 0.1
    S = mo.stft(y)
    assert S.shape == (1025, 64) and S.dtype == np.complex64
    back = mo.istft(S)
    assert back.shape == (512 * 63,) and np.max(np.abs(back - y[: back.shape[0]])) < 1e-6
    db = mo.power_to_db(np.array([[1.0, 1e-3], [1e-12, 1e-20]]))
    assert np.allclose(db, [[0.0, -30.0], [-80.0, -80.0]])        # amin = 1e-10 -> -100 dB, floored at max - 80
    mel, _ = mo.melspectrogram_db(y)
    assert mel.shape == (96, 64) and mel.max() <= 20.0 and mel.min() >= -100.0


def test_nnls_restatement_reaches_the_residual_floor():
    rng = np.random.default_rng(1)
    A = mo.mel_filters().astype(np.float64)
    X = rng.random((1025, 4)) ** 4
    B = A @ X
    Xh = mo.nnls_pg(A, B, iters=300)
    assert Xh.min() >= 0.0 and np.linalg.norm(A @ Xh - B) <= 1e-3 * np.linalg.norm(B)


def _real():
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_inversion.npz"))
    return d["mix_audio_int16"].astype(np.float32) / 32768.0, d["mixed_db"]


def test_front_end_conventions_against_the_reference_shipped_artefacts():
    """Real data the reference ships: its inverted mixture audio (librosa back end) and the mel spectrograms it was
    inverted from (librosa front end on the recording).  Re-analysing the audio must reproduce the spectrogram up to the
    inconsistency of an inverted STFT: a wrong mel scale, filter normalisation, window scaling or dB reference would show as a
    systematic offset of 10 dB or more (Slaney-normalised vs un-normalised filters alone differ by ~25 dB)."""
    audio, mixed = _real()
    for y, ref in zip(audio, mixed):
        db, _ = mo.melspectrogram_db(y)
        hi = ref > ref.max() - 40.0
        delta = (db - ref)[hi]
        assert np.median(np.abs(delta)) <= 3.0 and abs(float(np.mean(delta))) <= 3.0, (np.median(np.abs(delta)), np.mean(delta))


def test_back_end_against_the_reference_inverted_audio():
    """Same mel spectrogram, same phase (that of the reference's own output), FISTA instead of librosa's L-BFGS-B for the
    non-negative least squares: the waveform agrees with the reference's to ~11 dB SDR -- the size of the solver
    difference on an under-determined system, stated in INTEGRATION.md."""
    audio, mixed = _real()
    y, ref = audio[0], mixed[0]
    mine = mo.stft_inversion([ref], mo.stft(y), wiener_filter=False, iters=300)[0]
    sdr = 10 * np.log10(np.sum(y ** 2) / np.sum((mine - y) ** 2))
    assert sdr >= 8.0, sdr


def test_griffin_lim_restatement_converges():
    rng = np.random.default_rng(2)
    t = np.arange(32640) / 16000.0
    y = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 1320 * t)).astype(np.float32)
    m = np.abs(mo.stft(y)).astype(np.float64)
    phase0 = rng.uniform(0, 2 * np.pi, m.shape)
    sc = lambda a: np.linalg.norm(np.abs(mo.stft(a)) - m) / np.linalg.norm(m)
    start = sc(mo.istft((m * np.exp(1j * phase0)).astype(np.complex64)))
    end = sc(mo.griffinlim(m, phase0, n_iter=16))
    assert end < 0.3 * start
