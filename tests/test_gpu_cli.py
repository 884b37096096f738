"""End-to-end GPU test of the run_basis_sep CLI mirror: outputs contract + final-SDR parity with the oracle."""
import os

import numpy as np
import pytest
import torch

from audiosourcesep_b200 import GlowConfig, synthetic
from audiosourcesep_b200.weights import init_glow_params
from oracle import basis_oracle as bo
from oracle.glow_oracle import GlowOracle

pytestmark = pytest.mark.gpu


def test_run_basis_sep_glow_cli_contract_and_sdr(tmp_path):
    from audiosourcesep_b200 import ops
    from audiosourcesep_b200.run_basis_sep import build_parser, main
    out = tmp_path / "sep"
    n_mixed, T, L = 2, 3, 3
    argv = ["unused1", "unused2", "--output", str(out), "--model_type", "glow", "--synthetic", "--random_init", "7",
            "--n_mixed", str(n_mixed), "--T", str(T), "--K", "2", "--L", "3", "--n_filters", "512", "--learntop",
            "--sigma1", "0.05", "--sigmaL", "0.01", "--num_classes", str(L), "--progression", "logarithmic", "--seed", "5"]
    res = main(build_parser().parse_args(argv))
    # ---- files and keys the reference writes (run_basis_sep.py:310, 425, 435-436)
    z = np.load(out / "results.npz")
    assert sorted(z.files) == ["gt1", "gt2", "mixed", "stft_mixture", "x1", "x2"]
    for k in ("x1", "x2", "gt1", "gt2", "mixed"):
        assert z[k].shape == (n_mixed, 96, 64) and z[k].dtype == np.float32
        assert z[k].min() >= -100.0 and z[k].max() <= 20.0
    zc = np.load(out / "results_convergence.npz")
    assert zc["x1"].shape == (L + 1, n_mixed, 96, 64, 1) and zc["x2"].shape == (L + 1, n_mixed, 96, 64, 1)
    log = (out / "out.log").read_text()
    assert log.count("Sigma = ") == L and "Duration: " in log and "Data Loaded in" in log
    assert res is not None and np.array_equal(res["x1"], z["x1"])
    # ---- the same run on the oracle with the SAME Philox draws (keyed by seed / step / stream / element)
    cfg = GlowConfig(H=96, W=64, C=1, L=3, K=2, n_filters=512, learntop=True, minval=0.0, maxval=1.0)
    o1 = GlowOracle(cfg, init_glow_params(cfg, seed=7, mode="perturbed"))
    o2 = GlowOracle(cfg, init_glow_params(cfg, seed=8, mode="perturbed"))
    sig = bo.get_sigmas(0.05, 0.01, L, "logarithmic")
    gt1, gt2 = synthetic.mel_patches_db(n_mixed, 0), synthetic.mel_patches_db(n_mixed, 1)
    mixed = synthetic.normalise(synthetic.mixture_db(gt1, gt2))
    rng = np.random.Generator(np.random.PCG64(5))
    x1 = rng.uniform(0.0, 1.0, mixed.shape).astype(np.float32)
    x2 = rng.uniform(0.0, 1.0, mixed.shape).astype(np.float32)

    def noise(i, t):
        step = i * T + t
        return tuple(ops.philox_normal(mixed.shape, seed=5, step=step, stream_id=s).cpu().numpy() for s in (1, 2))

    def score(o):
        return lambda x, i: o.grad_log_prob(x)[0].numpy().astype(np.float32)

    y1, y2, arr = bo.basis_run(mixed, x1, x2, score(o1), score(o2), sig, T, noise)
    w1, w2 = bo.post_processing(y1.squeeze(-1)), bo.post_processing(y2.squeeze(-1))
    for got, want, gt in ((z["x1"], w1, gt1), (z["x2"], w2, gt2)):
        sdr_cuda = bo.sdr_db(synthetic.normalise(gt.squeeze(-1)), synthetic.normalise(got))
        sdr_orac = bo.sdr_db(synthetic.normalise(gt.squeeze(-1)), synthetic.normalise(want))
        print(f"mel-domain SDR: cuda {sdr_cuda:.3f} dB, oracle {sdr_orac:.3f} dB")
        assert abs(sdr_cuda - sdr_orac) <= 0.1                         # north-star gate: final SDR within 0.1 dB
        rel = np.linalg.norm(synthetic.normalise(got) - synthetic.normalise(want)) / np.linalg.norm(synthetic.normalise(want))
        assert rel <= 1e-2, rel
    np.testing.assert_allclose(zc["x1"][0].squeeze(-1), bo.post_processing(x1.squeeze(-1)), atol=1e-4)


def test_run_basis_sep_ncsn_cli_runs(tmp_path):
    from audiosourcesep_b200.run_basis_sep import build_parser, main
    out = tmp_path / "sep_ncsn"
    cfg_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "melspec_ncsnv2.yml")
    argv = ["unused1", "unused2", "--output", str(out), "--model_type", "ncsn", "--synthetic", "--random_init", "3",
            "--n_mixed", "2", "--config", cfg_path]
    args = build_parser().parse_args(argv)
    from audiosourcesep_b200.run_basis_sep import merge_config
    args = merge_config(args)
    args.config = None
    args.T, args.num_classes, args.sigma1 = 1, 3, 0.05     # 3 short noise levels so the test stays small
    res = main(args)
    z = np.load(out / "results.npz")
    assert z["x1"].shape == (2, 96, 64) and np.all(np.isfinite(z["x1"])) and np.all(np.isfinite(z["x2"]))
    assert np.load(out / "results_convergence.npz")["x1"].shape == (4, 2, 96, 64, 1)
    assert res["duration_s"] > 0


def test_ncsn_generate_samples_cli(tmp_path, capsys):
    """The sampler CLI mirror (reference: ncsn_generate_samples.py:24-116): output file, shape [L+1, n, H, W, 1], value
    range of the post-processing, the reference's print-outs; the final samples equal a direct call of the sampler."""
    from audiosourcesep_b200.ncsn_generate_samples import build_parser, main
    from audiosourcesep_b200.ncsn.utils import anneal_langevin_dynamics, get_sigmas, get_uncompiled_model_v2
    out = tmp_path / "samples"
    argv = [str(tmp_path / "ckpt"), "--filename", str(out), "--n_samples", "2", "--version", "v2", "--n_filters", "128",
            "--T", "2", "--num_classes", "3", "--sigma1", "0.05", "--random_init", "3", "--seed", "4", "--fast"]
    args = build_parser().parse_args(argv)
    arr = main(args)
    text = capsys.readouterr().out
    for line in ("SAMPLING PARAMETERS", "Weights loaded", "Start Generating 2 samples....", "Done. Duration:", "Shape: (4, 2, 96, 64, 1)",
                 "Generated Samples saved at"):
        assert line in text, line
    assert text.count("Sigma = ") == 3
    saved = np.load(str(out) + ".npy")
    assert saved.shape == (4, 2, 96, 64, 1) and np.array_equal(saved, arr)
    assert saved.min() >= -100.0 and saved.max() <= 20.0 and np.all(np.isfinite(saved))
    # direct call of the sampler with the same seeds
    sig = get_sigmas(0.05, 0.01, 3)
    ns = build_parser().parse_args(argv)
    ns.data_shape = [96, 64, 1]
    model = get_uncompiled_model_v2(ns, sigmas=sig, seed=3)
    x0 = torch.rand([2, 96, 64, 1], generator=torch.Generator().manual_seed(4))
    direct = anneal_langevin_dynamics(x0, [96, 64, 1], model, 2, sig, n_steps_each=2, step_lr=2e-5, seed=4)
    want = np.clip(direct * 120.0 - 100.0, -100.0, 20.0)
    np.testing.assert_allclose(saved[-1], want, rtol=0, atol=1e-4)



def test_wav_song_directory_to_separated_wav_files(tmp_path):
    """The reference's whole user path: <song_dir>/{mix,piano,violin}.wav -> run_basis_sep (GPU mel front end,
    run_basis_sep.py:342-351) -> results.npz with the mixture's STFT -> melspec_inversion_basis (GPU back end) -> wav."""
    import wave
    from audiosourcesep_b200 import melspec_inversion_basis as inv
    from audiosourcesep_b200.run_basis_sep import build_parser, main, merge_config
    sr, L, n_win = 16000, 32640, 4
    rng = np.random.default_rng(0)
    t = np.arange(L * n_win) / sr
    piano = 0.3 * np.sin(2 * np.pi * 330 * t) * (1 + 0.5 * np.sin(2 * np.pi * 2 * t)) + 0.002 * rng.standard_normal(t.size)
    violin = 0.2 * np.sin(2 * np.pi * 880 * t + 0.3 * np.sin(2 * np.pi * 5 * t)) + 0.002 * rng.standard_normal(t.size)
    song = tmp_path / "song"
    song.mkdir()
    for name, a in (("piano", piano), ("violin", violin), ("mix", piano + violin)):
        with wave.open(str(song / f"{name}.wav"), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
            w.writeframes((np.clip(a, -1, 1 - 1 / 32768) * 32768).astype("<i2").tobytes())
    out = tmp_path / "sep"
    cfg_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "melspec_ncsnv2.yml")
    args = merge_config(build_parser().parse_args(["unused1", "unused2", "--output", str(out), "--model_type", "ncsn", "--song_dir", str(song),
                                                   "--random_init", "3", "--n_mixed", "2", "--config", cfg_path]))
    args.config = None
    args.T, args.num_classes, args.sigma1 = 1, 2, 0.05
    main(args)
    z = np.load(out / "results.npz")
    assert z["stft_mixture"].shape == (2, 1025, 64) and z["stft_mixture"].dtype == np.complex64
    assert np.abs(z["stft_mixture"]).max() > 1.0                     # the real STFT of the mixture, not the placeholder zeros
    assert z["mixed"].shape == (2, 96, 64) and z["gt1"].max() <= 20.0 and z["gt1"].min() >= -100.0
    # the mel patches are those of windows 2 and 3 of the files (the first two are skipped, data_loader.py:131-134)
    from oracle import mel_oracle as mo
    pcm = (np.clip(piano + violin, -1, 1 - 1 / 32768) * 32768).astype("<i2").astype(np.float32) / 32768.0
    want, _ = mo.melspectrogram_db(pcm[2 * L: 3 * L])
    assert np.max(np.abs(z["mixed"][0] - want)) <= 2e-3
    flat = inv.main(inv.build_parser().parse_args([str(out), "--wiener_filter"]))
    assert flat["x1_audio"].shape == (2 * 512 * 63,) and np.all(np.isfinite(flat["x1_audio"]))
    assert os.path.exists(out / "inverse_reuse_phase_frame_wiener_filter" / "sep1.wav")
    # ground-truth spectrograms pushed through the same inversion come back close to the true piano track
    n = 512 * 63
    ref = np.concatenate([pcm[2 * L: 2 * L + n] * 0 + piano[2 * L: 2 * L + n], piano[3 * L: 3 * L + n]])
    sdr = 10 * np.log10(np.sum(ref ** 2) / np.sum((flat["gt1_audio"] - ref) ** 2))
    print(f"inversion of the ground-truth piano spectrogram: SDR {sdr:.1f} dB")
    assert sdr > 8.0


def test_train_ncsn_cli_runs_and_writes_weights(tmp_path, capsys):
    """train_ncsn.main (reference: train_ncsn.py main / train): a short run of the host loop on the real handle -- noise
    level drawn per replica batch, Adam + EMA, weights.npz with every parameter of the network."""
    from audiosourcesep_b200.train_ncsn import build_parser, main
    from audiosourcesep_b200.weights import ncsn_param_shapes
    from audiosourcesep_b200 import NCSNConfig
    out = tmp_path / "trained"
    args = build_parser().parse_args(["--output", str(out), "--version", "v2", "--n_train", "8", "--height", "32", "--width", "32",
                                      "--n_filters", "128", "--sigma1", "1.0", "--num_classes", "5", "--progression", "logarithmic",
                                      "--n_epochs", "2", "--batch_size", "4", "--learning_rate", "1e-5", "--ema", "--seed", "1"])
    hist = main(args)
    assert len(hist) == 4 and np.all(np.isfinite(hist))
    text = capsys.readouterr().out
    assert "Total Trainable Variables:" in text and text.count("Train Loss:") == 2 and "Training time:" in text
    w = np.load(out / "weights.npz")
    shapes = ncsn_param_shapes(NCSNConfig(version="v2", H=32, W=32, ngf=128, num_classes=5, sigma1=1.0))
    assert sorted(w.files) == sorted(shapes) and all(tuple(w[k].shape) == tuple(shapes[k]) for k in shapes)
