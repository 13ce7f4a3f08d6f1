"""Generates the golden vectors in tests/golden/ from the reference build
(oracle/_ref, i.e. the unmodified /root/reference sources + FFTW shim).

    make -C oracle && python tests/golden/make_golden.py

Inputs are stored next to the outputs so the fixtures do not depend on the
synthetic generator staying bit-stable.  Oracle compiler/flags: see
oracle/_ref/BUILD_FLAGS.txt (gcc 13.3, -O3 -ffast-math -mavx2 -mfma).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from libspecbleach_b200.synth import nonstationary_noise, speech_plus_noise  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = Path(__file__).resolve().parent

P2D = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0, whitening_factor=50.0,
           nlm_masking_protection=0.5, suppression_strength=50.0, aggressiveness=0.0,
           tonal_reduction=6.0)
P1D = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=30.0, whitening_factor=50.0,
           masking_depth=0.5, suppression_strength=50.0, aggressiveness=1.0, tonal_reduction=0.0)

CASES = {
    # name: (two_d, sr, frame_ms, seconds, learn_samples, params, chunk, generator, seed)
    "g2d_48k_manual": (True, 48000, 50.0, 1.6, 24000, P2D, 4096, "speech", 42),
    "g1d_44k_demo": (False, 44100, 50.0, 1.5, 8 * 512, P1D, 512, "speech", 43),
    "g2d_44k_spp": (True, 44100, 50.0, 1.5, 0,
                    dict(P2D, adaptive_noise=1, noise_estimation_method=0), 3000, "noise", 44),
    "g1d_48k_martin": (False, 48000, 20.0, 1.0, 0,
                       dict(P1D, adaptive_noise=1, noise_estimation_method=2), 1000, "noise", 45),
    "g2d_48k_brandt": (True, 48000, 50.0, 1.5, 0,
                       dict(P2D, adaptive_noise=1, noise_estimation_method=1), 5000, "noise", 46),
    "g1d_44k_brandt_learned": (False, 44100, 20.0, 1.6, 8820,
                               dict(P1D, adaptive_noise=1, noise_estimation_method=1), 2048, "speech", 47),
}


def main(only=None):
    for name, (two_d, sr, ms, secs, learn, prm, chunk, gen, seed) in CASES.items():
        if only and name not in only:
            continue
        x = speech_plus_noise(secs, sr, seed) if gen == "speech" else nonstationary_noise(secs, sr, seed)
        d = rb.RefDenoiser(sr, ms, two_d)
        ys = []
        if learn:
            d.load_parameters(learn_noise=1)
            ys.append(d.process(x[:learn], chunk))
        profiles = np.stack([d.get_profile(m) for m in (1, 2, 3, 4)])
        blocks = np.array([d.block_count(m) for m in (1, 2, 3, 4)])
        d.load_parameters(**prm)
        ys.append(d.process(x[learn:], chunk))
        y = np.concatenate(ys)
        np.savez_compressed(OUT / f"{name}.npz", x=x, y=y, profiles_after_learn=profiles,
                            blocks=blocks, latency=d.latency(), profile_size=d.profile_size())
        print(name, x.shape, float(np.sqrt(np.mean(y ** 2))))
    if only:
        return
    # NLM module on its own (nlm_filter.c), incl. a partial last block (203 % 8 = 3)
    rng = np.random.default_rng(7)
    s = rng.exponential(1.0, (27, 203)).astype(np.float32)
    s[:, 50:80] = 5.0 + 0.01 * rng.standard_normal((27, 30)).astype(np.float32)
    s[:, 150:160] = 0.0
    out = rb.ref_nlm(s, 1.0)
    np.savez_compressed(OUT / "nlm_module.npz", snr=s, out=out[20:], h=1.0)
    print("nlm_module", out[20:].shape)


if __name__ == "__main__":
    main(sys.argv[1:])
