"""Helper of test_long_adaptive_config4_matches_reference_build: ONE 10-minute run of the
reference build (oracle/_ref, 4 OpenMP threads) in a process of its own -- libgomp is not
fork-safe once the parent has used it, so the test starts these with subprocess, three at a time.

    python tests/ref_long_run.py <learn_samples> <perturb_seed> <out.npy>
"""
import os
import sys
from pathlib import Path

os.environ["OMP_NUM_THREADS"] = "4"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

from common import P2D, run_flow  # noqa: E402
from libspecbleach_b200.synth import nonstationary_noise  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402


def main():
    learn, seed, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    sr = 44100
    x = nonstationary_noise(600.0, sr, 20260004)
    if seed:
        rng = np.random.default_rng(seed)
        x = (x * (1.0 + 6e-8 * rng.standard_normal(x.size))).astype(np.float32)
    prm = dict(P2D, adaptive_noise=1, noise_estimation_method=0)
    y = run_flow(rb.RefDenoiser(sr, 50.0, True), x, learn, prm, 1 << 20)
    np.save(out, y)


if __name__ == "__main__":
    main()
