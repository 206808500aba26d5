import sys; sys.path.insert(0, ".")
import numpy as np
from onset_fingerprinting_b200 import synth, spectral
from oracle import spectral_np as sp
xs, _ = synth.drum_batch(3, seconds=0.6, seed=50, first_hit=10000)
for n_fft, hop in ((256, 32), (512, 128), (2048, 128), (4096, 256)):
    got = spectral.spectral_flux_batch(xs, n_fft, hop).cpu().numpy()
    for r in range(3):
        want = sp.onset_strength(xs[r], n_fft, hop)
        d = np.abs(got[r] - want)
        i = int(d.argmax())
        print(n_fft, hop, r, "max abs", d.max(), "at", i, got[r][i], want[i], "rel", (d / np.maximum(np.abs(want), 1e-3)).max())
