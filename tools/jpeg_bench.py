"""GPU box: throughput of the pool-ingest path -- scoring straight from JPEG files (device decode overlapped with the
previous chunk's forward passes) against scoring the same images handed over as decoded pixels, plus the decode alone.

    python tools/jpeg_bench.py [n_images]
"""
import io
import os
import random
import sys
import time

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cald_b200 import api, synth  # noqa: E402
from cald_b200.engine import Engine, expand_augs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, W, NC = 800, 1333, 91
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
eng = Engine(depth=50, num_classes=NC, min_size=800, max_size=1333)
eng.load_state_dict(synth.planted_frcnn_weights(50, NC, 0))
base = [synth.synth_image(100 + i, H, W) for i in range(16)]
files = []
for i in range(n):
    buf = io.BytesIO()
    Image.fromarray(np.roll(base[i % 16], 13 * (i // 16), axis=1)).save(buf, format="JPEG", quality=90)
    files.append(buf.getvalue())
print("%d files, mean %.0f KB (decoded: %.0f KB)" % (n, np.mean([len(f) for f in files]) / 1e3, H * W * 3 / 1e3))
t = time.time()
pix_cpu = [np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files[:32]]
t_pil = (time.time() - t) / 32
eng.decode_jpeg(files[:16])                                   # warm-up
t = time.time()
pix = eng.decode_jpeg(files)
t_dec = (time.time() - t) / n
assert all(np.array_equal(a, b) for a, b in zip(pix[:32], pix_cpu))
print("decode only: device %.2f ms / image (incl. D2H of the pixels), Pillow on one host core %.2f ms / image" % (
    1e3 * t_dec, 1e3 * t_pil))
views = expand_augs(AUGS)
u = np.random.RandomState(0).random_sample(200 * n)
eng.score(pix[:64], views, 1.3, u[:200 * 64])                 # warm-up
t = time.time()
c1, v1, _ = eng.score(pix, views, 1.3, u)
t_pix = time.time() - t
rates = {}
for walk in ("host", "device"):                                # where the entropy-coded segment is walked (UploadPipe)
    os.environ["CALD_JPEG_WALK"] = walk
    eng.score_jpeg(files[:64], views, 1.3, u[:200 * 64])      # warm-up (staging buffers of this mode)
    t = time.time()
    c2, v2, _, _, _ = eng.score_jpeg(files, views, 1.3, u)
    rates[walk] = n / (time.time() - t)
    assert np.array_equal(c1, c2) and np.array_equal(v1, v2)
print("scoring %d images: from decoded pixels %.1f img/s, from JPEG files %.1f img/s (entropy walk on host threads) / "
      "%.1f img/s (entropy walk on the device); identical scores" % (n, n / t_pix, rates["host"], rates["device"]))
