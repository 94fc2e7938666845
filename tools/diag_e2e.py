"""Where does the end-to-end (host buffer) leg spend its time?  Run on the GPU box."""
import os, sys, time, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cald_b200 import api, synth
from cald_b200.engine import Engine, expand_augs

H, W, B = 800, 1333, 8
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
kinds = expand_augs(AUGS)
eng = Engine(depth=50, num_classes=91, min_size=800, max_size=1333, max_views_per_pass=32)
eng.load_state_dict(synth.planted_frcnn_weights(50, 91, 0))
pinned = [torch.from_numpy(synth.synth_image(i, H, W, 0)).pin_memory() for i in range(B)]
pool = [t.numpy() for t in pinned]
pageable = [p.copy() for p in pool]
dev = [t.cuda() for t in pinned]
u = np.random.RandomState(0).random_sample(200 * B)

def t(fn, n=6):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.time()
    g = 0.0
    for _ in range(n):
        eng.event_record(0)
        fn()
        eng.event_record(1)
        g += eng.event_elapsed_ms(0, 1)
    torch.cuda.synchronize()
    sys.stdout.write("[gpu %.1f ms] " % (g / n))
    return 1000 * (time.time() - t0) / n

print("score_device      %.1f ms" % t(lambda: eng.score_device([d.data_ptr() for d in dev], [H] * B, [W] * B, kinds, 1.3, u)))
print("score pageable    %.1f ms" % t(lambda: eng.score(pageable, kinds, 1.3, u)))
print("score pinned      %.1f ms" % t(lambda: eng.score(pool, kinds, 1.3, u)))
print("score pageable    %.1f ms" % t(lambda: eng.score(pageable, kinds, 1.3, u)))
def apicall():
    random.seed(0)
    api.score_images(eng, pool, AUGS, chunk=B)
reg = [np.empty_like(p) for p in pool]
for r_, p_ in zip(reg, pool):
    r_[...] = p_
    assert torch.cuda.cudart().cudaHostRegister(r_.ctypes.data, r_.nbytes, 0) in (0, None) or True
print("score hostRegister'd %.1f ms" % t(lambda: eng.score(reg, kinds, 1.3, u)))
print("api.score_images  %.1f ms" % t(apicall))
x = torch.empty(B * H * W * 3, dtype=torch.uint8, device="cuda")
hp = torch.empty(B * H * W * 3, dtype=torch.uint8).pin_memory()
print("H2D 25.6MB pinned %.2f ms" % t(lambda: x.copy_(hp, non_blocking=True)))
