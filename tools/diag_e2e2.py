import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cald_b200 import synth, engine as E
from cald_b200.engine import Engine, expand_augs
H, W, B = 800, 1333, 8
kinds = expand_augs(['flip', 'cut_out', 'smaller_resize', 'rotation'])
eng = Engine(depth=50, num_classes=91, min_size=800, max_size=1333, max_views_per_pass=32)
eng.load_state_dict(synth.planted_frcnn_weights(50, 91, 0))
pinned = [torch.from_numpy(synth.synth_image(i, H, W, 0)).pin_memory() for i in range(B)]
pool = [t.numpy() for t in pinned]
pageable = [p.copy() for p in pool]
print('ascontig identity: pinned', np.ascontiguousarray(pool[0], dtype=np.uint8) is pool[0], 'pageable', np.ascontiguousarray(pageable[0], dtype=np.uint8) is pageable[0], pool[0].flags, type(pool[0]))
u = np.random.RandomState(0).random_sample(200 * B)
for name, imgs in (("pageable", pageable), ("pinned", pool), ("pageable", pageable), ("pinned", pool)):
    for it in range(3):
        t0 = time.time(); r = E._u8_list(imgs); t1 = time.time()
        eng.score(imgs, kinds, 1.3, u); t2 = time.time()
        print(name, "u8_list %.2f ms, score %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
