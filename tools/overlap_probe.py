"""GPU box: can a small kernel on a second stream run while the persistent conv kernels hold every SM?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200._lib import lib, check_ops
L = lib()
L.cald_op_overlap_probe.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_double)]
for name, shape in (("one-CTA kernel, 1x1 64->256 (226 KB smem)", (16, 200, 336, 64, 256, 1)),
                    ("pair kernel, 3x3 256->256 (202 KB smem)", (8, 200, 336, 256, 256, 3))):
    for smem in (0, 4096):
        out = (ctypes.c_double * 4)()
        check_ops(L.cald_op_overlap_probe(*shape, 40, 20, smem, out))
        print("%-44s spin smem %4d B: conv train %.2f ms alone, %.2f ms with the spin kernel; spin kernel %.2f ms alone, "
              "%.2f ms start-to-end inside the train" % (name, smem, out[0], out[1], out[2], out[3]))
