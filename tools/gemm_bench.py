"""Microbenchmark of the training GEMM kernels with the bring-up knock-out flags of train_tc.cu."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mirror_nerf_b200 import _lib
from mirror_nerf_b200.mirror_nerf import MirrorNeRF, packed_field
from mirror_nerf_b200.synthetic import make_state_dict
lib = _lib.load()
m = MirrorNeRF(predict_normal=True, predict_mirror_mask=True); m.load_state_dict(make_state_dict(1)); m = m.cuda()
pf = packed_field(m)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 786432
def run(kind, step, engine, dbg):
    ms = C.c_float()
    _lib.check(lib.mnrf_debug_gemm_bench(pf.handle, kind, step, P, engine, dbg, 5, C.byref(ms)), "bench")
    return ms.value
flop = 2.0 * P * 256 * 256
for kind, step, name in ((0, 1, "NN 256x256"), (0, 9, "NN N128 K256"), (0, 19, "NN N64 K256"), (1, 0, "TN 256x256")):
    base = run(kind, step, 0, 0) if True else 0
    print(f"{name:14s} simt {base:7.3f} ms", end="", flush=True)
    for dbg in (0, 1, 2, 3, 4, 8, 16, 1 | 8, 1 | 2 | 8 | 16, 1 | 2 | 4 | 8 | 16):
        if kind == 1 and dbg >= 16: continue
        t = run(kind, step, 1, dbg)
        print(f" | dbg{dbg}: {t:6.3f}", end="", flush=True)
    print(f" | tf32x1: {run(kind, step, 2, 0):6.3f}", end="")
    print(f"   (ideal tensor {flop * 3 / 1.13e15 * 1e3 * (128 if step == 9 else 64 if step == 19 else 256) / 256:.3f} ms)")
