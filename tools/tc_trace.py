"""Device-side timeline of the tcgen05 field kernel (CTA 0): prints per-tile event offsets in cycles.
Needs the trace flavour of the library:
    make -C mirror_nerf_b200/csrc OBJDIR=../lib/obj_trace OUT=../lib/libmnrf_trace.so EXTRA=-DMNRF_TC_TRACE
    MNRF_LIB=mirror_nerf_b200/lib/libmnrf_trace.so python tools/tc_trace.py tc2"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_models  # noqa: E402
from mirror_nerf_b200 import _lib  # noqa: E402
from mirror_nerf_b200.rendering import render_rays  # noqa: E402
from mirror_nerf_b200.synthetic import camera_rays  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "tc3"
lib = _lib.load()
models, emb = make_models()
rays = camera_rays(800, 800)[::19][:32768].contiguous().cuda()
cap = 200000
kw = dict(test_time=True, compute_normal=False, field_impl=impl, fused_composite=False)
with torch.no_grad():
    # a single full-model launch: coarse-only render with train-style (non sigma-only) coarse pass, 192 samples
    one = {"coarse": models["fine"]}
    kw["test_time"] = False
    render_rays(one, emb, rays, 192, False, 0, 0, 0, 32768, False, **kw)
    buf = torch.zeros(1 + 2 * cap, dtype=torch.int64, device="cuda")
    lib.mnrf_debug_set_trace(buf.data_ptr(), cap)
    render_rays(one, emb, rays, 192, False, 0, 0, 0, 32768, False, **kw)
    torch.cuda.synchronize()
    lib.mnrf_debug_set_trace(None, 0)
b = buf.cpu().numpy()
ev = []
for who in range(12):
    base = 1 + who * 2 * 8192
    for i in range(8192):
        t, tag = int(b[base + 2 * i]), int(b[base + 2 * i + 1])
        if t == 0:
            break
        ev.append((t, tag))
ev.sort()
print("events", len(ev))
# the trace holds the coarse launch then the fine launch: split at the largest time gap, keep the fine one
fine = ev
names = {1: "mma_begin", 2: "mma_A_ready", 3: "mma_last_chunk", 4: "mma_stage_issue", 5: "TILE wait PE cycles", 6: "TILE wait A cycles",
         7: "TILE wait W cycles", 20: "epi_first_ld_landed", 21: "epi_piece0_stored", 22: "epi_piece_stored", 10: "epi_acc_ready", 11: "epi_chunk_written", 12: "pe_begin", 13: "pe_end", 14: "tile_epilogue_done"}
# per-tile totals of the MMA issuer's blocked cycles (payload in the upper 32 bits of the tag)
tot = {5: [], 6: [], 7: []}
for t, tag in ev:
    e = (tag >> 16) & 255
    if (tag >> 24) & 255 == 1 and e in tot:
        tot[e].append(tag >> 32)
for e, v in tot.items():
    if v:
        v = v[2:-1] or v
        print(f"MMA issuer, {names[e]}: mean {sum(v) / len(v):.0f} over {len(v)} tiles")
starts = [i for i, (t, tag) in enumerate(fine)
          if ((tag >> 24) & 255) == 1 and ((tag >> 16) & 255) == 1 and ((tag >> 8) & 255) == 0 and (tag & 255) == 0]
print("tiles traced", len(starts))
print("tile period (cycles):", [fine[starts[j + 1]][0] - fine[starts[j]][0] for j in range(2, min(14, len(starts) - 1))])
k = min(20, len(starts) - 2)
lo, hi = starts[k], starts[k + 1]
t0 = fine[lo][0]
for t, tag in fine[lo:hi + 8]:
    who, e, a, bb = (tag >> 24) & 255, (tag >> 16) & 255, (tag >> 8) & 255, tag & 255
    if e in (5, 6, 7):
        a, bb = tag >> 32, 0
    print(f"{t - t0:8d}  {'MMA ' if who == 1 else 'EPI' + str(who - 4)}  {names.get(e, e):20s} {a} {bb}")
