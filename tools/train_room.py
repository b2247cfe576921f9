"""Train the renderer on the analytic room scene (mirror_nerf_b200/room_scene.py) with the training path of this repo
(mirror_nerf_b200/room_trainer.py) and save the resulting *scene-like* field as tests/golden/room_field.npz (fp16 storage;
loaded back as fp32).

    python tools/train_room.py [--steps 6000] [--rays 4096] [--out tests/golden/room_field.npz]

Prints the PSNR against the analytic ground truth on a held-out view as training proceeds."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mirror_nerf_b200.room_trainer import RoomTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6000)
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--lr", type=float, default=5e-4)
ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "room_field.npz"))
args = ap.parse_args()

tr = RoomTrainer(lr=args.lr, rays_per_step=args.rays)
t0 = time.time()
for it in range(args.steps + 1):
    if it % 500 == 0:
        p, mf = tr.psnr()
        print(f"step {it:5d}  psnr {p:6.2f} dB  predicted mirror fraction {mf:.3f}  ({time.time() - t0:.0f} s)", flush=True)
    if it == args.steps:
        break
    # step decay (R/opt.py: steplr)
    loss = tr.step(1.0 if it < args.steps // 2 else (0.4 if it < 5 * args.steps // 6 else 0.15))
    if it % 100 == 0:
        print(f"  it {it} loss {float(loss):.5f}", flush=True)

sd = {}
for k, m in tr.models.items():
    for name, p in m.state_dict().items():
        sd[f"{k}/{name}"] = p.detach().cpu().numpy().astype(np.float16)
np.savez_compressed(args.out, **sd)
print("saved", args.out, os.path.getsize(args.out) // 1024, "KiB")
