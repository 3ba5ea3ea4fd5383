"""How does the CPU oracle scale with torch threads on this host? (picks the thread count of the cpu_baseline)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import synthetic
from oracle import mingtok_oracle as O
cfg = synthetic.MINGTOK_CONFIG
sd = synthetic.mingtok_state_dict(cfg, 0)
img = synthetic.synthetic_images(4, 256)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    with torch.no_grad():
        O.mingtok_forward_enc_dec(sd, img[:1], cfg)
        t0 = time.perf_counter(); O.mingtok_forward_enc_dec(sd, img, cfg); dt = time.perf_counter() - t0
    print(f"threads {th}: 4 images {dt:.2f} s -> {256/dt:.1f} tokens/s", flush=True)
