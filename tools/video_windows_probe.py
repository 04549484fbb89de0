"""windows/s of the sliding-window inference (gen_pred.py:88-135) with and without the per-frame stem cache: a synthetic video of
T frames at 112 x 112, B windows per run, per-clip BatchNorm statistics.  python tools/video_windows_probe.py [T] [B]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sap3d_tensorflow_b200 as sp  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 272
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
size = 112
rng = np.random.RandomState(1)
frames = sp.video.preprocess_frames(rng.randint(0, 256, (T, 120, 160, 3)).astype(np.uint8), size=size)
full = sp.Session(sp.p3d.p3d_unetplusplus_ds(sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=False, per_sample_statistics=True), 0.0, B, False))
stem = sp.Session(sp.p3d.p3d_stem(sp.placeholder([32, 1, size, size, 3], dtype="bf16", training_graph=False)))
win = sp.Session(sp.p3d.p3d_unetplusplus_ds(sp.placeholder([B, 16, size // 2, size // 2, 64], dtype="bf16", training_graph=False, per_sample_statistics=True), 0.0, B, False))
nwin = len(sp.video.window_starts(T))
for name, fn in (("predict_video", lambda: sp.video.predict_video(full, frames, graph=True)),
                 ("predict_video_cached", lambda: sp.video.predict_video_cached(stem, win, frames, graph=True))):
    for _ in range(2):
        n = sum(1 for _ in fn())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        n = sum(1 for _ in fn())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {T} frames -> {nwin} windows ({n} maps) in {dt * 1e3:.1f} ms = {nwin / dt:.0f} windows/s (B = {B} windows per run, 112 x 112, bf16)")
