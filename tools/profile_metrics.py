"""The saliency-metric kernels between cudaProfilerStart/Stop, for ncu: native-resolution CC/SIM/NSS/KLdiv over a large batch
of 112 x 112 maps, and the test-time path (cv2-style resize to 1080 x 960 + CC/SIM/NSS + AUC_Judd/AUC_Borji) over a few.
    ncu --profile-from-start off --set full -k regex:metric ... python tools/profile_metrics.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sap3d_tensorflow_b200 as sp  # noqa: E402

torch.manual_seed(0)
B = 8192
pred = torch.rand(B, 112, 112, device="cuda")
dens = torch.rand(B, 112, 112, device="cuda")
fix = (torch.rand(B, 112, 112, device="cuda") < 0.01).float()
small = torch.rand(8, 16, 112, 112, 1, device="cuda")
dens_big = torch.rand(8, 1080, 960, device="cuda")
fix_big = (torch.rand(8, 1080, 960, device="cuda") < 0.001).float()
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    v = sp.metrics.saliency_metrics(pred, dens, fix)
    w = sp.metrics.evaluate_clips_test_time(small, dens_big, fix_big)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", v.shape, float(v[0, 0]))
