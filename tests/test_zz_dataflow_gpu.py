"""Device hand-off of the input pipeline: ClipLoader batches (decoded uint8 frames in pinned memory) through
`sap3d_preprocess_frames` equal the reference's `mapf` (dataflow.py:190-209, restated with its own cv2 calls in
oracle/dataflow_oracle.py), and feed a Session step."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_loader_to_device_matches_mapf_and_feeds_a_step(lib_built, tmp_path):
    import cv2
    import sap3d_tensorflow_b200 as sp
    from oracle import dataflow_oracle as DO

    rng = np.random.RandomState(0)
    for sub in ("frames", "density"):
        os.makedirs(tmp_path / sub / "video0")
    for k in range(1, 46):
        cv2.imwrite(str(tmp_path / "frames" / "video0" / f"frame_{k}.jpg"), cv2.GaussianBlur(rng.randint(0, 256, (90, 160, 3)).astype(np.uint8), (5, 5), 0))
        cv2.imwrite(str(tmp_path / "density" / "video0" / f"frame_{k}.jpg"), cv2.GaussianBlur(rng.randint(0, 256, (90, 160)).astype(np.uint8), (9, 9), 0))
    ds = sp.dataflow.VideoDataset([str(tmp_path / "frames")], [str(tmp_path / "density")], video_length=16, img_size=(112, 112),
                                  bgr_mean_list=[98, 102, 90], sort="rgb", seed=0)
    ds.setup_video_dataset_p3d(overlap=8, training_example_props=1.0, skip_head=11)
    ds.get_frame_p3d_tf()
    assert len(ds.final_train_list) == 3                      # first frames 11, 19, 27 of 45
    loader = sp.dataflow.ClipLoader(ds.final_train_list, batch=2, nr_thread=4, shuffle=False)
    batches = list(loader)
    assert len(batches) == 1
    x, y, fx = loader.to_device(batches[0])
    assert fx is None and x.is_cuda and tuple(x.shape) == (2, 16, 112, 112, 3) and tuple(y.shape) == (2, 16, 112, 112)
    for k in range(2):
        ref_frames, ref_density = DO.mapf(*ds.final_train_list[k])
        np.testing.assert_allclose(x[k].cpu().numpy(), np.stack(ref_frames), rtol=0, atol=2e-6)
        np.testing.assert_allclose(y[k].cpu().numpy(), np.stack(ref_density), rtol=0, atol=1e-7)
    xin = sp.placeholder([2, 16, 112, 112, 3], dtype="bf16", training_graph=True)
    sess = sp.Session(sp.p3d.p3d_unet(xin, 0.0, 2, True))
    loss = float(sess.train_step(x, y, graph=False).item())
    assert np.isfinite(loss) and loss > 0
