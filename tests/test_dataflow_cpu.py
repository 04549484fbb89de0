"""Host side of the input pipeline (sap3d_tensorflow_b200/dataflow.py vs the reference's dataflow.py, restated with its own cv2
calls in oracle/dataflow_oracle.py) on a synthetic on-disk dataset: clip indexing, the train / validation split and its
loop-bound quirks, file naming, the threaded loader's batches.  The GPU half (frames through sap3d_preprocess_frames) is covered
by tests/test_metrics.py::test_preprocess_frames_*; here the decoded frames are checked against cv2.imread."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.usefixtures("lib_built")

FRAMES = (40, 27, 61)          # frames per video: 27 < skip_head + 16 + ... gives exactly one clip at overlap 2


@pytest.fixture(scope="module")
def dataset_dir(tmp_path_factory):
    import cv2
    root = tmp_path_factory.mktemp("svsd")
    rng = np.random.RandomState(0)
    for v, n in enumerate(FRAMES):
        for sub in ("frames", "density", "fixation"):
            os.makedirs(root / sub / f"video{v}")
        for k in range(1, n + 1):
            img = cv2.GaussianBlur(rng.randint(0, 256, (54, 96, 3)).astype(np.uint8), (5, 5), 0)
            cv2.imwrite(str(root / "frames" / f"video{v}" / f"frame_{k}.jpg"), img)
            den = cv2.GaussianBlur(rng.randint(0, 256, (54, 96)).astype(np.uint8), (9, 9), 0)
            cv2.imwrite(str(root / "density" / f"video{v}" / f"frame_{k}.jpg"), den)
            fix = (rng.rand(54, 96) < 0.01).astype(np.uint8) * 255
            cv2.imwrite(str(root / "fixation" / f"video{v}" / f"frame_{k}.bmp"), fix)
    return root


def make(root, fixation=False, seed=3, **kw):
    from sap3d_tensorflow_b200.dataflow import VideoDataset
    return VideoDataset([str(root / "frames")], [str(root / "density")], fixation_dir=str(root / "fixation") if fixation else None,
                        video_length=16, img_size=(112, 112), bgr_mean_list=[98, 102, 90], sort="rgb", seed=seed, **kw)


def test_clip_index_and_split(dataset_dir):
    from oracle import dataflow_oracle as DO
    ds = make(dataset_dir)
    assert [os.path.basename(d) for d in ds.video_dir_list] == ["video0", "video1", "video2"]
    np.testing.assert_array_equal(ds.MEAN_VALUE, [[90, 102, 98]])                       # sort='rgb' reverses the BGR list
    for overlap in (2, 8, 15):
        ds.setup_video_dataset_p3d(overlap=overlap, training_example_props=0.9)
        want = DO.clip_tuples(FRAMES, 16, overlap, 11)
        assert sorted(ds.tuple_list) == sorted(want) and ds.num_examples == len(want)
        assert ds.tuple_list != want or len(want) < 3                                    # shuffled
        assert ds.num_training_examples == int(len(want) * 0.9)
        assert ds.training_tuple_list + ds.validation_tuple_list == ds.tuple_list
    # every clip stays inside its video: first frame >= skip_head, last frame < total
    assert all(11 <= j and j + 16 <= FRAMES[i] for i, j in ds.tuple_list)
    assert (1, 11) in ds.tuple_list and not any(i == 1 and j != 11 for i, j in ds.tuple_list)   # 27 frames: one clip only
    # same seed -> same shuffle; another seed -> another order
    a, b, c = make(dataset_dir, seed=5), make(dataset_dir, seed=5), make(dataset_dir, seed=6)
    for d in (a, b, c):
        d.setup_video_dataset_p3d(overlap=15, training_example_props=0.8)
    assert a.tuple_list == b.tuple_list != c.tuple_list
    ds.setup_video_dataset_p3d(overlap=2, shuffle_tuples=False)                          # dataflow_list.py: no shuffle
    assert ds.tuple_list == DO.clip_tuples(FRAMES, 16, 2, 11)
    with pytest.raises(AssertionError):
        ds.setup_video_dataset_p3d(overlap=16)
    with pytest.raises(TypeError):
        from sap3d_tensorflow_b200.dataflow import VideoDataset
        VideoDataset(str(dataset_dir / "frames"), [str(dataset_dir / "density")])


def test_file_lists_and_loop_bounds(dataset_dir):
    ds = make(dataset_dir, fixation=True)
    ds.setup_video_dataset_p3d(overlap=8, training_example_props=0.75)
    ds.get_frame_p3d_tf()
    assert len(ds.final_train_list) == ds.num_training_examples               # `while index <= n - 1`
    assert len(ds.final_valid_list) == ds.num_validation_examples - 1         # `while not index >= n - 1` drops the last clip
    (vi, j), entry = ds.training_tuple_list[0], ds.final_train_list[0]
    assert len(entry) == 3 and all(len(g) == 16 for g in entry)
    assert entry[0][0].endswith(os.path.join("frames", f"video{vi}", f"frame_{j + 1}.jpg"))      # 1-based file names
    assert entry[0][15].endswith(os.path.join("frames", f"video{vi}", f"frame_{j + 16}.jpg"))
    assert entry[1][3].endswith(os.path.join("density", f"video{vi}", f"frame_{j + 4}.jpg"))
    assert entry[2][0].endswith(os.path.join("fixation", f"video{vi}", f"frame_{j + 1}.bmp"))
    # test.py:80 uses training_example_props=0: everything is validation
    ds.setup_video_dataset_p3d(overlap=15, training_example_props=0)
    ds.get_frame_p3d_tf()
    assert ds.final_train_list == [] and len(ds.final_valid_list) == ds.num_examples - 1
    # a second density base directory that also holds the video takes precedence (dataflow.py:94-97 keeps the last match)
    from sap3d_tensorflow_b200.dataflow import VideoDataset
    two = VideoDataset([str(dataset_dir / "frames")], [str(dataset_dir / "fixation"), str(dataset_dir / "density")], seed=0)
    two.setup_video_dataset_p3d(overlap=2)
    two.get_frame_p3d_tf()
    assert os.sep + "density" + os.sep in two.final_train_list[0][1][0]
    # a hole in the data is an error, as in the reference (glob(...)[0] raises)
    missing = VideoDataset([str(dataset_dir / "frames")], [str(dataset_dir / "frames")], fixation_dir=str(dataset_dir / "density"), seed=0)
    missing.setup_video_dataset_p3d(overlap=2)
    with pytest.raises(FileNotFoundError):
        missing.get_frame_p3d_tf()


def test_loader_batches_match_the_reference_map_function(dataset_dir):
    import cv2
    from oracle import dataflow_oracle as DO
    from sap3d_tensorflow_b200.dataflow import ClipLoader
    ds = make(dataset_dir)
    ds.setup_video_dataset_p3d(overlap=8, training_example_props=0.9)
    ds.get_frame_p3d_tf()
    clips = ds.final_train_list
    loader = ClipLoader(clips, batch=2, nr_thread=4, buffer_size=3, shuffle=False)
    batches = list(loader)
    assert len(batches) == len(loader) == len(clips) // 2
    for bi, b in enumerate(batches):
        assert tuple(b["frames"].shape) == (2, 16, 54, 96, 3) and b["frames"].dtype.is_floating_point is False
        assert tuple(b["density"].shape) == (2, 16, 112, 112)
        for k in range(2):
            files = clips[bi * 2 + k]
            ref_frames, ref_density = DO.mapf(files[0], files[1])
            np.testing.assert_array_equal(b["frames"][k, 5].numpy(), cv2.imread(files[0][5], cv2.IMREAD_COLOR))
            np.testing.assert_allclose(b["density"][k].numpy(), np.stack(ref_density), rtol=0, atol=1e-7)
            # the arithmetic the GPU kernel applies to the uint8 frames, restated on the host, equals mapf's frames
            from oracle import metrics_oracle as MO
            got = MO.preprocess_frame(b["frames"][k, 5].numpy(), 112)
            np.testing.assert_allclose(got, ref_frames[5], rtol=0, atol=2e-6)
    # remainder=True keeps the ragged last batch; the pool size and read-ahead do not change the result
    n = len(clips)
    odd = ClipLoader(clips[: n - (n + 1) % 2], batch=2, nr_thread=1, buffer_size=2, shuffle=False, remainder=True)
    got = list(odd)
    assert len(got) == len(odd) and tuple(got[-1]["frames"].shape)[0] == 1
    for x, y in zip(got, batches):
        assert (x["frames"] == y["frames"]).all() and (x["density"] == y["density"]).all()
    # shuffling reorders whole clips, seeded
    s1 = [b["frames"][:, 0, 0, 0, 0].tolist() for b in ClipLoader(clips, 2, shuffle=True, seed=1)]
    s2 = [b["frames"][:, 0, 0, 0, 0].tolist() for b in ClipLoader(clips, 2, shuffle=True, seed=1)]
    assert s1 == s2


def test_test_time_loader(dataset_dir):
    from oracle import dataflow_oracle as DO
    from sap3d_tensorflow_b200.dataflow import ClipLoader
    ds = make(dataset_dir, fixation=True)
    ds.setup_video_dataset_p3d(overlap=2, training_example_props=0)
    ds.get_frame_p3d_tf()
    clips = ds.final_valid_list
    b = next(iter(ClipLoader(clips, batch=1, nr_thread=2, shuffle=False, test_time=True)))
    assert tuple(b["density"].shape) == (1, 16, 1080, 960) and tuple(b["fixation"].shape) == (1, 16, 54, 96)
    _, ref_density, ref_fix = DO.mapf_test(*clips[0])
    np.testing.assert_allclose(b["density"][0, -1].numpy(), ref_density[-1], atol=1e-7)
    np.testing.assert_allclose(b["fixation"][0, -1].numpy(), ref_fix[-1], atol=1e-7)
    assert set(np.unique(b["fixation"].numpy())) <= {0.0, 1.0}
