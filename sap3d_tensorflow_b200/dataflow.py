"""Clip indexing and the input pipeline of the reference (dataflow.py; used by train.py:96-135 and test.py), without tensorpack.

  VideoDataset        dataflow.py:15-156 -- same constructor / method names: `setup_video_dataset_p3d` cuts every video into
                      16-frame clips (first frame `skip_head`, stride 16 - overlap), shuffles the (video, first frame) tuples
                      and splits them 80 / 20; `get_frame_p3d_tf` expands them into lists of frame / density (/ fixation) file
                      names (`frame_%d.jpg`, 1-based; fixations `frame_%d.bmp`).
  ClipLoader          the tensorpack chain MultiThreadMapData(mapf) -> BatchData -> PrefetchDataZMQ (train.py:120-135;
                      dataflow.py:190-233) as a thread pool that decodes ahead of the consumer.  What `mapf` does per frame
                      on the CPU -- BGR->RGB, minus the channel mean, resize to 112, /255 -- is NOT done here: the loader hands
                      over the decoded uint8 BGR frames (pinned host memory when a GPU is present) and `ClipLoader.to_device`
                      runs that arithmetic as one kernel over the whole batch (`sap3d_preprocess_frames`, video.py).  The
                      1-channel density / fixation maps are small; they follow `mapf` / `mapf_test` on the host with the same
                      cv2 calls (uint8 bilinear resize, /255).

Differences from the reference, all deliberate: clip order through the pool is preserved (tensorpack's strict mode only
promises completeness); shuffles take an optional seed; `get_frame_p3d_tf` with mini_batch > 1 lists EVERY clip of a batch
(the reference's lists are rebuilt per clip but appended once per batch, so all but the last clip of each batch are dropped --
dataflow.py:84-117; its drivers only ever call it with the default mini_batch = 1, where the two agree); fixation names go to
the fixation list in the training phase too (dataflow.py:109 appends them to the density list).
"""
from __future__ import annotations

import glob
import os
import random
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

FRAME_WILDCARD = "frame_%d.jpg"       # dataflow.py:75-77
GT_WILDCARD = "frame_%d.jpg"
FIX_WILDCARD = "frame_%d.bmp"
TEST_DENSITY_SIZE = (960, 1080)       # cv2.resize(im, (960, 1080)) in mapf_test, dataflow.py:225 -> arrays [1080, 960]


class VideoDataset:
    def __init__(self, frame_basedir: Sequence[str], density_basedir: Sequence[str], fixation_dir: Optional[str] = None,
                 img_size=(480, 288), video_length: int = 16, stack: int = 5, bgr_mean_list=(103.939, 116.779, 123.68), sort: str = "bgr",
                 seed: Optional[int] = None):
        mean = np.array(bgr_mean_list, dtype=np.float32)
        if sort == "rgb":
            mean = mean[::-1]
        self.MEAN_VALUE = mean[None, ...]
        self.img_size = img_size
        self.video_length = video_length
        self.step = 1
        assert self.step < self.video_length
        if isinstance(frame_basedir, str) or isinstance(density_basedir, str):
            raise TypeError("frame_basedir / density_basedir are LISTS of directories (train.py:99-100)")
        self.frame_basedir = list(frame_basedir)
        self.density_basedir = list(density_basedir)
        self.video_dir_list: List[str] = []
        for each_video in self.frame_basedir:
            self.video_dir_list += sorted(glob.glob(os.path.join(each_video, "*")))   # sorted: glob order is filesystem order
        self.fixation_dir = fixation_dir
        self._rng = random.Random(seed) if seed is not None else random

    # dataflow.py:38-66
    def setup_video_dataset_p3d(self, overlap: int = 2, training_example_props: float = 0.8, skip_head: int = 11,
                                shuffle_tuples: bool = True):
        """shuffle_tuples=False is dataflow_list.py:57 (the same class with the shuffle commented out, used by gen / list runs)"""
        assert overlap < self.video_length, "overlap should smaller than videolength."
        self.tuple_list: List[Tuple[int, int]] = []
        step = self.video_length - overlap
        for i, video_dir in enumerate(self.video_dir_list):
            total_frame = len(glob.glob(os.path.join(video_dir, "*.*")))
            for j in range(skip_head, total_frame, step):
                if j + self.video_length > total_frame:
                    break
                self.tuple_list.append((i, j))          # video index, first frame index (0-based; file names are 1-based)
        self.num_examples = len(self.tuple_list)
        if shuffle_tuples:
            self._rng.shuffle(self.tuple_list)
        self.num_training_examples = int(self.num_examples * training_example_props)
        self.training_tuple_list = self.tuple_list[:self.num_training_examples]
        self.validation_tuple_list = self.tuple_list[self.num_training_examples:]
        self.num_validation_examples = len(self.validation_tuple_list)
        self.num_epoch = 0
        self.index_in_training_epoch = 0
        self.index_in_validation_epoch = 0
        self.final_train_list: List[list] = []
        self.final_valid_list: List[list] = []

    def _density_dir(self, video_name: str) -> str:
        found = None
        for each_density_dir in self.density_basedir:     # the LAST base directory that has the video wins (dataflow.py:94-97)
            if os.path.exists(os.path.join(each_density_dir, video_name)):
                found = os.path.join(each_density_dir, video_name)
        if found is None:
            raise FileNotFoundError(f"no density directory for video {video_name}")
        return found

    def _clip_files(self, tup: Tuple[int, int]) -> list:
        video_index, start = tup
        video_dir = self.video_dir_list[video_index]
        video_name = os.path.basename(video_dir)
        density_dir = self._density_dir(video_name)
        idx = range(start + 1, start + self.video_length + 1)
        out = [[os.path.join(video_dir, FRAME_WILDCARD % k) for k in idx], [os.path.join(density_dir, GT_WILDCARD % k) for k in idx]]
        if self.fixation_dir:
            out.append([os.path.join(self.fixation_dir, video_name, FIX_WILDCARD % k) for k in idx])
        for group in out:
            for f in group:
                if not os.path.exists(f):          # the reference indexes glob.glob(...)[0]: IndexError on a missing file
                    raise FileNotFoundError(f)
        return out

    # dataflow.py:68-156
    def get_frame_p3d_tf(self, mini_batch: int = 1, phase: str = "training", density_length: str = "full"):
        """fills final_train_list / final_valid_list with one [frame files, density files(, fixation files)] entry per clip.
        Loop bounds as in the reference: training takes floor(n / mini_batch) batches; validation stops one clip earlier
        (`while not index >= n - mini_batch`, dataflow.py:122), i.e. the last batch that would end exactly at n is dropped."""
        self.index_in_training_epoch += mini_batch
        index = 0
        while index <= self.num_training_examples - mini_batch:
            for tup in self.training_tuple_list[index:index + mini_batch]:
                self.final_train_list.append(self._clip_files(tup))
            index += mini_batch
        index = 0
        while not index >= self.num_validation_examples - mini_batch:
            for tup in self.validation_tuple_list[index:index + mini_batch]:
                self.final_valid_list.append(self._clip_files(tup))
            index += mini_batch


def _imread(path: str, flag):
    import cv2
    im = cv2.imread(path, flag)
    if im is None:
        raise IOError(f"cv2.imread failed: {path}")
    return im


def load_clip(files: list, size: int = 112, test_time: bool = False) -> Dict[str, np.ndarray]:
    """the host half of `mapf` / `mapf_test` (dataflow.py:190-233) for one clip: decoded frames stay uint8 BGR (their arithmetic runs
    on the GPU); density maps: IMREAD_GRAYSCALE -> cv2.resize to size x size on the uint8 image (tensorpack imgaug.Resize(112))
    -> / 255; test time: density resized to 960 x 1080, fixation maps / 255 at their own size."""
    import cv2
    frames = np.stack([_imread(f, cv2.IMREAD_COLOR) for f in files[0]])
    out = {"frames": frames}
    dens = []
    for f in files[1]:
        im = _imread(f, cv2.IMREAD_GRAYSCALE)
        im = cv2.resize(im, TEST_DENSITY_SIZE) if test_time else cv2.resize(im, (size, size), interpolation=cv2.INTER_LINEAR)
        dens.append(im / 255.0)
    out["density"] = np.stack(dens).astype(np.float32)
    if len(files) > 2:
        out["fixation"] = np.stack([_imread(f, cv2.IMREAD_GRAYSCALE) / 255.0 for f in files[2]]).astype(np.float32)
    return out


class ClipLoader:
    """for batch in ClipLoader(dataset.final_train_list, batch=8): x, y = loader.to_device(batch); sess.train_step(x, y)

    nr_thread decoders run `buffer_size` clips ahead of the consumer (MultiThreadMapData(nr_thread=16, buffer_size=1000),
    train.py:121-126); batches are assembled into pinned host tensors when a GPU is present.  `remainder=False` drops the
    ragged last batch (BatchData(..., remainder=False), train.py:133); `shuffle` reorders the clips every epoch
    (ImageFromFile.__iter__, dataflow.py:175-177)."""

    def __init__(self, clips: Sequence[list], batch: int, nr_thread: int = 16, buffer_size: int = 64, remainder: bool = False,
                 shuffle: bool = True, seed: Optional[int] = None, size: int = 112, test_time: bool = False):
        assert len(clips), "No image files given to ClipLoader!"
        self.clips = list(clips)
        self.batch, self.nr_thread, self.buffer_size = int(batch), int(nr_thread), max(int(buffer_size), int(batch))
        self.remainder, self.shuffle, self.size, self.test_time = remainder, shuffle, size, test_time
        self._rng = random.Random(seed) if seed is not None else random
        self._pin = torch.cuda.is_available()

    def __len__(self) -> int:
        n = len(self.clips)
        return (n + self.batch - 1) // self.batch if self.remainder else n // self.batch

    def _collate(self, items: List[Dict[str, np.ndarray]]) -> Dict[str, torch.Tensor]:
        out = {}
        for k in items[0]:
            shapes = {it[k].shape for it in items}
            if len(shapes) != 1:
                raise ValueError(f"clips of one batch differ in {k} size: {sorted(shapes)}")
            first = items[0][k]
            t = torch.empty((len(items),) + first.shape, dtype=torch.from_numpy(first[:0]).dtype, pin_memory=self._pin)
            for i, it in enumerate(items):
                t[i].copy_(torch.from_numpy(np.ascontiguousarray(it[k])))
            out[k] = t
        return out

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        order = list(range(len(self.clips)))
        if self.shuffle:
            self._rng.shuffle(order)
        with ThreadPoolExecutor(max_workers=self.nr_thread) as pool:
            pending: deque = deque()
            it = iter(order)

            def fill():
                while len(pending) < self.buffer_size:
                    i = next(it, None)
                    if i is None:
                        return
                    pending.append(pool.submit(load_clip, self.clips[i], self.size, self.test_time))

            fill()
            items: List[Dict[str, np.ndarray]] = []
            while pending:
                items.append(pending.popleft().result())     # clip order is preserved
                fill()
                if len(items) == self.batch:
                    yield self._collate(items)
                    items = []
            if items and self.remainder:
                yield self._collate(items)

    def to_device(self, batch: Dict[str, torch.Tensor], dtype: str = "f32"):
        """(x [B,16,size,size,3], density [B,16,...], fixation or None) on the GPU: the frames go through
        `sap3d_preprocess_frames` (BGR->RGB, minus [90,102,98], bilinear resize, /255 -- dataflow.py:194-209 on the device)."""
        from . import video
        f = batch["frames"]
        B, T = f.shape[0], f.shape[1]
        x = video.preprocess_frames(f.reshape((B * T,) + tuple(f.shape[2:])).cuda(non_blocking=True), self.size, dtype)
        x = x.reshape(B, T, self.size, self.size, 3)
        y = batch["density"].cuda(non_blocking=True)
        fx = batch["fixation"].cuda(non_blocking=True) if "fixation" in batch else None
        return x, y, fx
