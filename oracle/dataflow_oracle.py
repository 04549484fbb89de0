"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement of the reference's input pipeline, dataflow.py, with the same cv2 calls it makes (OpenCV is the reference's own
dependency and is present here, so these functions ARE the reference arithmetic; tensorpack's imgaug.Resize(112) is
`cv2.resize(img, (112, 112), interpolation=cv2.INTER_LINEAR)`):
  clip_tuples      dataflow.py:38-50   (video index, first frame) tuples before the shuffle
  mapf             dataflow.py:190-209 frames: imread -> [:, :, ::-1] -> minus mean (float32) -> Resize(112) -> / 255.
                                        density: imread gray -> Resize(112) on uint8 -> / 255.
  mapf_test        dataflow.py:212-233 density resized to (960, 1080); fixation / 255.
"""
import numpy as np

MEAN_VALUE = np.array([98, 102, 90], dtype=np.float32)[::-1][None, ...]     # dataflow.py:186-188


def clip_tuples(frames_per_video, video_length=16, overlap=2, skip_head=11):
    out = []
    step = video_length - overlap
    for i, total in enumerate(frames_per_video):
        j = skip_head
        while j < total:
            if j + video_length > total:
                break
            out.append((i, j))
            j += step
    return out


def mapf(frame_files, density_files):
    import cv2
    ret_frame, ret_density = [], []
    for f in frame_files:
        im = cv2.imread(f, cv2.IMREAD_COLOR)
        im = im[:, :, ::-1]
        im = im - MEAN_VALUE
        im = cv2.resize(im, (112, 112), interpolation=cv2.INTER_LINEAR)
        ret_frame.append(im / 255.)
    for f in density_files:
        im = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        im = cv2.resize(im, (112, 112), interpolation=cv2.INTER_LINEAR)
        ret_density.append(im / 255.)
    return ret_frame, ret_density


def mapf_test(frame_files, density_files, fixation_files):
    import cv2
    ret_frame = mapf(frame_files, [])[0]
    ret_density = [cv2.resize(cv2.imread(f, cv2.IMREAD_GRAYSCALE), (960, 1080)) / 255. for f in density_files]
    ret_fixation = [cv2.imread(f, cv2.IMREAD_GRAYSCALE) / 255. for f in fixation_files]
    return ret_frame, ret_density, ret_fixation
