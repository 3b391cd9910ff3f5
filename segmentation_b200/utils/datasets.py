"""Device-side input path for the models (reference utils/datasets.py).

The reference feeds its graphs from TF queue runners: files are decoded to uint8, divided by
255, image and mask are cropped TOGETHER at a random position and shuffle-batched
(`ImageMaskDataSet._preprocessing`, utils/datasets.py:176-190), and `load_images` does the same
with numpy for inference (`:19-45`).  File decoding (TF readers, cv2) is outside this build;
everything after it is here, with the arithmetic on the GPU:

  * the dataset hands the model RAW batches - uint8 images [B,Hs,Ws,3], uint8 masks
    [B,Hs,Ws,1] with values 0/255 - and one crop corner per image;
  * `BaseModel.train_step()` uploads those bytes (4x fewer than fp32) and ONE staging launch
    (`seg_stage_input`, include/segb200.h) does the `/255`, the joint crop, `uint8(mask/255)`
    and the bf16 packing that the first convolution reads.

`ArrayImageMaskDataSet` is the duck type the models consume (`batch_size`, `use_feed`,
`has_masks`, `set_tf_sess`, `next_batch`) over in-memory arrays.
"""
import numpy as np
import torch


class ArrayImageMaskDataSet(object):
    """images: uint8 [N,Hs,Ws,3]; masks: uint8 [N,Hs,Ws,1] (0 / 255, as the mask files decode).
    next_batch() -> (images[B], masks[B], crop_yx[B,2]): `batch_size` images drawn at random
    (the shuffle batch of :163-168) with a uniform random crop corner per image
    (tf.random_crop of the stacked image+mask, :184-185).  Host tensors are pinned so that the
    model's copy stream uploads them asynchronously."""
    use_feed, has_masks = False, True

    def __init__(self, images, masks, batch_size=96, crop_size=256, n_classes=2, seed=5555,
                 pinned=True):
        assert images.dtype == np.uint8 and masks.dtype == np.uint8
        assert images.shape[:3] == masks.shape[:3] and images.shape[3] == 3
        assert images.shape[1] >= crop_size and images.shape[2] >= crop_size
        self.images, self.masks = images, masks
        self.batch_size, self.crop_size, self.n_classes = batch_size, crop_size, n_classes
        self.rng = np.random.default_rng(seed)
        self.pinned = pinned and torch.cuda.is_available()

    def set_tf_sess(self, sess):
        pass

    def draw(self):
        """Indices and crop corners of the next batch (numpy)."""
        n, hs, ws = self.images.shape[:3]
        idx = self.rng.integers(0, n, self.batch_size)
        cy = self.rng.integers(0, hs - self.crop_size + 1, self.batch_size)
        cx = self.rng.integers(0, ws - self.crop_size + 1, self.batch_size)
        return idx, np.stack([cy, cx], 1).astype(np.int32)

    def next_batch(self):
        idx, crop = self.draw()
        x = torch.from_numpy(self.images[idx])
        y = torch.from_numpy(self.masks[idx])
        c = torch.from_numpy(crop)
        if self.pinned:
            x, y, c = x.pin_memory(), y.pin_memory(), c.pin_memory()
        return x, y, c

    def reference_batch(self, idx, crop):
        """What utils/datasets.py:176-190 computes for the same draw, on the host: float32
        images / 255 and uint8(mask / 255), both cropped at `crop` (the parity oracle of the
        staging launch)."""
        s = self.crop_size
        xs = np.stack([self.images[i, cy:cy + s, cx:cx + s] for i, (cy, cx) in zip(idx, crop)])
        ys = np.stack([self.masks[i, cy:cy + s, cx:cx + s] for i, (cy, cx) in zip(idx, crop)])
        return xs.astype(np.float32) / np.float32(255.0), (ys / 255).astype(np.uint8)


def load_images(images, batchsize, crop_size, rng=None):
    """`load_images(paths, batchsize, crop_size)` (utils/datasets.py:19-45) over decoded
    uint8 arrays: a random choice of `batchsize` images, one random crop each.  Returns
    (uint8 [B,Hs,Ws,3], crop_yx int32 [B,2]) for `model.infer_raw`-style staging; the `/255`
    runs on the device."""
    rng = rng or np.random.default_rng()
    n, hs, ws = images.shape[:3]
    idx = rng.choice(n, batchsize)
    cy = rng.integers(0, hs - crop_size, batchsize) if hs > crop_size else np.zeros(batchsize, int)
    cx = rng.integers(0, ws - crop_size, batchsize) if ws > crop_size else np.zeros(batchsize, int)
    return images[idx], np.stack([cy, cx], 1).astype(np.int32)
