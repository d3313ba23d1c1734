"""Output side of the reference's ``predict.py`` for the TAI path: writes the frames of a batch of clips in the
directory layout ``compute_quant_results.py`` / the figure scripts read back (predict.py:59-100):

    <qual_result_root>/<clip_label>/gt_preceding_%04d.png   (0 .. K-1)
                                     gt_middle_%04d.png      (K .. K+T-1; contiguous clips only)
                                     gt_following_%04d.png   (K+T ..)
                                     pred_middle_%04d.png    (K .. K+T-1)
                                     pred_middle_forward_ / pred_middle_backward_ / interp_net_outputs_{1,2}_%04d.png

The float -> 8-bit conversion (clamp, inverse transform, *255, truncation, BGR -> RGB; predict.py:124-134) runs on
the device (``frames_to_uint8_b200``) so that a quarter of the bytes cross PCIe; only the PNG encoding (PIL, as in
the reference) is host work.  Dataset reading, option parsing and the quantitative evaluation stay out of scope
(SURVEY.md section 2)."""
import os

import torch

from . import ops


def frames_to_uint8(video):
    """[T,C,H,W] (or [B,T,C,H,W]) float CUDA tensor in [-1,1] -> uint8 host array [..., H, W, C]."""
    return ops.frames_to_uint8(video.detach().contiguous().float()).cpu().numpy()


def save_video_frames(video, image_root_dir, image_name_prefix, counter_start=0):
    """predict.py:113-134: ``video`` [T,C,H,W] in [-1,1] (BGR if C == 3) -> ``<prefix>_%04d.png`` files."""
    from PIL import Image
    u8 = frames_to_uint8(video)
    os.makedirs(image_root_dir, exist_ok=True)
    paths = []
    for t in range(u8.shape[0]):
        path = os.path.join(image_root_dir, '%s_%04d.png' % (image_name_prefix, t + counter_start))
        Image.fromarray(u8[t, :, :, 0] if u8.shape[-1] == 1 else u8[t]).save(path)
        paths.append(path)
    return paths


def write_clip_predictions(gen_output, preceding_frames, following_frames, clip_labels, qual_result_root, image_size,
                           gt_middle_frames=None, intermediate_preds=False):
    """The per-batch body of predict.py:59-100.  ``gen_output`` is ``env.gen_output``; frames are cropped to
    ``image_size`` (the padding of the SloMo models is dropped) exactly as the reference slices them."""
    h, w = image_size
    K, T = preceding_frames.shape[1], gen_output['pred'].shape[1]
    dev = gen_output['pred'].device
    written = []
    for b in range(gen_output['pred'].shape[0]):
        root = os.path.join(qual_result_root, clip_labels[b])
        crop = lambda v: v[b, :, :, :h, :w].to(dev)
        written += save_video_frames(crop(preceding_frames), root, 'gt_preceding')
        written += save_video_frames(crop(following_frames), root, 'gt_following', counter_start=K + T)
        if gt_middle_frames is not None:
            written += save_video_frames(crop(gt_middle_frames), root, 'gt_middle', counter_start=K)
        written += save_video_frames(crop(gen_output['pred']), root, 'pred_middle', counter_start=K)
        if intermediate_preds:
            for key, prefix in (('pred_forward', 'pred_middle_forward'), ('pred_backward', 'pred_middle_backward'),
                                ('interp_net_outputs_1', 'interp_net_outputs_1'),
                                ('interp_net_outputs_2', 'interp_net_outputs_2')):
                if gen_output.get(key) is not None:
                    written += save_video_frames(crop(gen_output[key]), root, prefix, counter_start=K)
    return written
