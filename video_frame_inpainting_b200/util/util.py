"""The four helpers of the reference's ``src/util/util.py`` that sit on the TAI call path
(util.py:22-41, 188-202); everything else in that file is plotting / IO and out of scope."""
import torch.nn as nn
from torch.nn import init


def inverse_transform(images):
    """[-1, 1] -> [0, 1]   (util.py:22-23)."""
    return (images + 1.) / 2


def fore_transform(images):
    """[0, 1] -> [-1, 1]   (util.py:26-27)."""
    return images * 2 - 1


_GRAY = (0.1140, 0.5870, 0.2989)  # B, G, R weights, util.py:32,39


def bgr2gray(image):
    """[B,3,H,W] (BGR) -> [B,1,H,W]   (util.py:30-34)."""
    return (_GRAY[0] * image[:, 0] + _GRAY[1] * image[:, 1] + _GRAY[2] * image[:, 2]).unsqueeze(1)


def bgr2gray_batched(image):
    """[B,T,3,H,W] (BGR) -> [B,T,1,H,W]   (util.py:37-41)."""
    return (_GRAY[0] * image[:, :, 0] + _GRAY[1] * image[:, :, 1] + _GRAY[2] * image[:, :, 2]).unsqueeze(2)


def weights_init(m):
    """xavier-normal weights / zero bias for (transposed) convolutions, U(0, 0.02) for linear layers
    (util.py:193-202; applied to the generator at environments.py:80 and the discriminator at :284)."""
    from ..discriminators.SNDiscriminator import SNConv2d, SNLinear
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, SNConv2d)):
        init.xavier_normal_(m.weight.data, gain=1)
        if m.bias is not None:
            init.constant_(m.bias.data, 0.0)
    elif isinstance(m, (nn.Linear, SNLinear)):
        init.uniform_(m.weight.data, 0.0, 0.02)
        init.constant_(m.bias.data, 0.0)
    elif isinstance(m, nn.BatchNorm2d):
        init.uniform_(m.weight.data, 1.0, 0.02)
        init.constant_(m.bias.data, 0.0)


def move_to_devices(model):
    """Moves the model to the current CUDA device (util.py:188-190)."""
    return model.cuda()
