"""Generates tests/golden/l2_gdl_ref.npz by IMPORTING the reference's own GDL class
(/root/reference/src/losses/losses.py, unmodified, torch CPU) and torch.nn.MSELoss, spelled as the
reference's training step spells them (environments.py:363-371: inverse_transform, then loss_Lp and
loss_gdl).  Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_loss_golden.py
"""
import importlib.util
import os
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_losses", "/root/reference/src/losses/losses.py")
ref_losses = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_losses)


def main():
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(1234)
    out = {}
    shapes = [(3, 2, 1, 5, 7), (2, 3, 3, 12, 16), (4, 1, 9, 4), (2, 2, 1, 33, 20)]
    gdl = ref_losses.GDL()
    mse = torch.nn.MSELoss()
    for i, shape in enumerate(shapes):
        x = rng.uniform(-1, 1, shape).astype(np.float32)
        y = rng.uniform(-1, 1, shape).astype(np.float32)
        if i == 1:  # ties: identical neighbours make some |.| arguments exactly zero (sign(0) = 0)
            x[..., 3:6, 4:9] = 0.25
            y[..., 3:6, 4:9] = -0.5
        tx = torch.from_numpy(x).requires_grad_()
        ty = torch.from_numpy(y)
        a = (tx + 1.) / 2           # util.py:22-23
        b = (ty + 1.) / 2
        l_mse = mse(a, b)
        l_gdl = gdl(a, b)
        (0.7 * l_mse + 1.3 * l_gdl).backward()
        out["x%d" % i] = x
        out["y%d" % i] = y
        out["mse%d" % i] = np.float64(l_mse.item())
        out["gdl%d" % i] = np.float64(l_gdl.item())
        out["grad%d" % i] = tx.grad.numpy().copy()
    out["g_mse"] = np.float64(0.7)
    out["g_gdl"] = np.float64(1.3)
    out["n"] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(HERE, "l2_gdl_ref.npz"), **out)
    print("wrote", os.path.join(HERE, "l2_gdl_ref.npz"))


if __name__ == "__main__":
    main()
