"""Training / evaluation step of the TAI path (callers of the hot path).

Restates ``src/environments/environments.py`` of the reference for torch 2 -- the classes on the TAI path
only (``BaseVideoFillInEnvironment`` :64-119, ``BaseTrainingEnvironment`` :122-259,
``L2GDLDiscTrainingEnvironment`` :262-397, ``TAITrainingEnvironment`` :415-485, ``SloMoTrainingEnvironment``
:523-620) with the same method
names, loss composition and checkpoint dictionary.  Differences, all on the host side:

* ``Variable`` / ``volatile`` / ``.cuda(async=True)`` (a SyntaxError on Python 3, environments.py:94)
  become plain tensors, ``torch.no_grad()`` and ``.cuda(non_blocking=True)``;
* gradients live in one flat buffer per network and are all-reduced bucket by bucket while backward is
  still running (``parallel.FlatGradAllReducer``) when the process group has more than one rank -- the
  reference is single-GPU;
* the reconstruction losses (MSELoss + GDL on the inverse-transformed frames, three prediction tensors per
  step) go through the fused loss kernels (``losses.L2GDLLoss``); every other kernel is reached through
  ``generator(T, preceding, following)``.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .. import ops
from ..discriminators.SNDiscriminator import SNDiscriminator
from ..losses.losses import GDL, L2GDLLoss
from ..parallel import FlatGradAllReducer, broadcast_module
from ..util.util import inverse_transform, move_to_devices, weights_init


def create_training_environment(fill_in_model, c_dim, checkpoints_dir, name, max_K, max_T, max_F, image_size, alpha,
                                beta, lr, beta1, df_dim, Ip, disc_window_size, tf_p_min=None, tf_p_max=None,
                                tf_offset=None, tf_decay=None, padding_size=(0, 0), lambda_r=0.8, lambda_p=0.005,
                                lambda_w=0.4, lambda_s=1.0, lr_decay_count=200, lr_decay_rate=0.1,
                                vgg16_state_dict=None):
    """Factory with the reference's argument order (environments.py:24-52): dispatches on the model family -- bi-TAI /
    bi-TWI -> ``TAITrainingEnvironment``, MC-Net -> ``MCNetTrainingEnvironment``, Super SloMo ->
    ``SloMoTrainingEnvironment``.  The teacher-forcing arguments belong to the self-attention models (out of scope)
    and are accepted and ignored.  Resumes from model_latest.ckpt if present."""
    from ..models.mcnet.mcnet import MCNetFillInModel
    from ..models.slomo.slomo import SloMoFillInModel
    from ..models.tai.tai import TAIFillInModel
    if isinstance(fill_in_model, TAIFillInModel):          # TimeWeightedInterpolationFillInModel derives from it
        env = TAITrainingEnvironment(fill_in_model, checkpoints_dir, name, image_size, c_dim, alpha, beta, lr, beta1,
                                     df_dim, Ip, disc_window_size, max_K, max_T, max_F, padding_size)
    elif isinstance(fill_in_model, MCNetFillInModel):
        env = MCNetTrainingEnvironment(fill_in_model, checkpoints_dir, name, image_size, c_dim, alpha, beta, lr, beta1,
                                       df_dim, Ip, disc_window_size, max_K, max_T, max_F, padding_size)
    elif isinstance(fill_in_model, SloMoFillInModel):
        env = SloMoTrainingEnvironment(fill_in_model, checkpoints_dir, name, lr, beta1, max_K, max_T, max_F, padding_size,
                                       lambda_r, lambda_p, lambda_w, lambda_s, lr_decay_count, lr_decay_rate,
                                       vgg16_state_dict=vgg16_state_dict)
    else:
        raise RuntimeError('Tried to create a training environment for object of unsupported type %s'
                           % type(fill_in_model).__name__)
    if os.path.isfile(os.path.join(checkpoints_dir, name, 'model_latest.ckpt')):
        env.load('model_latest.ckpt')
    return env


def create_eval_environment(fill_in_model, checkpoints_dir, name, snapshot_file_name, padding_size=(0, 0)):
    env = BaseVideoFillInEnvironment(fill_in_model, checkpoints_dir, name, padding_size)
    if snapshot_file_name is not None:
        env.load(snapshot_file_name)
    return env


class BaseVideoFillInEnvironment(object):
    """Owns the generator; ``set_test_inputs`` / ``forward_test``   (environments.py:64-119)."""

    def __init__(self, video_fill_in_model, checkpoints_dir, name, padding_size):
        self.save_dir = os.path.join(checkpoints_dir, name)
        self.padding_size = padding_size
        self.generator = move_to_devices(video_fill_in_model)
        self.generator.apply(weights_init)
        self.K = self.T = self.F = None

    def set_test_inputs(self, preceding_frames, following_frames):
        self.preceding_frames = preceding_frames.contiguous().cuda(non_blocking=True)
        self.following_frames = following_frames.contiguous().cuda(non_blocking=True)

    def set_gt_middle_frames_test(self, gt_middle_frames):
        self.gt_middle_frames = gt_middle_frames.contiguous().cuda(non_blocking=True)

    def forward_test(self):
        if getattr(self, '_graphs', None) is not None:
            return self._forward_test_graphed()
        with torch.no_grad():
            self.gen_output = self.generator(self.T, self.preceding_frames, self.following_frames)

    # -- CUDA-graph replay of the inference forward --------------------------
    # At batch 1 the forward pass is a few thousand kernels of a few microseconds each and the step time is
    # the host's launch time, not the GPU's.  The whole ``generator(T, preceding, following)`` call is
    # captured once per (T, input shapes) into a CUDA graph over static input / output buffers and replayed;
    # the kernels of this library take device pointers, sizes and a stream only, allocate nothing and never
    # synchronise, so they are captured like any other launch.
    def enable_cuda_graph(self, enabled=True):
        self._graphs = {} if enabled else None

    def _forward_test_graphed(self):
        pre, fol = self.preceding_frames, self.following_frames
        key = (int(self.T), tuple(pre.shape), tuple(fol.shape), pre.device.index)
        entry = self._graphs.get(key)
        if entry is None:
            s_pre, s_fol = pre.clone(), fol.clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):  # cuDNN autotuning, lazy initialisation of this library's launchers
                    self.generator(self.T, s_pre, s_fol)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                out = self.generator(self.T, s_pre, s_fol)
            entry = (graph, s_pre, s_fol, out)
            self._graphs[key] = entry
        graph, s_pre, s_fol, out = entry
        s_pre.copy_(pre, non_blocking=True)
        s_fol.copy_(fol, non_blocking=True)
        graph.replay()
        self.gen_output = out  # static buffers: overwritten by the next replay of the same shape

    def load(self, snapshot_file_name):
        save_path = os.path.join(self.save_dir, snapshot_file_name)
        if not os.path.isfile(save_path):
            raise RuntimeError('Failed to find snapshot at path %s' % save_path)
        # The reference writes its checkpoints with torch 0.3.1 / Python 2 and keeps numpy scalars in them
        # (train.py:163 passes np.sum(np.mean(...)) as sum_avg_psnr_err): the weights_only unpickler of
        # torch >= 2.6 rejects those, and py2 pickles need latin1.  Checkpoints are trusted input here, as
        # they are for the reference's torch.load (environments.py:104).
        try:
            snapshot = torch.load(save_path, map_location='cuda', weights_only=False)
        except UnicodeDecodeError:
            snapshot = torch.load(save_path, map_location='cuda', weights_only=False, encoding='latin1')
        self.generator.load_state_dict(snapshot['generator'])
        return snapshot

    def eval(self):
        self.generator.eval()


class BaseTrainingEnvironment(BaseVideoFillInEnvironment):
    def __init__(self, fill_in_model, checkpoints_dir, name, lr, beta1, max_K, max_T, max_F, padding_size):
        super(BaseTrainingEnvironment, self).__init__(fill_in_model, checkpoints_dir, name, padding_size)
        self.start_update = 0
        self.total_updates = 0
        self.start_sum_avg_psnr_err = 0
        self.start_sum_avg_ssim_err = 0
        self.max_K, self.max_T, self.max_F = max_K, max_T, max_F
        broadcast_module(self.generator)
        self.reducer_G = FlatGradAllReducer(self.generator)
        self.optimizer_G = torch.optim.Adam(self.generator.parameters(), lr=lr, betas=(beta1, 0.999))

    def sample_KTF(self, allow_random_sampling):
        if allow_random_sampling:
            ktf = [np.random.randint(1, self.max_K + 1), np.random.randint(1, self.max_T + 1),
                   np.random.randint(1, self.max_F + 1)]
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # every rank must run the same graph (K switches the batched-history path, T the number of
                # sepconv launches): rank 0's draw wins (the reference is single-process, environments.py:417-427)
                t = torch.tensor(ktf, device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
                dist.broadcast(t, 0)
                ktf = [int(v) for v in t.tolist()]
            return tuple(ktf)
        return self.max_K, self.max_T, self.max_F

    def set_train_inputs(self, preceding_frames, following_frames, gt_middle_frames):
        self.preceding_frames = preceding_frames.contiguous().cuda(non_blocking=True)
        self.following_frames = following_frames.contiguous().cuda(non_blocking=True)
        self.gt_middle_frames = gt_middle_frames.contiguous().cuda(non_blocking=True)

    def forward_train(self):
        self.gen_output = self.generator(self.T, self.preceding_frames, self.following_frames)

    def get_current_state_dict(self, total_updates, sum_avg_psnr_err, sum_avg_ssim_err):
        return {
            'updates': int(total_updates),
            'sum_avg_psnr_err': float(sum_avg_psnr_err),   # numpy scalars in the reference's train loop (train.py:163)
            'sum_avg_ssim_err': float(sum_avg_ssim_err),
            'generator': self.generator.state_dict(),
            'optimizer_G': self.optimizer_G.state_dict(),
        }

    def load(self, snapshot_file_name):
        snapshot = super(BaseTrainingEnvironment, self).load(snapshot_file_name)
        self.start_update = snapshot['updates']
        self.start_sum_avg_psnr_err = snapshot['sum_avg_psnr_err']
        self.start_sum_avg_ssim_err = snapshot['sum_avg_ssim_err']
        self.optimizer_G.load_state_dict(snapshot['optimizer_G'])
        return snapshot

    def save(self, snapshot_file_name, total_updates, sum_avg_psnr_err, sum_avg_ssim_err):
        """environments.py:186-194.  Replicas are bit-identical, so rank 0 alone writes (to a temporary file
        moved into place: a reader never sees a torn checkpoint) and the other ranks wait for it."""
        multi = dist.is_available() and dist.is_initialized()
        if not multi or dist.get_rank() == 0:
            os.makedirs(self.save_dir, exist_ok=True)
            path = os.path.join(self.save_dir, snapshot_file_name)
            tmp = path + '.tmp.%d' % os.getpid()
            torch.save(self.get_current_state_dict(total_updates, sum_avg_psnr_err, sum_avg_ssim_err), tmp)
            os.replace(tmp, path)
        if multi:
            dist.barrier()

    def optimize_parameters(self):
        """One generator update (environments.py:222-228)."""
        self.reducer_G.zero_grad()
        self.compute_loss_G()
        self.reducer_G.arm()
        self.loss_G.backward()
        self.reducer_G.finish()
        self.optimizer_G.step()

    def compute_loss_G(self):
        self.loss_G = torch.zeros(1, device='cuda')

    def get_current_errors(self):
        return {'G_loss': float(self.loss_G.detach())}

    def train(self):
        self.generator.train()


class L2GDLDiscTrainingEnvironment(BaseTrainingEnvironment):
    """L2 + GDL + adversarial loss; spectral-norm discriminator with its own Adam   (environments.py:262-397)."""

    def __init__(self, fill_in_model, checkpoints_dir, name, image_size, c_dim, alpha, beta, lr, beta1, df_dim, Ip,
                 disc_t, max_K, max_T, max_F, padding_size):
        super(L2GDLDiscTrainingEnvironment, self).__init__(fill_in_model, checkpoints_dir, name, lr, beta1, max_K,
                                                           max_T, max_F, padding_size)
        self.loss_Lp = torch.nn.MSELoss()
        self.loss_gdl = GDL()
        self.loss_Lp_gdl = L2GDLLoss()   # both of the above in one kernel pass (used by compute_loss_G)
        self.loss_d = torch.nn.BCEWithLogitsLoss()
        self.alpha, self.beta = alpha, beta
        self.disc_t = disc_t
        discriminator = SNDiscriminator((image_size[0] + padding_size[0], image_size[1] + padding_size[1]), c_dim,
                                        disc_t, df_dim, Ip)
        self.discriminator = move_to_devices(discriminator)
        self.discriminator.apply(weights_init)
        broadcast_module(self.discriminator)
        self.reducer_D = FlatGradAllReducer(self.discriminator)
        self.optimizer_D = torch.optim.Adam(self.discriminator.parameters(), lr=lr, betas=(beta1, 0.999))

    def get_current_state_dict(self, total_updates, sum_avg_psnr_err, sum_avg_ssim_err):
        state = super(L2GDLDiscTrainingEnvironment, self).get_current_state_dict(total_updates, sum_avg_psnr_err,
                                                                                 sum_avg_ssim_err)
        state['discriminator'] = self.discriminator.state_dict()
        state['optimizer_D'] = self.optimizer_D.state_dict()
        return state

    def load(self, snapshot_file_name):
        snapshot = super(L2GDLDiscTrainingEnvironment, self).load(snapshot_file_name)
        self.discriminator.load_state_dict(snapshot['discriminator'])
        self.optimizer_D.load_state_dict(snapshot['optimizer_D'])
        return snapshot

    def create_fake_labels(self):
        """Windows that contain only real (preceding / following) frames are labelled 1 (environments.py:308-323)."""
        n = self.K + self.T + self.F - self.disc_t + 1
        ones_P = max(0, self.K - self.disc_t + 1)
        ones_F = max(0, self.F - self.disc_t + 1)
        labels = torch.zeros(n)
        labels[:ones_P] = 1
        if ones_F > 0:
            labels[n - ones_F:] = 1
        return labels

    def _video(self, middle):
        return torch.cat([self.preceding_frames, middle, self.following_frames], dim=1)

    def compute_loss_D(self):
        h = self.discriminator(self._video(self.gen_output['pred']).detach())
        fake_labels = self.create_fake_labels().to(h.device).view(1, -1).expand(h.size(0), -1)
        self.loss_d_fake = self.loss_d(h, fake_labels)
        h_ = self.discriminator(self._video(self.gt_middle_frames).detach())
        self.loss_d_real = self.loss_d(h_, torch.ones_like(h_))
        self.loss_D = self.loss_d_fake + self.loss_d_real

    def optimize_parameters(self):
        super(L2GDLDiscTrainingEnvironment, self).optimize_parameters()
        self.reducer_D.zero_grad()
        self.compute_loss_D()
        self.reducer_D.arm()
        self.loss_D.backward()
        self.reducer_D.finish()
        self.optimizer_D.step()

    @staticmethod
    def _time_major01(frames):
        """[B,T,C,H,W] in [-1,1] -> [T*B,C,H,W] in [0,1], same-time frames grouped (environments.py:363-369)."""
        _, _, c, H, W = frames.shape
        return inverse_transform(frames.permute(1, 0, 2, 3, 4).contiguous().view(-1, c, H, W))

    def compute_loss_G(self):
        super(L2GDLDiscTrainingEnvironment, self).compute_loss_G()
        # environments.py:363-371: time-major regrouping + inverse_transform + MSELoss + GDL, fused
        self.Lp, self.gdl = self.loss_Lp_gdl(self.gen_output['pred'], self.gt_middle_frames)
        h = self.discriminator(self._video(self.gen_output['pred']))
        self.L_GAN = self.loss_d(h, torch.ones_like(h))
        self.loss_G = self.loss_G + self.alpha * (self.Lp + self.gdl) + self.beta * self.L_GAN

    def get_current_errors(self):
        errors = super(L2GDLDiscTrainingEnvironment, self).get_current_errors()
        errors.update(G_Lp=float(self.Lp.detach()), G_gdl=float(self.gdl.detach()),
                      D_real=float(self.loss_d_real.detach()), D_fake=float(self.loss_d_fake.detach()),
                      G_GAN=float(self.L_GAN.detach()))
        return errors

    def train(self):
        super(L2GDLDiscTrainingEnvironment, self).train()
        self.discriminator.train()


class MCNetTrainingEnvironment(L2GDLDiscTrainingEnvironment):
    """L2 + GDL + adversarial step for the forward-only MC-Net baseline (environments.py:398-411): at least two
    preceding frames (one difference frame) when K, T, F are sampled."""

    def sample_KTF(self, allow_random_sampling):
        if allow_random_sampling:
            return (np.random.randint(2, self.max_K + 1), np.random.randint(1, self.max_T + 1),
                    np.random.randint(1, self.max_F + 1))
        return self.max_K, self.max_T, self.max_F


class TAITrainingEnvironment(L2GDLDiscTrainingEnvironment):
    """Adds the reconstruction losses of the two intermediate predictions   (environments.py:415-485)."""

    def sample_KTF(self, allow_random_sampling):
        if allow_random_sampling:
            return (np.random.randint(2, self.max_K + 1), np.random.randint(1, self.max_T + 1),
                    np.random.randint(2, self.max_F + 1))
        return self.max_K, self.max_T, self.max_F

    def compute_loss_G(self):
        super(TAITrainingEnvironment, self).compute_loss_G()
        # environments.py:437-451
        self.Lp_forward, self.gdl_forward = self.loss_Lp_gdl(self.gen_output['pred_forward'], self.gt_middle_frames)
        self.Lp_backward, self.gdl_backward = self.loss_Lp_gdl(self.gen_output['pred_backward'], self.gt_middle_frames)
        self.loss_G = self.loss_G + self.alpha * (self.Lp_forward + self.Lp_backward + self.gdl_forward
                                                  + self.gdl_backward)

    def get_current_errors(self):
        errors = super(TAITrainingEnvironment, self).get_current_errors()
        errors.update(G_Lp_forward=float(self.Lp_forward.detach()), G_gdl_forward=float(self.gdl_forward.detach()),
                      G_Lp_backward=float(self.Lp_backward.detach()), G_gdl_backward=float(self.gdl_backward.detach()))
        return errors


class SloMoTrainingEnvironment(BaseTrainingEnvironment):
    """Super SloMo training step: reconstruction (L1) + perceptual (VGG-16 conv4_3 features, MSE) + warping (L1 of
    six-plus-2T backward warps) + smoothness (GDL of the two flows against zero)   (environments.py:523-620).

    On this library's kernels: every warp goes through ``FlowWarper`` (``flow_warp_{forward,backward}_b200``, the
    reference builds a meshgrid on the host and calls ``grid_sample`` per warp, slomo.py:265-286) and the two
    smoothness terms through the fused MSE + GDL kernel with an identity transform (``add = 0, mul = 1``; its MSE
    output is unused).  ``vgg16_state_dict``: the reference loads torchvision's ImageNet weights
    (``vgg16(pretrained=True)``, environments.py:532); there is no network here, so the caller passes the
    state_dict (or None: random features -- enough to exercise and time the step, not to train a model)."""

    def __init__(self, fill_in_model, checkpoints_dir, name, lr, beta1, max_K, max_T, max_F, padding_size, lambda_r,
                 lambda_p, lambda_w, lambda_s, lr_decay_count, lr_decay_rate, vgg16_state_dict=None):
        super(SloMoTrainingEnvironment, self).__init__(fill_in_model, checkpoints_dir, name, lr, beta1, max_K, max_T,
                                                       max_F, padding_size)
        import torchvision
        from ..models.slomo.slomo import FlowWarper
        self.l1_loss = torch.nn.L1Loss()
        self.MSE_loss = torch.nn.MSELoss()
        self.gdl = GDL()
        vgg16 = torchvision.models.vgg16(weights=None)
        if vgg16_state_dict is not None:
            vgg16.load_state_dict(vgg16_state_dict)
        elif lambda_p != 0:
            import warnings
            warnings.warn("SloMoTrainingEnvironment: no vgg16_state_dict given -- the perceptual loss (lambda_p=%g) "
                          "is computed on RANDOMLY INITIALISED VGG-16 features; the reference uses the ImageNet "
                          "weights (environments.py:541).  Pass torchvision's vgg16 state_dict for real training."
                          % lambda_p)
        self.vgg16_conv = torch.nn.Sequential(*list(vgg16.features.children())[:22]).cuda()   # features[:22]: up to conv4_3, no ReLU (as the reference)
        for param in self.vgg16_conv.parameters():
            param.requires_grad = False
        self.warper = FlowWarper()
        self.reuse_warps = True   # warping loss: take the per-t warped frames from the generator's stage kernel
        self.lambda_r, self.lambda_p, self.lambda_w, self.lambda_s = lambda_r, lambda_p, lambda_w, lambda_s
        self.lr_decay_count, self.lr_decay_rate, self.lr = lr_decay_count, lr_decay_rate, lr

    def _smoothness(self, flow):
        # GDL(flow, 0) (environments.py:589-590) through the fused loss kernel: identity transform, GDL output
        return ops.l2_gdl_loss(flow.contiguous(), torch.zeros_like(flow), 0.0, 1.0)[1]

    def compute_loss_G(self):
        super(SloMoTrainingEnvironment, self).compute_loss_G()
        B, T, c_dim, H, W = self.gt_middle_frames.shape
        I0 = self.preceding_frames[:, -1].contiguous()
        I1 = self.following_frames[:, 0].contiguous()
        out = self.gen_output
        pred, F_0_1, F_1_0 = out['pred'], out['F_0_1'], out['F_1_0']
        gt = self.gt_middle_frames
        self.reconstruction_loss = self.l1_loss(pred, gt)
        # perceptual loss: the T frames as one batch (the reference loops over t, environments.py:574-581; the
        # feature extractor is frozen and stateless, so the result is the same)
        as_rgb = lambda v: (v.expand(B, T, 3, H, W) if c_dim == 1 else v).reshape(B * T, 3, H, W)
        self.perceptual_loss = self.MSE_loss(self.vgg16_conv(as_rgb(pred)), self.vgg16_conv(as_rgb(gt)).detach())
        # warping loss (environments.py:584-586)
        slomo = getattr(getattr(self.generator, 'module', self.generator), 'generator', None)
        X = getattr(slomo, 'last_interp_input', None) if self.reuse_warps else None
        if X is not None and X.shape[0] == T * B and X.shape[1] == 4 * c_dim + 4:
            # g(I0, F_t_0) and g(I1, F_t_1) were formed by the time-batched stage kernel as channels [C, 2C) and
            # [2C+4, 3C+4) of the refinement input (sample t*B + b); collector index i is time step T-1-i.  Same
            # values bit for bit as warping again; the gradient reaches the flows through that kernel's adjoint.
            blk = lambda i: X[(T - 1 - i) * B:(T - i) * B]
            per_t = [self.l1_loss(blk(i)[:, c_dim:2 * c_dim], gt[:, i])
                     + self.l1_loss(blk(i)[:, 2 * c_dim + 4:3 * c_dim + 4], gt[:, i]) for i in range(T)]
        else:
            per_t = [self.l1_loss(self.warper(I0, out['F_t_0_collector'][:, i].contiguous()), gt[:, i])
                     + self.l1_loss(self.warper(I1, out['F_t_1_collector'][:, i].contiguous()), gt[:, i]) for i in range(T)]
        self.warping_loss = (self.l1_loss(self.warper(I0, F_1_0), I1) + self.l1_loss(self.warper(I1, F_0_1), I0)
                             + sum(per_t) / len(per_t))
        self.smooth_loss = self._smoothness(F_1_0) + self._smoothness(F_0_1)
        self.loss_G = (self.loss_G + self.lambda_r * self.reconstruction_loss + self.lambda_p * self.perceptual_loss
                       + self.lambda_w * self.warping_loss + self.lambda_s * self.smooth_loss)

    def get_current_errors(self):
        errors = super(SloMoTrainingEnvironment, self).get_current_errors()
        errors.update(reconstruction_loss=float(self.reconstruction_loss.detach()),
                      perceptual_loss=float(self.perceptual_loss.detach()),
                      warping_loss=float(self.warping_loss.detach()), smooth_loss=float(self.smooth_loss.detach()))
        return errors

    def optimize_parameters(self):
        """Learning-rate step decay, then one generator update (environments.py:609-616)."""
        for param_group in self.optimizer_G.param_groups:
            param_group['lr'] = self.lr * (self.lr_decay_rate ** (self.total_updates // self.lr_decay_count))
        super(SloMoTrainingEnvironment, self).optimize_parameters()
