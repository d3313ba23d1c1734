"""Host-side mirror of the reference model interface: registry keys, constructor signatures, state_dict
names, output contracts.  The hot-path kernels need a GPU; on CPU the same module tree is exercised with
the reference formulation of those operators (oracle/reference_model.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.reference_model import CpuTAITrainingStep, to_cpu_reference
from video_frame_inpainting_b200.losses.losses import GDL
from video_frame_inpainting_b200.models.create_model import create_model
from video_frame_inpainting_b200.models.mcnet.mcnet import DecCnn
from video_frame_inpainting_b200.models.slomo.slomo import SloMoFillInModel
from video_frame_inpainting_b200.models.tai.tai import TAIFillInModel
from video_frame_inpainting_b200.models.twi.twi import TimeWeightedInterpolationFillInModel
from video_frame_inpainting_b200.util.util import weights_init


def test_registry_and_parameter_counts():
    """Parameter counts derived from the reference layer shapes: ~38.3 M for TAI_gray."""
    gray = create_model('TAI_gray')
    assert sum(p.numel() for p in gray.parameters()) == 38320861
    keys = gray.state_dict().keys()
    for expected in ('generator.motion_enc.dyn_conv1.0.weight', 'generator.conv_lstm_cell.conv.weight',
                     'generator.dec_cnn.dec1.2.bias', 'merge_residual1.res.0.weight',
                     'kernelnet.moduleConv.0.0.weight', 'kernelnet.moduleUpsample.3.1.weight',
                     'kernelnet.moduleVertical1.7.weight', 'kernelnet.moduleHorizontal2.4.bias'):
        assert expected in keys, expected
    # num_block = 5: the ratio conv takes one extra channel (tai.py:335-340); num_block = 4 has no such block
    assert gray.state_dict()['kernelnet.moduleUpsample.3.1.weight'].shape[1] == 65
    color = create_model('TAI_color')
    assert all(w.shape[1] % 2 == 0 for k, w in color.state_dict().items() if 'moduleUpsample' in k and k.endswith('1.weight'))
    assert 'mcnet.motion_enc.dyn_conv1.0.weight' in create_model('TimeWeightedInterpolationFillInModel_gray').state_dict()
    assert 'generator.compute_enc.enc1.0.weight' in create_model('SloMoFillInModel_color').state_dict()
    with pytest.raises(RuntimeError):
        create_model('OFFillInModel')


def _tiny(cls=TAIFillInModel, c=1, num_block=5):
    torch.manual_seed(0)
    m = cls(8, c, 3, 5, num_block=num_block, kf_dim=4)
    m.apply(weights_init)
    return to_cpu_reference(m)


@pytest.mark.parametrize("c,num_block", [(1, 5), (3, 4)])
def test_forward_contract_and_time_conditioning(c, num_block):
    m = _tiny(c=c, num_block=num_block)
    B, K, T, F_ = 2, 3, 2, 4
    pre, fol = torch.rand(B, K, c, 32, 32) * 2 - 1, torch.rand(B, F_, c, 32, 32) * 2 - 1
    with torch.no_grad():
        out = m(T, pre, fol)
    assert set(out) == {'pred', 'pred_forward', 'pred_backward', 'interp_net_outputs_1', 'interp_net_outputs_2'}
    for v in out.values():
        assert v.shape == (B, T, c, 32, 32)
    assert torch.allclose(out['pred'], 0.5 * out['interp_net_outputs_1'] + 0.5 * out['interp_net_outputs_2'], atol=1e-6)
    # the ratio plane reaches the network only when num_block >= 5 (tai.py:213 with create_model.py:28,30)
    args = [torch.randn(1, c, 32, 32), torch.randn(1, c, 32, 32)] + [torch.randn(1, 32, 4, 4) for _ in range(4)]
    res = [torch.randn(1, 4, 32, 32), torch.randn(1, 8, 16, 16), torch.randn(1, 16, 8, 8)]
    with torch.no_grad():
        d_a = m.kernelnet(*args, res, ratio=0.25)[0]
        d_b = m.kernelnet(*args, res, ratio=0.75)[0]
    assert torch.equal(d_a, d_b) == (num_block < 5)


def test_twi_blend_weights_and_slomo_order():
    twi = _tiny(TimeWeightedInterpolationFillInModel)
    assert twi.blend_weights(3) == [(0.75, 0.25, 0), (0.5, 0.5, 0), (0.25, 0.75, 0)]
    with torch.no_grad():
        out = twi(3, torch.rand(1, 2, 1, 32, 32), torch.rand(1, 2, 1, 32, 32))
    assert torch.allclose(out['pred'][:, 0], 0.75 * out['interp_net_outputs_1'][:, 0] + 0.25 * out['interp_net_outputs_2'][:, 0], atol=1e-6)
    torch.manual_seed(1)
    s = SloMoFillInModel(4, 3)
    s.apply(weights_init)
    s = to_cpu_reference(s)
    with torch.no_grad():
        out = s(3, torch.rand(1, 2, 3, 64, 64) * 2 - 1, torch.rand(1, 2, 3, 64, 64) * 2 - 1)
    assert out['pred'].shape == (1, 3, 3, 64, 64) and out['F_t_0_collector'].shape == (1, 3, 2, 64, 64)
    # collectors are in REVERSE time order (slomo.py:332-340): entry 0 belongs to t = 3/4
    t = 3 / 4
    assert torch.allclose(out['F_t_0_collector'][:, 0], -(1 - t) * t * out['F_0_1'] + t * t * out['F_1_0'], atol=1e-6)


def test_fixed_unpooling_and_gdl():
    x = torch.arange(12.).view(1, 2, 2, 3)
    up = DecCnn(1, 4).fixed_unpooling(x)
    assert up.shape == (1, 2, 4, 6) and torch.equal(up[:, :, ::2, ::2], x) and up.sum() == x.sum()
    a, b = torch.rand(2, 3, 5, 6), torch.rand(2, 3, 5, 6)
    wa, wb = a[..., :, :-1] - a[..., :, 1:], b[..., :, :-1] - b[..., :, 1:]
    ha, hb = a[..., 1:, :] - a[..., :-1, :], b[..., 1:, :] - b[..., :-1, :]
    expect = ((wa - wb).abs()[..., 1:, :] + (ha - hb).abs()[..., :, 1:]).reshape(2, -1).mean()
    assert torch.allclose(GDL()(a, b), expect)


def test_cpu_reference_training_step_decreases_nothing_weird():
    torch.manual_seed(0)
    st = CpuTAITrainingStep(TAIFillInModel(8, 1, 3, 5, num_block=5, kf_dim=4), (32, 32), 1, 3, 2, 3, df_dim=8)
    pre, fol, gt = (torch.rand(2, n, 1, 32, 32) * 2 - 1 for n in (3, 3, 2))
    before = [p.detach().clone() for p in st.generator.parameters()]
    lg, ld = st.step(pre, fol, gt)
    assert np.isfinite(lg) and np.isfinite(ld)
    changed = sum(int(not torch.equal(a, b)) for a, b in zip(before, st.generator.parameters()))
    # merge_residual1 is computed but never consumed by the kernel net (tai.py:47,93 vs 221-226): no gradient
    assert changed >= len(before) - 4


def _golden_case(z, ci):
    tag = 'c%d_' % ci
    cfg = {k[len(tag) + 4:]: int(z[k]) for k in z.files if k.startswith(tag + 'cfg_')}
    sd = {str(n): torch.from_numpy(z[tag + 'sd_' + str(n)]) for n in z[tag + 'sd_names']}
    model = TAIFillInModel(cfg['gf_dim'], cfg['c_dim'], cfg['feature_size'], cfg['ks'], num_block=cfg['num_block'],
                           kf_dim=cfg['kf_dim'])
    return tag, cfg, sd, model


@pytest.mark.parametrize("ci", [0, 1])
def test_reference_model_classes_golden(ci):
    """tests/golden/tai_model_ref.npz holds a state_dict, inputs, outputs and gradients produced by the REFERENCE's
    own TAIFillInModel / MCNet / TAI classes (imported unmodified by tests/golden/make_model_golden.py, whose
    header lists the shims).  (1) checkpoint compatibility: the reference state_dict loads with strict=True --
    every key and shape matches; (2) the CPU port used as the end-to-end checker (oracle/reference_model.py)
    reproduces the reference's outputs and gradients."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_model_ref.npz"))
    tag, cfg, sd, model = _golden_case(z, ci)
    assert list(model.state_dict().keys()) == list(sd.keys())          # same names in the same order
    model.load_state_dict(sd, strict=True)
    model = to_cpu_reference(model)
    pre, fol = torch.from_numpy(z[tag + 'pre']), torch.from_numpy(z[tag + 'fol'])
    out = model(cfg['T'], pre, fol)
    for k in ('pred', 'pred_forward', 'pred_backward', 'interp_net_outputs_1', 'interp_net_outputs_2'):
        ref = z[tag + 'out_' + k]
        assert out[k].shape == ref.shape
        assert O.rel_err(out[k].detach().numpy(), ref) < 1e-5, k
    (out['pred'].pow(2).mean() + out['pred_forward'].mean() + out['pred_backward'].pow(2).mean()).backward()
    params = dict(model.named_parameters())
    for n in z[tag + 'grad_names']:
        assert O.rel_err(params[str(n)].grad.numpy(), z[tag + 'grad_' + str(n)]) < 1e-4, n


def test_reference_full_configuration_golden_forward():
    """tests/golden/tai_full_config_ref.npz: the reference's own TAIFillInModel(64, 1, 3, 51, num_block=5) -- the
    network registered as TAI_gray (create_model.py:27-28) -- on one 64x64 clip, K = F = T = 5, with weights that are
    a function of the state_dict key names (tests/helpers.py:name_seeded_state_dict; 38.3 M parameters cannot be
    committed).  Here: same keys in the same order, and the CPU port reproduces the five outputs to 1e-5 (forward
    only: the backward pass is another 80 s of CPU time; gradients are checked on the GPU, test_models_gpu.py)."""
    import os
    from tests.helpers import name_seeded_state_dict
    from video_frame_inpainting_b200.models.create_model import create_model
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_full_config_ref.npz"))
    model = create_model('TAI_gray')
    assert list(model.state_dict().keys()) == [str(n) for n in z['sd_names']]
    model.load_state_dict(name_seeded_state_dict(model.state_dict()), strict=True)
    model = to_cpu_reference(model)
    with torch.no_grad():
        out = model(int(z['cfg_T']), torch.from_numpy(z['pre']), torch.from_numpy(z['fol']))
    for k in ('pred', 'pred_forward', 'pred_backward', 'interp_net_outputs_1', 'interp_net_outputs_2'):
        assert out[k].shape == z['out_' + k].shape
        assert O.rel_err(out[k].numpy(), z['out_' + k]) < 1e-5, k


def test_reference_slomo_classes_golden():
    """Same pin for the Super SloMo baseline (the reference's own slomo.py classes): strict state_dict load, the
    five outputs (incl. the reversed time order of the collectors, slomo.py:331-340) and a gradient."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_model_ref.npz"))
    sd = {str(n): torch.from_numpy(z['s_sd_' + str(n)]) for n in z['s_sd_names']}
    model = SloMoFillInModel(gf_dim=2, c_input_dim=3)
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    model = to_cpu_reference(model)
    out = model(3, torch.from_numpy(z['s_pre']), torch.from_numpy(z['s_fol']))
    for k in ('pred', 'F_0_1', 'F_1_0', 'F_t_0_collector', 'F_t_1_collector'):
        assert out[k].shape == z['s_out_' + k].shape
        assert O.rel_err(out[k].detach().numpy(), z['s_out_' + k]) < 1e-5, k
    out['pred'].pow(2).mean().backward()
    name = str(z['s_grad_name'][0])
    assert O.rel_err(dict(model.named_parameters())[name].grad.numpy(), z['s_grad']) < 1e-4


@pytest.mark.parametrize("cls", [TAIFillInModel, TimeWeightedInterpolationFillInModel])
def test_batched_loop_structure_equals_reference_loops_on_cpu(cls):
    """Host logic of the batching steps (two MC-Net streams as one pass, kernel network once over T*B frames, motion
    history as one batch) without a GPU: on the CPU port the batched route must reproduce the reference's loop
    structure (tai.py:77-84, 91-105; mcnet.py:405-409) -- outputs and parameter gradients."""
    torch.manual_seed(1)
    m = cls(4, 1, 3, 5, num_block=5, kf_dim=2)
    m.apply(weights_init)
    g = torch.Generator().manual_seed(2)
    for n, p in m.named_parameters():
        if n.endswith('bias'):
            p.data.uniform_(-0.1, 0.1, generator=g)
    m = to_cpu_reference(m)                        # switches every batching flag off
    mcnet = m.generator
    pre = torch.rand(2, 4, 1, 32, 32, generator=g) * 2 - 1
    fol = torch.rand(2, 4, 1, 32, 32, generator=g) * 2 - 1
    results = []
    for batched in (False, True):
        m.batch_streams = m.batch_time = mcnet.batch_history = batched
        m.zero_grad(set_to_none=True)
        out = m(3, pre, fol)
        (out['pred'].pow(2).mean() + out['pred_forward'].mean() + out['pred_backward'].pow(2).mean()).backward()
        results.append(({k: v.detach().numpy() for k, v in out.items()},
                        np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().ravel()
                                        for p in m.parameters()])))
    for k in results[0][0]:
        assert O.rel_err(results[1][0][k], results[0][0][k]) < 1e-4, k
    assert O.rel_err(results[1][1], results[0][1]) < 1e-3
    m.batch_streams = m.batch_time = mcnet.batch_history = True
    assert m(2, pre, fol[:, :3])['pred'].shape == (2, 2, 1, 32, 32)      # K != F: separate MC-Net passes


def test_fused_sequential_keeps_keys_and_cpu_route():
    """FusedSequential is nn.Sequential for state_dict purposes (the reference's keys) and evaluates CPU tensors with
    the plain children -- the route the CPU port of the reference model takes."""
    import torch.nn as nn
    from video_frame_inpainting_b200.models.layers import FusedSequential
    torch.manual_seed(0)
    mods = [nn.Conv2d(2, 4, 3, padding=1), nn.ReLU(), nn.ConvTranspose2d(4, 3, 3, padding=1), nn.LeakyReLU(0.1),
            nn.Conv2d(3, 1, 1)]
    plain, fused = nn.Sequential(*mods), FusedSequential(*mods)
    assert list(plain.state_dict().keys()) == list(fused.state_dict().keys())
    x = torch.randn(2, 2, 8, 10)
    assert torch.equal(plain(x), fused(x))


def test_reference_sn_discriminator_golden():
    """The spectral-norm discriminator against the reference's own SNDiscriminator classes (tests/golden/
    tai_model_ref.npz, 'd_*'): strict state_dict load, logits of two consecutive calls (the in-place weight
    normalisation and the carried u), the normalised weight and gradients.  The initial u vectors come from the
    fixture (the reference draws them from the global RNG)."""
    import os
    from video_frame_inpainting_b200.discriminators.SNDiscriminator import SNDiscriminator
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_model_ref.npz"))
    sd = {str(n): torch.from_numpy(z['d_sd_' + str(n)]) for n in z['d_sd_names']}
    disc = SNDiscriminator((32, 32), 1, 3, 4, 3)
    assert list(disc.state_dict().keys()) == list(sd.keys())
    disc.load_state_dict(sd, strict=True)
    for i, m in enumerate([m for m in disc.modules() if hasattr(m, 'Ip')]):
        m.u = torch.from_numpy(z['d_u%d' % i])
    video = torch.from_numpy(z['d_video'])
    assert O.rel_err(disc(video).detach().numpy(), z['d_logits_call1']) < 1e-5
    logits2 = disc(video)
    assert O.rel_err(logits2.detach().numpy(), z['d_logits_call2']) < 1e-5
    logits2.sum().backward()
    assert O.rel_err(disc.conv_layers[0].weight.detach().numpy(), z['d_weight0_after']) < 1e-5
    assert O.rel_err(disc.conv_layers[0].weight.grad.numpy(), z['d_grad_weight0']) < 1e-4
    assert O.rel_err(disc.linear_layer.weight.grad.numpy(), z['d_grad_linear']) < 1e-4


def _load_step_golden():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_step_ref.npz"))
    cfg = {k[4:]: float(z[k]) for k in z.files if k.startswith('cfg_')}
    ints = {k: int(cfg[k]) for k in ('gf_dim', 'c_dim', 'feature_size', 'ks', 'num_block', 'kf_dim', 'K', 'T', 'F_', 'H', 'W',
                                     'B', 'df_dim', 'Ip', 'disc_t')}
    gsd = {str(n): torch.from_numpy(z['g_sd_' + str(n)]) for n in z['g_sd_names']}
    dsd = {str(n): torch.from_numpy(z['d_sd_' + str(n)]) for n in z['d_sd_names']}
    return z, cfg, ints, gsd, dsd


STEP_TERMS = ('Lp', 'gdl', 'L_GAN', 'Lp_forward', 'Lp_backward', 'gdl_forward', 'gdl_backward', 'loss_G', 'loss_d_fake',
              'loss_d_real', 'loss_D')


def test_reference_training_step_golden():
    """One training step of the reference's OWN TAITrainingEnvironment (tests/golden/tai_step_ref.npz, produced by
    tests/golden/make_step_golden.py from environments.py with the single token async=True replaced): the CPU port
    of the step must reproduce all eleven loss terms and the parameters after both Adam updates."""
    z, cfg, i, gsd, dsd = _load_step_golden()
    gen = TAIFillInModel(i['gf_dim'], i['c_dim'], i['feature_size'], i['ks'], num_block=i['num_block'], kf_dim=i['kf_dim'])
    st = CpuTAITrainingStep(gen, (i['H'], i['W']), i['c_dim'], i['K'], i['T'], i['F_'], alpha=cfg['alpha'], beta=cfg['beta'],
                            lr=cfg['lr'], beta1=cfg['beta1'], df_dim=i['df_dim'], Ip=i['Ip'], disc_t=i['disc_t'])
    st.generator.load_state_dict(gsd, strict=True)
    st.discriminator.load_state_dict(dsd, strict=True)
    for k, m in enumerate([m for m in st.discriminator.modules() if hasattr(m, 'Ip')]):
        m.u = torch.from_numpy(z['u%d' % k])
    clip = torch.from_numpy(z['clip'])
    K, T = i['K'], i['T']
    st.step(clip[:, :K], clip[:, K + T:], clip[:, K:K + T])
    for name in STEP_TERMS:
        ref = float(z['loss_' + name])
        assert abs(st.terms[name] - ref) <= 1e-5 * abs(ref), (name, st.terms[name], ref)
    gp, dp = dict(st.generator.named_parameters()), dict(st.discriminator.named_parameters())
    for n in z['g_after_names']:
        assert O.rel_err(gp[str(n)].detach().numpy(), z['g_after_' + str(n)]) < 1e-5, n
    for n in z['d_after_names']:
        assert O.rel_err(dp[str(n)].detach().numpy(), z['d_after_' + str(n)]) < 1e-5, n


@pytest.mark.parametrize("tag", ["twi", "mcnet"])
def test_reference_twi_and_mcnet_classes_golden(tag):
    """bi-TWI (src/models/twi/twi.py: per-t blend weights, no time-ratio plane) and the forward-only MC-Net baseline
    (mcnet.py:301-347) against outputs of the reference's own classes: strict state_dict load, every output, a
    gradient."""
    import os
    from video_frame_inpainting_b200.models.mcnet.mcnet import MCNetFillInModel
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "tai_model_ref.npz"))
    sd = {str(n): torch.from_numpy(z[tag + '_sd_' + str(n)]) for n in z[tag + '_sd_names']}
    model = (TimeWeightedInterpolationFillInModel(4, 1, 3, 5, num_block=5, kf_dim=2) if tag == 'twi'
             else MCNetFillInModel(4, 3, 3))
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    model = to_cpu_reference(model)
    out = model(3, torch.from_numpy(z[tag + '_pre']), torch.from_numpy(z[tag + '_fol']))
    keys = [k[len(tag) + 5:] for k in z.files if k.startswith(tag + '_out_')]
    assert sorted(out) == sorted(keys)
    for k in keys:
        assert O.rel_err(out[k].detach().numpy(), z[tag + '_out_' + k]) < 1e-5, k
    out['pred'].pow(2).mean().backward()
    name = str(z[tag + '_grad_name'][0])
    assert O.rel_err(dict(model.named_parameters())[name].grad.numpy(), z[tag + '_grad']) < 1e-4


def test_training_environment_factory_keeps_reference_signature():
    """create_training_environment(...) takes the reference's positional arguments in the reference's order
    (environments.py:24-26) and rejects model families that are out of scope with the reference's error."""
    import inspect
    from video_frame_inpainting_b200.environments import environments as E
    ref_order = ['fill_in_model', 'c_dim', 'checkpoints_dir', 'name', 'max_K', 'max_T', 'max_F', 'image_size', 'alpha',
                 'beta', 'lr', 'beta1', 'df_dim', 'Ip', 'disc_window_size', 'tf_p_min', 'tf_p_max', 'tf_offset',
                 'tf_decay', 'padding_size', 'lambda_r', 'lambda_p', 'lambda_w', 'lambda_s', 'lr_decay_count',
                 'lr_decay_rate']
    assert list(inspect.signature(E.create_training_environment).parameters)[:len(ref_order)] == ref_order
    with pytest.raises(RuntimeError, match='unsupported type'):
        E.create_training_environment(object(), 1, '/tmp', 'x', 2, 2, 2, (32, 32), 1.0, 0.02, 1e-4, 0.5, 8, 3, 3)
    for cls in ('BaseVideoFillInEnvironment', 'BaseTrainingEnvironment', 'L2GDLDiscTrainingEnvironment',
                'MCNetTrainingEnvironment', 'TAITrainingEnvironment', 'SloMoTrainingEnvironment'):
        assert hasattr(E, cls)
    np.random.seed(0)
    env = E.MCNetTrainingEnvironment.__new__(E.MCNetTrainingEnvironment)
    env.max_K, env.max_T, env.max_F = 5, 4, 3
    for _ in range(50):
        K, T, F_ = env.sample_KTF(True)
        assert 2 <= K <= 5 and 1 <= T <= 4 and 1 <= F_ <= 3
    assert env.sample_KTF(False) == (5, 4, 3)
