"""CPU tests pinning the oracle: against committed golden vectors (made from the live reference by
tests/golden/make_golden.py) and, where /root/reference exists, against the reference itself."""
import copy
import os
import warnings

import pytest
import torch

from oracle import live_reference as lr, nets, step


def _sample(t, k=257):
    f = t.detach().reshape(-1).float()
    return f[:: max(1, f.numel() // k)][:k]


def _rel(a, b):
    return float((a.detach() - b).norm() / b.norm().clamp_min(1e-12))


@pytest.fixture(scope="module")
def gold_nets(golden_dir):
    return torch.load(os.path.join(golden_dir, "golden_nets_n2.pt"))


@pytest.fixture(scope="module")
def gold_step(golden_dir):
    return torch.load(os.path.join(golden_dir, "golden_step_n2.pt"))


def test_oracle_networks_match_golden(gold_nets):
    g = gold_nets
    state = nets.init_model_state(seed=g["seed_w"], perturb=g["perturb"])
    a, b, z = step.synthetic_batch(g["n"], seed=g["seed_x"])
    om = step.OracleModel(state=state)
    gn = g["nets"]
    ar = a.clone().requires_grad_(True); zr = z.clone().requires_grad_(True)
    y = om.G_A_B(ar, zr)
    assert _rel(y, gn["G_A_B"]["out"]) < 1e-5
    (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
    assert _rel(_sample(ar.grad), gn["G_A_B"]["dx"]) < 1e-4
    assert _rel(zr.grad, gn["G_A_B"]["dz"]) < 1e-4
    for k, v in om.params("netG_A_B"):
        ref = gn["G_A_B"]["dw_norm"][k]
        if nets.is_noise_grad("netG_A_B", k):
            assert float(v.grad.norm()) < 1e-3 and ref < 1e-3, k
            continue
        assert abs(float(v.grad.norm()) - ref) <= 1e-4 * ref + 1e-6, k
    assert _rel(om.G_B_A(b), gn["G_B_A"]["out"]) < 1e-5
    assert _rel(om.D_A(a), gn["D_A"]["out"]) < 1e-5
    assert _rel(om.D_B(b), gn["D_B"]["out"]) < 1e-5
    mu, lv = om.E_B(torch.cat((a, b), 1))
    assert _rel(mu, gn["E_B"]["mu"]) < 1e-5 and _rel(lv, gn["E_B"]["logvar"]) < 1e-5
    assert _rel(om.nets["netE_B"]["conv_modules.3.running_mean"], gn["E_B"]["running_mean_3"]) < 1e-5
    assert _rel(om.nets["netE_B"]["conv_modules.12.running_var"], gn["E_B"]["running_var_12"]) < 1e-5
    assert _rel(om.D_z_B(z), gn["D_z_B"]["out"]) < 1e-5


def test_oracle_train_instance_matches_golden(gold_step):
    g = gold_step
    state = nets.init_model_state(seed=g["seed_w"], perturb=g["perturb"])
    a, b, z = step.synthetic_batch(g["n"], seed=g["seed_x"])
    om = step.OracleModel(state=state)
    for it, rec in enumerate(g["steps"]):
        losses, visuals, gnorms = om.train_instance(a, b, z)
        # step 2 runs on Adam-updated weights; Adam turns rounding-noise gradients (biases that feed
        # an instance norm) into +-lr moves, so tolerances loosen after the first update.
        tol = 2e-5 if it == 0 else 2e-3
        for k, v in rec["losses"].items():
            assert abs(losses[k] - v) <= tol * max(1.0, abs(v)), (it, k, losses[k], v)
        for k, v in rec["gnorms"].items():
            assert abs(gnorms[k] - v) <= max(tol, 1e-4) * max(1.0, abs(v)), (it, k, gnorms[k], v)
        for k, v in rec["visuals"].items():
            assert _rel(_sample(visuals[k], 1025), v) < (1e-4 if it == 0 else 1e-2), (it, k)
        if it == 0:
            for key, v in rec["grad_norm"].items():
                name, k = key.split("/")
                gr = om.nets[name][k].grad
                if nets.is_noise_grad(name, k):
                    assert float(gr.norm()) < 1e-3 and v < 1e-3, key
                    continue
                assert abs(float(gr.double().norm()) - v) <= 1e-3 * v + 1e-6, key


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_step():
    warnings.simplefilter("ignore")
    opt = step.default_opt()
    state = nets.init_model_state(seed=7, perturb=0.03)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    om = step.OracleModel(opt, state)
    a, b, z = step.synthetic_batch(3, seed=99)
    lo, vo, go = om.train_instance(a, b, z)
    lr_, vr, gr = ref.train_instance(a, b, z)
    for k in lo:
        assert abs(lo[k] - float(lr_[k])) < 2e-5 * max(1.0, abs(lo[k])), k
    for k in vo:
        assert (vo[k] - vr[k]).abs().max() < 1e-4, k
    for k in go:
        assert abs(go[k] - float(gr[k])) < 1e-4 * max(1.0, abs(go[k])), k
    # G-side gradients are still in .grad after the step on both sides
    for name in ("netG_A_B", "netG_B_A", "netE_B"):
        rp = dict(getattr(ref, name).named_parameters())
        for k, v in om.params(name):
            if v.grad is None:
                assert rp[k].grad is None
                continue
            if nets.is_noise_grad(name, k):
                assert float(v.grad.norm()) < 1e-3 and float(rp[k].grad.norm()) < 1e-3, (name, k)
                continue
            scale = float(v.grad.norm()) + 1e-12
            assert float((v.grad - rp[k].grad).norm()) <= 1e-3 * scale + 1e-6, (name, k)


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
def test_state_dict_keys_match_reference():
    warnings.simplefilter("ignore")
    opt = step.default_opt()
    state = nets.init_model_state(seed=1)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    for name in nets.NET_NAMES:
        full = set(getattr(ref, name).state_dict().keys())
        ours = set(state[name].keys())
        assert ours <= full
        for k in full - ours:   # only the CINResnetBlock alias keys may be missing
            assert name == "netG_A_B" and k.split(".")[1] in ("10", "11", "12"), (name, k)


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("ignore_noise", [False, True])
def test_oracle_stoch_matches_live_reference_step(ignore_noise):
    """StochCycleGAN.train_instance (model.py:126-208) at 64x64 with the climate channel counts (3 -> 1)."""
    warnings.simplefilter("ignore")
    opt = step.default_opt(output_nc=1)
    state = nets.init_model_state(seed=17, perturb=0.03, output_nc=1)
    ref = lr.build_reference_model(copy.deepcopy(opt), state, stoch=True, ignore_noise=ignore_noise)
    om = step.OracleStochModel(opt, state, ignore_noise=ignore_noise)
    a, b, z = step.synthetic_batch(2, seed=5, output_nc=1, kind="climate")
    for it in range(2):
        lo, vo, go = om.train_instance(a, b, z)
        lr_, vr, gr = ref.train_instance(a, b, z)
        tol = 2e-5 if it == 0 else 2e-3
        assert list(lo.keys()) == list(lr_.keys()) and list(go.keys()) == list(gr.keys())
        for k in lo:
            assert abs(lo[k] - float(lr_[k])) < tol * max(1.0, abs(lo[k])), (it, k)
        for k in vo:
            assert (vo[k] - vr[k]).abs().max() < (1e-4 if it == 0 else 1e-2), (it, k)
        for k in go:
            assert abs(go[k] - float(gr[k])) < max(tol, 1e-4) * max(1.0, abs(go[k])), (it, k)


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
def test_oracle_supervised_matches_live_reference_step():
    """AugmentedCycleGAN.supervised_train_instance (model.py:541-604)."""
    warnings.simplefilter("ignore")
    opt = step.default_opt()
    state = nets.init_model_state(seed=23, perturb=0.03)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    om = step.OracleModel(opt, state)
    a, b, z = step.synthetic_batch(3, seed=31)
    for it in range(2):
        lo = om.supervised_train_instance(a, b, z)
        lr_ = ref.supervised_train_instance(a, b, z)
        assert list(lo.keys()) == list(lr_.keys())
        tol = 2e-5 if it == 0 else 5e-3
        for k in lo:
            assert abs(lo[k] - float(lr_[k])) < tol * max(1.0, abs(lo[k])), (it, k, lo[k], float(lr_[k]))
    for name in ("netG_A_B", "netG_B_A", "netE_B", "netD_z_B", "netD_B"):
        rp = dict(getattr(ref, name).named_parameters())
        for k, v in om.params(name):
            assert float((v.detach() - rp[k].detach()).norm()) <= 2e-3 * float(v.detach().norm()) + 1e-5, (name, k)


def test_oracle_stoch_matches_golden(golden_dir):
    """golden_stoch_n2.pt: live reference StochCycleGAN on config 3's shape (climate fields 3 -> 1, 128x128)."""
    g = torch.load(os.path.join(golden_dir, "golden_stoch_n2.pt"))
    opt = step.default_opt(output_nc=g["output_nc"])
    state = nets.init_model_state(seed=g["seed_w"], perturb=g["perturb"], output_nc=g["output_nc"])
    a, b, z = step.synthetic_batch(g["n"], size=g["size"], seed=g["seed_x"], output_nc=g["output_nc"], kind=g["kind"])
    om = step.OracleStochModel(opt, state)
    for it, rec in enumerate(g["steps"]):
        losses, visuals, gnorms = om.train_instance(a, b, z)
        tol = 2e-5 if it == 0 else 2e-3
        for k, v in rec["losses"].items():
            assert abs(losses[k] - v) <= tol * max(1.0, abs(v)), (it, k, losses[k], v)
        for k, v in rec["gnorms"].items():
            assert abs(gnorms[k] - v) <= max(tol, 1e-4) * max(1.0, abs(v)), (it, k, gnorms[k], v)
        for k, v in rec["visuals"].items():
            assert _rel(_sample(visuals[k], 1025), v) < (1e-4 if it == 0 else 1e-2), (it, k)


def test_oracle_supervised_matches_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "golden_sup_n2.pt"))
    state = nets.init_model_state(seed=g["seed_w"], perturb=g["perturb"])
    a, b, z = step.synthetic_batch(g["n"], seed=g["seed_x"])
    om = step.OracleModel(state=state)
    for it, rec in enumerate(g["steps"]):
        losses = om.supervised_train_instance(a, b, z)
        tol = 2e-5 if it == 0 else 5e-3
        for k, v in rec.items():
            assert abs(losses[k] - v) <= tol * max(1.0, abs(v)), (it, k, losses[k], v)


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
def test_parameter_order_matches_reference():
    """torch.optim.Adam state_dicts index parameters by position (model.py:379-389 build the optimizers from
    itertools.chain(net.parameters() ...)), so reference-format checkpoints only interoperate if the drop-in modules
    register their parameters in the reference's order with the reference's shapes."""
    warnings.simplefilter("ignore")
    import dtg  # noqa: F401
    from dtg_b200 import networks as N
    opt = step.default_opt()
    ref = lr.build_reference_model(copy.deepcopy(opt), nets.init_model_state(seed=1))
    ours = {
        "netG_A_B": N.define_stochastic_G(nlatent=16, input_nc=3, output_nc=3, ngf=32, which_model_netG="resnet",
                                          norm="instance", use_dropout=False, gpu_ids=[]),
        "netG_B_A": N.define_G(input_nc=3, output_nc=3, ngf=32, which_model_netG="resnet", norm="instance",
                               use_dropout=False, gpu_ids=[]),
        "netE_B": N.define_E(nlatent=16, input_nc=6, nef=32, norm="batch", gpu_ids=[]),
        "netD_A": N.define_D_A(input_nc=3, ndf=32, which_model_netD="basic", norm="instance", use_sigmoid=False, gpu_ids=[]),
        "netD_B": N.define_D_B(input_nc=3, ndf=64, which_model_netD="basic", norm="instance", use_sigmoid=False, gpu_ids=[]),
        "netD_z_B": N.define_LAT_D(nlatent=16, ndf=64, use_sigmoid=False, gpu_ids=[]),
    }
    for k, net in ours.items():
        a = [(n, tuple(p.shape)) for n, p in getattr(ref, k).named_parameters()]
        b = [(n, tuple(p.shape)) for n, p in net.named_parameters()]
        assert a == b, k
        assert list(getattr(ref, k).state_dict().keys()) == list(net.state_dict().keys()), k


def test_skimage014_resize_known_answers():
    """oracle/fields.py:skimage014_resize against values computed by hand from the algorithm (bilinear, input coordinate
    = scale * (o + 0.5) - 0.5, out-of-image samples = cval 0, clip to the image range) -- the resize has no golden
    vectors (scikit-image is absent from the reference tree and from this image): PARITY UNPINNED, see the module header"""
    import numpy as np
    from oracle import fields as ofields
    x = np.array([[-1.5, -0.5], [0.5, 1.5]])[:, :, None]
    up = ofields.skimage014_resize(x, (4, 4))[:, :, 0]
    # corner: coordinate (-0.25, -0.25): only pixel (0,0) is inside -> 0.75 * 0.75 * -1.5
    assert abs(up[0, 0] - (-0.84375)) < 1e-12
    # (0, 1): row -0.25 (weight 0.75 on row 0), column 0.25 -> 0.75 * (0.75 * -1.5 + 0.25 * -0.5)
    assert abs(up[0, 1] - (-0.9375)) < 1e-12
    # interior (1, 1): coordinate (0.25, 0.25): 0.75*0.75*-1.5 + 0.75*0.25*-0.5 + 0.25*0.75*0.5 + 0.25*0.25*1.5
    assert abs(up[1, 1] - (-0.75)) < 1e-12
    # point symmetry of the data survives
    assert np.allclose(up, -up[::-1, ::-1])
    # 2x shrink of a 4x4 ramp: coordinate 2o + 0.5 -> plain 2x2 box means, no border effect
    r = np.arange(16, dtype=np.float64).reshape(4, 4, 1)
    dn = ofields.skimage014_resize(r, (2, 2))[:, :, 0]
    assert np.allclose(dn, [[2.5, 4.5], [10.5, 12.5]])
    # same size: identity
    assert np.array_equal(ofields.skimage014_resize(r, (4, 4)), r)
    # clip: an all-positive image keeps exact cval outputs but clamps the darkened border up to the image minimum
    p = ofields.skimage014_resize(np.full((2, 2, 1), 3.0), (4, 4))[:, :, 0]
    assert np.all(p == 3.0)


def test_oracle_preprocess_fields_scaling():
    """dataloader.py:17-25 arithmetic of the oracle against first principles"""
    import numpy as np
    from oracle import fields as ofields
    g = np.random.RandomState(0)
    arr = g.rand(3, 5, 6, 4)
    arr[0, 0, 0, 0] = np.nan
    arr[1, :, :, 2] = 2.0
    out = ofields.preprocess_fields(arr)
    assert out.shape == (3, 3, 5, 6) and out.dtype == np.float32
    f = np.nan_to_num(arr[0, :, :, 0])
    assert np.allclose(out[0, 0], -1 + 2 * (f - f.min()) / (f.max() - f.min()), atol=1e-6)
    assert np.all(out[1, 2] == 0) and out.min() >= -1 and out.max() <= 1
