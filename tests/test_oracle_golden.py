"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(oracle/make_golden.py). This is the pin required before any CUDA parity claim."""
import numpy as np
import pytest
import torch

from helpers import golden, rel_l2, seeded_sd, t
from oracle import loss_oracle as LO
from oracle import lpips_oracle as LP
from oracle import seeded
from oracle import step_oracle as SO
from oracle import vae_oracle as VO
from oracle import vq_oracle as VQ

VAE_CASES = {
    "vae_small": (VO.VAEConfig(resolution=32, ch=32, ch_mult=(1, 2), num_res_blocks=2, z_channels=4), 2, 32),
    "vae_attn": (VO.VAEConfig(resolution=32, ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=4, use_attn=True), 2, 32),
    "vae_hr": (VO.VAEConfig(resolution=32, ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=4,
                            decoder_also_perform_hr=True), 1, 32),
}


@pytest.mark.parametrize("name", sorted(VAE_CASES))
def test_vae_oracle_matches_reference_golden(name):
    cfg, N, R = VAE_CASES[name]
    g = golden(name)
    sd = {k: v.requires_grad_(True) for k, v in seeded_sd(VO.state_dict_shapes(cfg), name).items()}
    x = seeded.tensor(name + "/x", (N, 3, R, R), 1.0, "uniform")
    dec, z = VO.vae_forward(sd, x, cfg)
    assert rel_l2(z, g["z"]) < 1e-5
    assert rel_l2(dec, g["dec"]) < 1e-5
    (dec.pow(2).mean() + z.pow(2).mean()).backward()
    keys = [str(k) for k in g["grad_keys"]]
    assert sorted(sd) == keys
    norms = np.array([sd[k].grad.norm().item() for k in keys])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-6)
    for k in g:
        if k.startswith("grad::"):
            assert rel_l2(sd[k[6:]].grad, g[k]) < 1e-4, k


def test_lpips_oracle_matches_reference_golden():
    g = golden("lpips_small")
    sd = seeded_sd(LP.lpips_state_dict_shapes(), "lpips")
    a = seeded.tensor("lpips_small/a", (2, 3, 32, 32), 1.0, "uniform").requires_grad_(True)
    b = seeded.tensor("lpips_small/b", (2, 3, 32, 32), 1.0, "uniform")
    val = LP.lpips_forward(sd, a, b)
    assert val.shape == (2, 1, 1, 1)
    assert rel_l2(val, g["val"]) < 1e-5
    val.mean().backward()
    assert rel_l2(a.grad, g["grad_a"]) < 1e-4


def test_patchd_oracle_matches_reference_golden():
    g = golden("patchd_small")
    sd = seeded_sd(LP.patchd_state_dict_shapes(), "patchd")
    for k, v in sd.items():
        if "scaling" not in k:
            v.requires_grad_(True)
    x = seeded.tensor("patchd_small/x", (2, 3, 32, 32), 1.0, "uniform").requires_grad_(True)
    y = LP.patchd_forward(sd, x)
    assert rel_l2(y, g["logits"]) < 1e-5
    (y * seeded.tensor("patchd_small/gy", y.shape)).sum().backward()
    assert rel_l2(x.grad, g["grad_x"]) < 1e-4
    keys = [str(k) for k in g["grad_keys"]]
    norms = np.array([sd[k].grad.norm().item() for k in keys])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-6)


def test_loss_oracle_matches_reference_golden():
    g = golden("losses")
    r, f = seeded.tensor("losses/real", (4, 16)), seeded.tensor("losses/fake", (4, 16))
    for dt in ("hinge", "bce"):
        l, ar, af, acc = LO.gan_disc_loss(r, f, dt)
        np.testing.assert_allclose([l.item(), ar, af, acc], g[dt], rtol=1e-5, atol=1e-7)
    x = seeded.tensor("losses/x", (2, 3, 32, 32), 1.0, "uniform")
    xr = seeded.tensor("losses/xr", (2, 3, 32, 32), 1.0, "uniform")
    z = seeded.tensor("losses/z", (2, 4, 8, 8))
    vl, st = LO.vae_loss_function(x, xr, z)
    np.testing.assert_allclose(vl.item(), g["vae_loss"], rtol=1e-6)
    np.testing.assert_allclose(st["kl_loss"], g["kl_loss"], rtol=1e-6)
    np.testing.assert_allclose(st["average_of_abs_z"], g["abs_z"], rtol=1e-6)
    np.testing.assert_allclose(st["std_of_abs_z"], g["std_abs_z"], rtol=1e-5)
    assert rel_l2(LO.blurriness_heatmap(x), g["heat"]) < 1e-5
    _, st2 = LO.vae_loss_function(x, xr, z, do_pool=False, do_recon=True)
    np.testing.assert_allclose(st2["recon_loss"], g["lowpass_recon"], rtol=1e-5)
    gi = seeded.tensor("losses/gn_x", (2, 3, 8, 8)).requires_grad_(True)
    (LO.gradnorm(gi, 0.5) * seeded.tensor("losses/gn_gy", (2, 3, 8, 8))).sum().backward()
    assert rel_l2(gi.grad, g["gradnorm_grad"]) < 1e-6
    assert abs(gi.grad.norm().item() - 0.5) < 1e-5  # the rescaled gradient has norm == weight
    assert rel_l2(LP.wavelet_transform_multi_channel(x), g["wavelet"]) < 1e-6


def test_step_oracle_matches_reference_golden():
    g = golden("step_small")
    cfg = VO.VAEConfig(resolution=32, ch=32, ch_mult=(1, 2), num_res_blocks=1, z_channels=4)
    lsd = seeded_sd(LP.lpips_state_dict_shapes(), "lpips")
    dsd = seeded_sd(LP.patchd_state_dict_shapes(), "patchd")
    real = seeded.tensor("step_small/real", (2, 3, 32, 32), 1.0, "uniform")
    keys = [str(k) for k in g["grad_keys"]]
    for gan, tag in ((False, "nogan"), (True, "gan")):
        vsd = {k: v.requires_grad_(True) for k, v in seeded_sd(VO.state_dict_shapes(cfg), "step_small/vae").items()}
        o = SO.generator_step(vsd, lsd, dsd, real, cfg, do_clamp=True, do_ganloss=gan, disc_type="hinge")
        np.testing.assert_allclose(o["loss"].item(), g[tag + "_loss"], rtol=1e-5)
        np.testing.assert_allclose(o["percep"].item(), g[tag + "_percep"], rtol=1e-5)
        norms = np.array([vsd[k].grad.norm().item() for k in keys])
        np.testing.assert_allclose(norms, g[tag + "_grad_norms"], rtol=5e-4, atol=1e-7)
        assert rel_l2(vsd["encoder.conv_in.weight"].grad, g[tag + "_grad_conv_in"]) < 2e-4
    assert rel_l2(o["recon"], g["recon"]) < 1e-5
    dsd2 = {k: (v.requires_grad_(True) if "scaling" not in k else v) for k, v in
            seeded_sd(LP.patchd_state_dict_shapes(), "patchd").items()}
    od = SO.discriminator_step(dsd2, real, t(g["recon"]), "hinge", True, (0.1, 0.05))
    np.testing.assert_allclose(od["d_loss"].item(), g["d_loss"], rtol=1e-5)
    dkeys = [str(k) for k in g["d_grad_keys"]]
    dn = np.array([dsd2[k].grad.norm().item() for k in dkeys])
    np.testing.assert_allclose(dn, g["d_grad_norms"], rtol=5e-4, atol=1e-7)


def test_vq_oracle_properties():
    """Config-4 codebook oracle: parity UNPINNED by the reference (no VQ there); check the canonical semantics."""
    rng = np.random.default_rng(0)
    e = rng.uniform(-1 / 64, 1 / 64, size=(64, 16)).astype(np.float32)
    z = rng.normal(size=(200, 16)).astype(np.float32) * 0.01
    z[:10] = e[5:15]  # exact hits
    e[40] = e[7]      # a duplicate code: the FIRST index must win
    zq, idx, loss, gap = VQ.vq_forward(z, e)
    assert (idx[:10] == np.arange(5, 15)).all()
    assert idx[2] == 7 and gap[2] == 0.0
    d = ((z[:, None, :].astype(np.float64) - e[None].astype(np.float64)) ** 2).sum(-1)
    assert (np.take_along_axis(d, idx[:, None], 1)[:, 0] <= d.min(1) + 1e-9).all()
    assert np.allclose(zq, e[idx]) and loss >= 0


def test_oracle_matches_reference_golden_at_the_flux_config_forward():
    """The oracle is pinned at the BASELINE shapes too (ch=128, mult 1,2,4,4, z=16, 256x256): encoder -> clamp -> decoder of
    the step_flux fixture (generated by the unmodified reference) in fp32 on CPU. (Forward only here: the full step with
    every gradient was compared when the fixture was written — oracle/make_golden.py refuses to write otherwise — and
    takes ~15 s per pass.)"""
    cfg = VO.VAEConfig(resolution=256, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16)
    g = golden("step_flux")
    sd = seeded_sd(VO.state_dict_shapes(cfg), "step_flux/vae")
    x = seeded.tensor("step_flux/real", (1, 3, 256, 256), 1.0, "uniform")
    with torch.no_grad():
        z = VO.encoder_forward(sd, x, cfg).clamp(-8.0, 8.0)
        rec = VO.decoder_forward(sd, VO.reg(z), cfg)
    assert rel_l2(z, g["z"]) < 1e-5
    assert rel_l2(rec, g["recon"].astype(np.float32)) < 5e-4  # the fixture stores the image in fp16
    assert len(g["grad_keys"]) == 224 and sorted(sd) == [str(k) for k in g["grad_keys"]]
