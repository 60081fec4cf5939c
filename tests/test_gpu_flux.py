"""GPU (-m gpu): parity at the BASELINE.json configurations (ch=128, ch_mult 1,2,4,4, z=16, 256x256 — the shapes bench.py
times), against goldens produced by the UNMODIFIED reference on CPU fp32 (oracle/make_golden.py flux_step / flux_hr).

These are the shapes where the halo-tile conv path (C % 64 == 0, Cout >= 128), the N = 256 tiles, the 256-pixel double
accumulator, one-wave split-K at realistic K and GroupNorm with 4/8/16 channels per group are live.

Every tolerance is tied to a PEER: the reference's own arithmetic (oracle restatement = plain PyTorch/cuDNN) executed on
this GPU in reduced precision, measured against the same fp32 CPU golden. Two peers are run and printed:
  * "mix"  — the reference's own precision mix: TF32 encoder / LPIPS / D, bf16-autocast decoder (vae_trainer.py:18-19,
             453,623). The bound for everything downstream of the decoder (recon, losses, decoder gradients).
  * "bf16" — the same arithmetic with bf16 autocast around the encoder too: BASELINE.json's configs name bf16 as the
             compute dtype of this path and this implementation stores every activation in bf16 (DESIGN.md deviation 2),
             so the encoder output z and the encoder gradients are bounded by the all-bf16 peer: a TF32 encoder keeps
             fp32 activation storage, which bf16 storage cannot match by construction (measured z rel-L2: TF32 ~1e-3,
             bf16 eager and this implementation ~1e-2).
For every quantity q:  err_ours(q) <= max(1.5 * err_peer(q), floor)  with `floor` stated next to each assert; where the
peer reaches cosine >= 0.999 we must too.
"""
import numpy as np
import pytest
import torch

from helpers import cosine, golden, rel_l2, seeded_sd, t
from oracle import lpips_oracle as LP
from oracle import seeded
from oracle import step_oracle as SO
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu

CFG = VO.VAEConfig(resolution=256, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16)
CFG_HR = VO.VAEConfig(resolution=256, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16,
                      decoder_also_perform_hr=True)


def _build_vae(cfg, tag):
    import ae

    m = ae.VAE(resolution=cfg.resolution, in_channels=3, ch=cfg.ch, out_ch=3, ch_mult=list(cfg.ch_mult),
               num_res_blocks=cfg.num_res_blocks, z_channels=cfg.z_channels, use_attn=False,
               decoder_also_perform_hr=cfg.decoder_also_perform_hr, use_wavelet=False)
    m.load_state_dict(seeded_sd(VO.state_dict_shapes(cfg), tag), strict=True)
    return m.cuda()


def _sub_cos(grad, g, key, tag=""):
    s = int(g[f"{tag}stride::{key}"])
    return cosine(grad.detach().flatten()[::s], g[f"{tag}grad::{key}"])


def _peer_tf32(on=True):
    torch.backends.cuda.matmul.allow_tf32 = on
    torch.backends.cudnn.allow_tf32 = on


def _norm_ratio(named_grads, g, tag, only=None, scale=1.0):
    keys = [str(k) for k in g["grad_keys"]]
    ref = g[tag + "grad_norms"]
    norms = np.array([named_grads[k].float().norm().item() for k in keys]) / scale
    big = ref > 1e-3 * ref.max()
    if only is not None:
        big = big & np.array([k.startswith(only) for k in keys])
    return np.abs(norms[big] / ref[big] - 1.0), [k for k, b in zip(keys, big) if b]


def _bound(ours, peer, floor):
    return ours <= max(1.5 * peer, floor)


def _cos_bound(ours, peer, floor_gap):
    """1 - cos is the error: ours <= max(1.5 x peer's, floor). Where the peer is clearly above 0.999 (>= 0.9995; a peer
    sitting AT 0.999 +- 1e-4 would turn run-to-run bf16 noise of either side into a coin flip) we must reach 0.999 too."""
    ok = (1 - ours) <= max(1.5 * (1 - peer), floor_gap)
    if peer >= 0.9995:
        ok = ok and ours >= 0.999
    return ok


@pytest.fixture(scope="module")
def flux_models():
    import utils

    lsd = seeded_sd(LP.lpips_state_dict_shapes(), "lpips")
    dsd = seeded_sd(LP.patchd_state_dict_shapes(), "patchd")
    lp = utils.LPIPS().eval()
    lp.load_state_dict(lsd)
    disc = utils.PatchDiscriminator()
    disc.load_state_dict(dsd)
    return lp.cuda(), disc.cuda(), lsd, dsd


def _our_step(vae, lp, disc, real, gan):
    import vae_trainer as vt

    vae.zero_grad(set_to_none=True)
    z = vae.encoder(real).clamp(-8.0, 8.0)
    recon = vae.decoder(vae.reg(z))
    percep = lp(vt.gradnorm(recon), real).mean()
    vl, _ = vt.vae_loss_function(real, vt.gradnorm(recon, weight=0.001), z)
    loss = percep + vl
    if gan:
        disc.requires_grad_(False)
        loss = loss - disc(vt.gradnorm(recon, weight=1.0)).mean()
        disc.requires_grad_(True)
    loss.backward()
    return loss.detach(), percep.detach(), z.detach(), recon.detach(), \
        {k: p.grad.detach() for k, p in vae.named_parameters()}


def _peer_step(vsd, lsd, dsd, real, gan, all_bf16=False):
    """The reference arithmetic in plain PyTorch on this GPU: reference precision mix (TF32 + bf16-autocast decoder), or
    with all_bf16 the encoder under bf16 autocast as well."""
    import contextlib

    _peer_tf32(True)
    try:
        osd = {k: v.cuda().requires_grad_(True) for k, v in vsd.items()}
        lsd_c, dsd_c = {k: v.cuda() for k, v in lsd.items()}, {k: v.cuda() for k, v in dsd.items()}
        if all_bf16:  # generator_step with the encoder inside autocast too (z back to fp32 like the module boundary)
            from oracle import loss_oracle as LO

            with torch.autocast("cuda", dtype=torch.bfloat16):
                z = VO.encoder_forward(osd, real, CFG)
            z = z.float().clamp(-8.0, 8.0)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                recon = VO.decoder_forward(osd, VO.reg(z), CFG)
            with torch.autocast("cuda", dtype=torch.bfloat16):  # LPIPS and D trunks in bf16 as well
                percep = LP.lpips_forward(lsd_c, LO.gradnorm(recon, 1.0), real).float().mean()
            vl, _ = LO.vae_loss_function(real, LO.gradnorm(recon, 0.001), z, do_pool=True, do_recon=False)
            loss = percep + vl
            if gan:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    fake = LP.patchd_forward(dsd_c, LO.gradnorm(recon, 1.0))
                loss = loss + LO.gan_gen_loss(fake.float(), "hinge")
            loss.backward()
            o = {"loss": loss.detach(), "percep": percep.detach(), "z": z.detach(), "recon": recon.detach()}
        else:
            o = SO.generator_step(osd, lsd_c, dsd_c, real, CFG, do_clamp=True, do_ganloss=gan, disc_type="hinge",
                                  amp_decoder=True)
        return o["loss"], o["percep"], o["z"], o["recon"].float(), {k: v.grad.detach() for k, v in osd.items()}
    finally:
        _peer_tf32(False)


@pytest.mark.parametrize("gan", [False, True])
@pytest.mark.parametrize("batch", [1, 32])
def test_flux_generator_step_vs_reference_golden(flux_models, gan, batch):
    """configs[1] (gan=False) / configs[2] generator pass (gan=True) at B=1 and at the bench batch B=32 (the B=1 golden
    tiled: GroupNorm/LPIPS are per-sample and every loss is a batch mean, so z / recon / losses repeat per sample).
    GradNorm divides by ||dL/d recon||_2 over the WHOLE batch tensor (vae_trainer.py:27-48): with B identical samples
    that norm is 1/sqrt(B) of the B=1 one, so every gradient that flows through the decoder is exactly sqrt(B) x the B=1
    golden, while the un-normalised 0.1*mean(z^2) path into the encoder is not scaled. At B=32 the decoder gradients are
    therefore compared with sqrt(32) x golden, and the encoder gradients (a B-dependent mix) at B=1 only."""
    lp, disc, lsd, dsd = flux_models
    g = golden("step_flux")
    tag = "gan_" if gan else "nogan_"
    real1 = seeded.tensor("step_flux/real", (1, 3, 256, 256), 1.0, "uniform").cuda()
    real = real1.repeat(batch, 1, 1, 1).contiguous()
    vsd = seeded_sd(VO.state_dict_shapes(CFG), "step_flux/vae")
    vae = _build_vae(CFG, "step_flux/vae")
    loss, percep, z, recon, grads = _our_step(vae, lp, disc, real, gan)
    torch.cuda.synchronize()

    # every sample of the tiled batch is compared with the golden (atomics in the fused GroupNorm statistics make
    # identical samples differ by bf16 rounding noise, so "sample i == sample 0" only holds to the parity tolerance)
    gz = np.repeat(g["z"], batch, 0)
    gr = np.repeat(g["recon"].astype(np.float32), batch, 0)
    ez, er = rel_l2(z, gz), rel_l2(recon, gr)
    if batch > 1:
        worst = max(rel_l2(z[i:i + 1], g["z"]) for i in range(batch))
        assert worst < 1.5 * ez + 1e-3, "one sample of the tiled batch is off: cross-sample leakage in the tiles?"
    el = abs(loss.item() - float(g[tag + "loss"])) / abs(float(g[tag + "loss"]))
    ep = abs(percep.item() - float(g[tag + "percep"])) / abs(float(g[tag + "percep"]))
    picks = [k[len(tag) + 6:] for k in g if k.startswith(tag + "grad::")]
    if batch > 1:
        picks = [k for k in picks if k.startswith("decoder.")]
    cos = {k: _sub_cos(grads[k], g, k, tag) for k in picks}
    sc = batch ** 0.5
    nr_dec, nk_dec = _norm_ratio(grads, g, tag, "decoder.", sc)
    nr_enc, nk_enc = _norm_ratio(grads, g, tag, "encoder.", 1.0) if batch == 1 else (np.zeros(1), [""])

    peers = {}
    for name, allbf in (("mix", False), ("bf16", True)):
        pl, pp, pz, pr, pg = _peer_step(vsd, lsd, dsd, real1, gan, all_bf16=allbf)
        peers[name] = dict(
            ez=rel_l2(pz, g["z"]), er=rel_l2(pr, g["recon"].astype(np.float32)),
            el=abs(pl.item() - float(g[tag + "loss"])) / abs(float(g[tag + "loss"])),
            ep=abs(pp.item() - float(g[tag + "percep"])) / abs(float(g[tag + "percep"])),
            nr_dec=_norm_ratio(pg, g, tag, "decoder.")[0], nr_enc=_norm_ratio(pg, g, tag, "encoder.")[0],
            cos={k: _sub_cos(pg[k], g, k, tag) for k in picks})
    M, Bf = peers["mix"], peers["bf16"]

    print(f"\nflux step gan={gan} B={batch}   ours | peer 'mix' (TF32 enc + bf16-autocast dec) | peer 'bf16' (all autocast)"
          f"  — all vs the fp32 reference golden")
    print(f"  z rel_l2      {ez:.3e} | {M['ez']:.3e} | {Bf['ez']:.3e}")
    print(f"  recon rel_l2  {er:.3e} | {M['er']:.3e} | {Bf['er']:.3e}")
    print(f"  loss rel      {el:.3e} | {M['el']:.3e} | {Bf['el']:.3e}     percep rel {ep:.3e} | {M['ep']:.3e} | {Bf['ep']:.3e}")
    print(f"  decoder grad-norm |ratio-1| max {nr_dec.max():.4f} | {M['nr_dec'].max():.4f} | {Bf['nr_dec'].max():.4f}"
          f"   mean {nr_dec.mean():.4f} | {M['nr_dec'].mean():.4f} | {Bf['nr_dec'].mean():.4f}  (worst ours {nk_dec[int(nr_dec.argmax())]})")
    if batch == 1:
        print(f"  encoder grad-norm |ratio-1| max {nr_enc.max():.4f} | {M['nr_enc'].max():.4f} | {Bf['nr_enc'].max():.4f}"
              f"   mean {nr_enc.mean():.4f} | {M['nr_enc'].mean():.4f} | {Bf['nr_enc'].mean():.4f}  (worst ours {nk_enc[int(nr_enc.argmax())]})")
    for k in picks:
        print(f"  cos {k:48s} {cos[k]:.5f} | {M['cos'][k]:.5f} | {Bf['cos'][k]:.5f}")

    # encoder-side quantities: bounded by the all-bf16 peer (bf16 activation storage, DESIGN deviation 2);
    # everything downstream of the decoder: bounded by the reference's own precision mix
    assert _bound(ez, Bf["ez"], 5e-3), "z"
    assert _bound(er, M["er"], 1e-2), "recon"
    # the GAN term is a mean over 256 patch logits whose bf16 errors are spatially coherent (weight rounding acts on
    # positive post-ReLU features): the mean inherits the per-logit error level, 1.5e-2 for this implementation AND for
    # eager bf16 D (test_flux_discriminator_step...; measured here 1.35e-2 .. 1.5e-2 over runs), hence the 3e-2 floor with
    # the GAN term, 5e-3 without
    assert _bound(ep, M["ep"], 5e-3) and _bound(el, max(M["el"], Bf["el"]), 3e-2 if gan else 5e-3), "losses"
    # with the GAN term every gradient first crosses the 13 bf16 layers of the discriminator, which the "mix" peer runs
    # in TF32: the decoder quantities are then bounded by the all-bf16 peer as well
    D = Bf if gan else M
    # (floors: this implementation's own run-to-run spread — atomics reorder the fused GroupNorm statistics and bf16
    #  rounding amplifies that — measured mean |ratio-1| 0.005 .. 0.012 with the GAN term over repeated runs)
    assert _bound(nr_dec.max(), D["nr_dec"].max(), 0.04 if gan else 0.02) and \
        _bound(nr_dec.mean(), D["nr_dec"].mean(), 0.02 if gan else 0.01), "dec norms"
    if batch == 1:
        # run-to-run spread of this implementation (fp32 atomics in the fused GroupNorm statistics reorder sums, and bf16
        # rounding amplifies that through ~60 layers): measured mean |ratio-1| 0.002-0.011 (gan) over repeated runs, so
        # the floors are 0.05 (max) / 0.02 (mean) with the GAN term, 0.02 / 0.01 without
        assert _bound(nr_enc.max(), Bf["nr_enc"].max(), 0.05 if gan else 0.02) and \
            _bound(nr_enc.mean(), Bf["nr_enc"].mean(), 0.02 if gan else 0.01), "enc norms"
    bad = [k for k in picks
           if not _cos_bound(cos[k], (Bf if (gan or k.startswith("encoder.")) else M)["cos"][k], 2e-3)]
    assert not bad, [(k, cos[k], M["cos"][k], Bf["cos"][k]) for k in bad]


def test_flux_discriminator_step_vs_reference_golden(flux_models):
    """configs[2] discriminator pass at 256^2: hinge + LeCam (anchors 0.1 / 0.05), logits and all 42 gradient norms."""
    import vae_trainer as vt

    lp, disc, lsd, dsd = flux_models
    g = golden("step_flux")
    real = seeded.tensor("step_flux/real", (1, 3, 256, 256), 1.0, "uniform").cuda()
    recon = t(g["recon"].astype(np.float32)).cuda()
    disc.zero_grad(set_to_none=True)
    rp, fp = disc(real), disc(recon)
    dl, ar, af, acc = vt.gan_disc_loss(rp, fp, "hinge")
    total = dl.mean() + 0.1 * ((rp - 0.05).pow(2).mean() + (fp - 0.1).pow(2).mean())
    total.backward()
    dgr = {k: p.grad.detach() for k, p in disc.named_parameters()}

    _peer_tf32(True)
    try:
        osd = {k: (v.cuda().requires_grad_(True) if torch.is_floating_point(v) and "scaling" not in k else v.cuda())
               for k, v in dsd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):  # the closest reference-style reduced-precision D pass
            prp, pfp = LP.patchd_forward(osd, real), LP.patchd_forward(osd, recon)
        pdl = (torch.relu(1 - prp.float()).mean() + torch.relu(1 + pfp.float()).mean()) * 0.5
        ptotal = pdl + 0.1 * ((prp.float() - 0.05).pow(2).mean() + (pfp.float() - 0.1).pow(2).mean())
        ptotal.backward()
        pgr = {k: v.grad.detach() for k, v in osd.items() if v.grad is not None}
    finally:
        _peer_tf32(False)

    dkeys = [str(k) for k in g["d_grad_keys"]]
    ref = g["d_grad_norms"]
    big = ref > 1e-3 * ref.max()
    ours_r = np.abs(np.array([dgr[k].norm().item() for k in dkeys])[big] / ref[big] - 1)
    peer_r = np.abs(np.array([pgr[k].float().norm().item() for k in dkeys])[big] / ref[big] - 1)
    e_real, e_fake = rel_l2(rp, g["d_logits_real"]), rel_l2(fp, g["d_logits_fake"])
    pe_real, pe_fake = rel_l2(prp.float(), g["d_logits_real"]), rel_l2(pfp.float(), g["d_logits_fake"])
    ed = abs(total.item() - float(g["d_loss"])) / abs(float(g["d_loss"]))
    ped = abs(ptotal.item() - float(g["d_loss"])) / abs(float(g["d_loss"]))
    picks = [k[8:] for k in g if k.startswith("d_grad::")]
    print(f"\nflux D step (ours | eager bf16-autocast peer): loss rel {ed:.3e} | {ped:.3e}; logits real {e_real:.3e} | "
          f"{pe_real:.3e} fake {e_fake:.3e} | {pe_fake:.3e}; grad-norm |ratio-1| max {ours_r.max():.4f} | {peer_r.max():.4f}")
    # the loss is a mean of hinge terms of logits that themselves carry ~1.5e-2 (ours and peer alike): floor 1.5e-2
    assert _bound(ed, ped, 1.5e-2) and _bound(e_real, pe_real, 1e-2) and _bound(e_fake, pe_fake, 1e-2)
    assert _bound(ours_r.max(), peer_r.max(), 0.02)
    for k in picks:
        c, pc = _sub_cos(dgr[k], g, k, "d_"), _sub_cos(pgr[k].float(), g, k, "d_")
        print(f"  cos {k:40s} {c:.5f} | {pc:.5f}")
        assert _cos_bound(c, pc, 2e-3), k


def test_flux_hr_decoder_vs_reference_golden():
    """configs[4] topology: ch=128 encoder at 256^2 + the decoder's extra x2 level (ae.py:381) -> 512^2."""
    name = "vae_flux_hr"
    g = golden(name)
    vae = _build_vae(CFG_HR, name)
    x = seeded.tensor(name + "/x", (1, 3, 256, 256), 1.0, "uniform").cuda()
    dec, z = vae(x)
    assert tuple(dec.shape) == (1, 3, 512, 512)
    (dec.pow(2).mean() + z.pow(2).mean()).backward()
    grads = {k: p.grad.detach() for k, p in vae.named_parameters()}

    _peer_tf32(True)
    try:
        osd = {k: v.cuda().requires_grad_(True) for k, v in seeded_sd(VO.state_dict_shapes(CFG_HR), name).items()}
        pz = VO.encoder_forward(osd, x, CFG_HR)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pdec = VO.decoder_forward(osd, VO.reg(pz), CFG_HR)
            pz_bf16 = VO.encoder_forward(osd, x, CFG_HR).float()  # all-bf16 peer for the encoder output (see module doc)
        (pdec.float().pow(2).mean() + pz.pow(2).mean()).backward()
        pg = {k: v.grad.detach() for k, v in osd.items()}
    finally:
        _peer_tf32(False)
    pez_bf16 = rel_l2(pz_bf16, g["z"])
    ez, ed = rel_l2(z, g["z"]), rel_l2(dec, g["dec"].astype(np.float32))
    pez, ped = rel_l2(pz, g["z"]), rel_l2(pdec.float(), g["dec"].astype(np.float32))
    nr, nk = _norm_ratio(grads, g, "")
    pnr, _ = _norm_ratio(pg, g, "")
    print(f"\n{name} (ours | peer mix): z {ez:.3e} | {pez:.3e} (all-bf16 peer {pez_bf16:.3e})  dec {ed:.3e} | {ped:.3e}  "
          f"grad-norm |ratio-1| max {nr.max():.4f} | {pnr.max():.4f} (worst ours {nk[int(nr.argmax())]})")
    assert _bound(ez, pez_bf16, 5e-3) and _bound(ed, ped, 1e-2)
    assert _bound(nr.max(), pnr.max(), 0.02)
    for k in [k[6:] for k in g if k.startswith("grad::")]:
        c, pc = _sub_cos(grads[k], g, k), _sub_cos(pg[k], g, k)
        print(f"  cos {k:48s} {c:.5f} | {pc:.5f}")
        assert _cos_bound(c, pc, 2e-3), k
