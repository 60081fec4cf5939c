"""One training step of the reference loop (vae_trainer.py:530-708) restated over the oracle functions.
TEST INFRASTRUCTURE (see oracle/__init__.py). Pinned by tests/golden/step_*.npz (generated from a harness that
drives the reference's own modules the same way; the reference's train_ddp cannot run outside the author's
machine: hard-coded dataset path :386-387, wandb entity, webdataset).

Deliberate, documented deviations: LPIPS in eval mode (no dropout), `recon_weight`
exposed (0.0 = HEAD), data-parallel gradient averaging is the caller's job.
"""
from __future__ import annotations

import torch

from . import loss_oracle as LO
from . import lpips_oracle as LP
from . import vae_oracle as VO


def generator_step(vae_sd, lpips_sd, disc_sd, real, cfg: VO.VAEConfig, do_clamp=True, clamp_th=8.0,
                   do_ganloss=False, disc_type="hinge", recon_weight=0.0, avg_fn=None, amp_decoder=False):
    """Forward + backward of the generator (VAE) loss for one batch. `vae_sd` tensors must require grad.
    Returns dict(loss, percep, zloss, g_gan, recon, z) after calling backward (grads land in vae_sd tensors)."""
    z = VO.encoder_forward(vae_sd, real, cfg)                       # :538
    if do_clamp:
        z = z.clamp(-clamp_th, clamp_th)                            # :561-562
    z_s = VO.reg(z)                                                 # :563
    if amp_decoder:  # the reference's own precision mix on a GPU: decoder under bf16 autocast (:453,623), rest TF32 (:18-19)
        with torch.autocast(device_type=z_s.device.type, dtype=torch.bfloat16):
            recon = VO.decoder_forward(vae_sd, z_s, cfg)
    else:
        recon = VO.decoder_forward(vae_sd, z_s, cfg)                # :623-624
    rec_p = LO.gradnorm(recon, 1.0, avg_fn)                         # :662
    percep = LP.lpips_forward(lpips_sd, rec_p, real).mean()         # :676
    rec_m = LO.gradnorm(recon, 0.001, avg_fn)                       # :679
    vae_loss, stats = LO.vae_loss_function(real, rec_m, z, do_pool=True, do_recon=recon_weight != 0.0,
                                           recon_weight=recon_weight)  # :680
    g_gan = torch.zeros(())
    if do_ganloss:
        fake = LP.patchd_forward(disc_sd, LO.gradnorm(recon, 1.0, avg_fn))  # :683-684
        g_gan = LO.gan_gen_loss(fake, disc_type)                    # :688-693
        loss = percep + g_gan + vae_loss                            # :695
    else:
        loss = percep + vae_loss                                    # :698
    loss.backward()                                                 # :701
    return {"loss": loss.detach(), "percep": percep.detach(), "zloss": torch.tensor(stats["kl_loss"]),
            "g_gan": g_gan.detach(), "recon": recon.detach(), "z": z.detach()}


def discriminator_step(disc_sd, real, recon_detached, disc_type="hinge", use_lecam=False, anchors=(0.0, 0.0),
                       lecam_weight=0.1):
    """vae_trainer.py:629-659. disc_sd tensors must require grad. Returns dict(d_loss, avg_real, avg_fake, acc)."""
    real_p = LP.patchd_forward(disc_sd, real)
    fake_p = LP.patchd_forward(disc_sd, recon_detached)
    d_loss, avg_r, avg_f, acc = LO.gan_disc_loss(real_p, fake_p, disc_type)
    total = d_loss.mean()
    if use_lecam:
        total = total + LO.lecam_loss(real_p, fake_p, anchors[0], anchors[1]) * lecam_weight
    total.backward()
    return {"d_loss": total.detach(), "avg_real": avg_r, "avg_fake": avg_f, "acc": acc}
