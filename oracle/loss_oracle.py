"""Functional fp32 CPU restatement of the loss / autograd glue in reference vae_trainer.py:27-217,636-699.
TEST INFRASTRUCTURE (see oracle/__init__.py). Pinned by tests/golden/losses_*.npz and step_*.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class GradNormFunction(torch.autograd.Function):
    """vae_trainer.py:27-48: identity forward (clone); backward g -> weight * g / (mean_ranks(||g||_2) + 1e-8).
    `avg_fn` maps the local norm (python float) to the rank-average (vae_trainer.py:56-60); identity on 1 rank."""

    @staticmethod
    def forward(ctx, x, weight, avg_fn):
        ctx.weight = float(weight)
        ctx.avg_fn = avg_fn
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        n = torch.norm(g).mean().item()  # whole local tensor, in g's dtype (vae_trainer.py:40)
        if ctx.avg_fn is not None:
            n = ctx.avg_fn(n)
        return ctx.weight * g / (n + 1e-8), None, None


def gradnorm(x, weight=1.0, avg_fn=None):  # vae_trainer.py:51-53
    return GradNormFunction.apply(x, weight, avg_fn)


def gan_disc_loss(real_preds, fake_preds, disc_type="bce"):
    """vae_trainer.py:63-90 -> (loss, avg_real, avg_fake, acc)."""
    if disc_type == "bce":
        real_loss = F.binary_cross_entropy_with_logits(real_preds, torch.ones_like(real_preds))
        fake_loss = F.binary_cross_entropy_with_logits(fake_preds, torch.zeros_like(fake_preds))
    elif disc_type == "hinge":
        real_loss = F.relu(1 - real_preds).mean()
        fake_loss = F.relu(1 + fake_preds).mean()
    else:
        raise ValueError(disc_type)
    with torch.no_grad():
        acc = ((real_preds > 0).sum().item() + (fake_preds < 0).sum().item()) / (
            real_preds.numel() + fake_preds.numel())
    return (real_loss + fake_loss) * 0.5, real_preds.mean().item(), fake_preds.mean().item(), acc


def gan_gen_loss(fake_preds, disc_type):  # vae_trainer.py:688-693
    if disc_type == "bce":
        return F.binary_cross_entropy_with_logits(fake_preds, torch.ones_like(fake_preds))
    return -fake_preds.mean()


def lecam_loss(real_preds, fake_preds, anchor_real, anchor_fake):  # vae_trainer.py:651-653
    return (real_preds - anchor_fake).pow(2).mean() + (fake_preds - anchor_real).pow(2).mean()


def gaussian_kernel1d(ksize=13, sigma=2.0):  # torchvision GaussianBlur: linspace(-6,6,13), exp(-x^2/2s^2), normalised
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def blurriness_heatmap(img):
    """vae_trainer.py:143-176."""
    gray = img.mean(dim=1, keepdim=True)
    lap = torch.tensor([[0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [1, 1, -20, 1, 1], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0]],
                       dtype=torch.float32).view(1, 1, 5, 5).to(img.device)
    edge = F.conv2d(gray, lap, padding=2).abs()
    k1 = gaussian_kernel1d().to(img.device)
    k2 = (k1[:, None] * k1[None, :]).view(1, 1, 13, 13)
    edge = F.conv2d(F.pad(edge, (6, 6, 6, 6), mode="reflect"), k2)
    edge = (edge - edge.min()) / (edge.max() - edge.min() + 1e-8)
    blur = 1 - edge
    blur = torch.where(blur < 0.8, torch.zeros_like(blur), blur)
    return blur.repeat(1, 3, 1, 1)


def vae_loss_function(x, x_rec, z, do_pool=True, do_recon=False, recon_weight=0.0):
    """vae_trainer.py:179-217. `recon_weight` is the hard-coded 0.0 at :209 made explicit (0.0 reproduces HEAD).
    The reference's enabled+pooled branch leaves recon_loss_item unbound (UnboundLocalError); the restatement
    defines it as recon_loss.item() in both branches (documented fix)."""
    if do_recon:
        if do_pool:
            xr = F.interpolate(x_rec, scale_factor=1 / 16, mode="area")
            xd = F.interpolate(x, scale_factor=1 / 16, mode="area")
            recon = (xr - xd).abs().mean()
        else:
            recon = ((x_rec - x) * blurriness_heatmap(x)).abs().mean()
        recon_item = recon.item()
    else:
        recon, recon_item = 0, 0
    zloss = z.pow(2).mean()
    loss = recon * recon_weight + zloss * 0.1
    stats = {"recon_loss": recon_item, "kl_loss": zloss.item(), "average_of_abs_z": z.abs().mean().item(),
             "std_of_abs_z": z.abs().std().item(), "average_of_logvar": 0.0, "std_of_logvar": 0.0}
    return loss, stats


def flip_latent_h(z):  # vae_trainer.py:567-570
    z = torch.flip(z, [-1]).clone()
    z[:, -4:-2] = -z[:, -4:-2]
    return z


def flip_latent_v(z):  # vae_trainer.py:572-575
    z = torch.flip(z, [-2]).clone()
    z[:, -2:] = -z[:, -2:]
    return z


def cosine_lr_factor(step, warmup, total):  # transformers.get_cosine_schedule_with_warmup (vae_trainer.py:486-490)
    if step < warmup:
        return step / max(1, warmup)
    progress = (step - warmup) / max(1, total - warmup)
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 0.5 * 2.0 * progress)))


def latent_augment(z, z_s, real_hr, flip_invariance, crop_invariance, downscale_factor=16, hr=False, rnd=None):
    """vae_trainer.py:566-621 restated line by line (python `random` draw order included); `rnd` = the random module or
    a random.Random instance."""
    import random as _random

    rnd = rnd or _random
    if rnd.random() < 0.5 and flip_invariance:
        z_s = torch.flip(z_s, [-1])
        z_s[:, -4:-2] = -z_s[:, -4:-2]
        real_hr = torch.flip(real_hr, [-1])
    if rnd.random() < 0.5 and flip_invariance:
        z_s = torch.flip(z_s, [-2])
        z_s[:, -2:] = -z_s[:, -2:]
        real_hr = torch.flip(real_hr, [-2])
    if rnd.random() < 0.5 and crop_invariance:
        z_h, z_w = z.shape[-2:]
        new_z_h = rnd.randint(12, z_h - 1)
        new_z_w = rnd.randint(12, z_w - 1)
        offset_z_h = rnd.randint(0, z_h - new_z_h - 1)
        offset_z_w = rnd.randint(0, z_w - new_z_w - 1)
        m = downscale_factor * 2 if hr else downscale_factor
        new_h, new_w, offset_h, offset_w = new_z_h * m, new_z_w * m, offset_z_h * m, offset_z_w * m
        real_hr = real_hr[:, :, offset_h:offset_h + new_h, offset_w:offset_w + new_w]
        z_s = z_s[:, :, offset_z_h:offset_z_h + new_z_h, offset_z_w:offset_z_w + new_z_w]
    return z_s, real_hr
