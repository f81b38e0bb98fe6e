"""TEST INFRASTRUCTURE — CPU oracle of the policy hot path (never imported by the product).

Functional restatement (plain differentiable torch ops, driven by a reference-format
state_dict) of ConditionalUnet1D.forward and of the epsilon-prediction training loss
around it; backward comes from torch autograd over these ops.

PARITY PIN: no reference tests exist for this path.  Pinned against the
reference ITSELF: tests/golden/make_policy_golden.py runs the unmodified
diffuser/diffusion_policy/model/conditional_unet1d.py and commits outputs / gradient
fingerprints; tests/test_policy_oracle.py checks this file against them and, when
/root/reference is mounted, against the live module (forward and every gradient).

Third-party arithmetic: `diffusers` DDPMScheduler (unpinned in the reference's
requirements.txt, not installed here) — `ddpm_alphas_cumprod` / `add_noise` restate its
published squaredcos_cap_v2 schedule (twin in-repo formula:
flowdiffusion/.../guided_diffusion/gaussian_diffusion.py:45-62); unpinned against diffusers
itself, pinned against that vendored twin run unmodified (tests/golden/make_scheduler_golden.py,
tests/test_scheduler_pin.py: betas, alphas_cumprod, add_noise = q_sample).

dp/ = diffuser/diffusion_policy/model/ in the reference checkout.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def sinusoidal_pos_emb(t: Tensor, dim: int) -> Tensor:
    """[sin | cos], freq exp(-ln(1e4) i/(half-1)) (dp/positional_embedding.py:5-17)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=t.device) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def conv1d_block(sd: SD, p: str, x: Tensor, n_groups: int) -> Tensor:
    """Conv1d(k, pad k//2) -> GroupNorm -> Mish (dp/conv1d_components.py:23-40)."""
    w = sd[p + "block.0.weight"]
    y = F.conv1d(x, w, sd[p + "block.0.bias"], padding=w.shape[-1] // 2)
    return F.mish(F.group_norm(y, n_groups, sd[p + "block.1.weight"], sd[p + "block.1.bias"], 1e-5))


def cond_res_block(sd: SD, p: str, x: Tensor, cond: Tensor, n_groups: int) -> Tensor:
    """ConditionalResidualBlock1D.forward, cond_predict_scale=True: FiLM is scale*out+bias,
    embed.reshape(B, 2, C, 1) -> [scale | bias] (dp/conditional_unet1d.py:46-66)."""
    out = conv1d_block(sd, p + "blocks.0.", x, n_groups)
    e = F.linear(F.mish(cond), sd[p + "cond_encoder.1.weight"], sd[p + "cond_encoder.1.bias"])
    C = out.shape[1]
    if e.shape[1] == 2 * C:
        e = e.reshape(e.shape[0], 2, C, 1)
        out = e[:, 0] * out + e[:, 1]
    else:
        out = out + e[:, :, None]
    out = conv1d_block(sd, p + "blocks.1.", out, n_groups)
    if p + "residual_conv.weight" in sd:
        x = F.conv1d(x, sd[p + "residual_conv.weight"], sd[p + "residual_conv.bias"])
    return out + x


def unet1d_forward(sd: SD, sample: Tensor, timestep: Tensor, global_cond: Optional[Tensor],
                   n_groups: int = 8, p: str = "") -> Tensor:
    """ConditionalUnet1D.forward with local_cond=None (dp/conditional_unet1d.py:178-246).
    sample [B, T, D] -> [B, T, D]."""
    x = sample.permute(0, 2, 1)
    t = timestep.expand(sample.shape[0])
    dsed = sd[p + "diffusion_step_encoder.1.weight"].shape[1]
    g = sinusoidal_pos_emb(t, dsed)
    g = F.linear(g, sd[p + "diffusion_step_encoder.1.weight"], sd[p + "diffusion_step_encoder.1.bias"])
    g = F.linear(F.mish(g), sd[p + "diffusion_step_encoder.3.weight"], sd[p + "diffusion_step_encoder.3.bias"])
    if global_cond is not None:
        g = torch.cat([g, global_cond], dim=-1)
    hs: List[Tensor] = []
    i = 0
    while f"{p}down_modules.{i}.0.blocks.0.block.0.weight" in sd:
        q = f"{p}down_modules.{i}."
        x = cond_res_block(sd, q + "0.", x, g, n_groups)
        x = cond_res_block(sd, q + "1.", x, g, n_groups)
        hs.append(x)
        if q + "2.conv.weight" in sd:  # Downsample1d: Conv1d(dim, dim, 3, 2, 1)
            x = F.conv1d(x, sd[q + "2.conv.weight"], sd[q + "2.conv.bias"], stride=2, padding=1)
        i += 1
    for j in range(2):
        x = cond_res_block(sd, f"{p}mid_modules.{j}.", x, g, n_groups)
    i = 0
    while f"{p}up_modules.{i}.0.blocks.0.block.0.weight" in sd:
        q = f"{p}up_modules.{i}."
        x = torch.cat((x, hs.pop()), dim=1)
        x = cond_res_block(sd, q + "0.", x, g, n_groups)
        x = cond_res_block(sd, q + "1.", x, g, n_groups)
        if q + "2.conv.weight" in sd:  # Upsample1d: ConvTranspose1d(dim, dim, 4, 2, 1)
            x = F.conv_transpose1d(x, sd[q + "2.conv.weight"], sd[q + "2.conv.bias"], stride=2, padding=1)
        i += 1
    x = conv1d_block(sd, p + "final_conv.0.", x, 8)  # final Conv1dBlock uses the default n_groups=8
    x = F.conv1d(x, sd[p + "final_conv.1.weight"], sd[p + "final_conv.1.bias"])
    return x.permute(0, 2, 1)


def ddpm_alphas_cumprod(num_train_timesteps: int = 100, max_beta: float = 0.999) -> Tensor:
    """diffusers DDPMScheduler(beta_schedule='squaredcos_cap_v2'): betas_for_alpha_bar, fp32."""
    ab = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    T = num_train_timesteps
    betas = torch.tensor([min(1 - ab((i + 1) / T) / ab(i / T), max_beta) for i in range(T)], dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(acp: Tensor, x0: Tensor, noise: Tensor, t: Tensor) -> Tensor:
    """DDPMScheduler.add_noise: sqrt(acp_t) x0 + sqrt(1 - acp_t) noise."""
    a = acp.to(x0.device)[t] ** 0.5
    b = (1 - acp.to(x0.device)[t]) ** 0.5
    return a[:, None, None] * x0 + b[:, None, None] * noise


def epsilon_loss(sd: SD, trajectory: Tensor, global_cond: Tensor, noise: Tensor, timesteps: Tensor,
                 acp: Tensor) -> Tensor:
    """The UNet1D part of compute_loss (diffusion_unet_image_policy.py:246-276), prediction_type epsilon."""
    noisy = add_noise(acp, trajectory, noise, timesteps)
    pred = unet1d_forward(sd, noisy, timesteps, global_cond)
    loss = F.mse_loss(pred, noise, reduction="none")
    return loss.reshape(loss.shape[0], -1).mean(dim=1).mean()


def seeded_policy_state_dict(shapes: Dict[str, tuple], seed: int) -> SD:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "weight" and len(shp) == 1:
            v = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif leaf == "bias":
            v = 0.1 * torch.randn(shp, generator=g)
        else:
            # ConvTranspose1d weights are [Cin, Cout, k]; fan-in is still prod(shape[1:]) up to k/stride
            v = torch.randn(shp, generator=g) / math.sqrt(max(1, math.prod(shp[1:])))
        sd[name] = v
    return sd


def seeded_full_policy_state_dict(layout: Dict[str, list], seed: int) -> SD:
    """Weights for the WHOLE DiffusionUnetImagePolicy state_dict (414 keys: UNet1D + two VisualCore
    encoders, whose Sequential aliases 'nets.0...' / 'backbone...' must carry the SAME tensor)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    canon: Dict[str, Tensor] = {}
    for name in layout:
        shp = tuple(layout[name])
        leaf = name.rsplit(".", 1)[-1]
        # VisualCore registers backbone/pool both directly and inside .nets (Sequential): alias by canonical key
        key = re.sub(r"^(.*key_model_map\.[^.]+)\.nets\.0\.", r"\1.backbone.", name)
        key = re.sub(r"^(.*key_model_map\.[^.]+)\.nets\.1\.", r"\1.pool.", key)
        if key in canon:
            sd[name] = canon[key]
            continue
        if leaf in ("pos_x", "pos_y", "temperature"):
            v = None                                  # buffers keep their constructed values (merge with
                                                      # the module's own state_dict before loading)
        elif len(shp) == 0 or math.prod(shp) == 0:
            v = torch.zeros(shp)
        elif leaf == "weight" and len(shp) == 1:
            v = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif leaf == "bias":
            v = 0.1 * torch.randn(shp, generator=g)
        else:
            v = torch.randn(shp, generator=g) / math.sqrt(max(1, math.prod(shp[1:])))
        if v is not None:
            canon[key] = v
            sd[name] = v
    return sd
