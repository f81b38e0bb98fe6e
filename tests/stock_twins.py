"""Stock-torch-op twins of product sub-paths, evaluated on the product's own parameter-holder modules.

Test infrastructure only (the product has no torch-op path: these lived behind V2A_ATTNPOOL=torch /
V2A_ENCODER=torch switches inside the package in round 1 and were moved out).  They give the tests and the
A/B tools a second, independent evaluation of the same parameters.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _task_pool(seq: nn.Sequential, y: torch.Tensor) -> torch.Tensor:
    """task_attnpool(y).mean(1) (gd/unet.py:491-494,671; gd/imagen.py:254-372).

    Step-invariant conditioning: evaluated ONCE per sample() call (0.92 GFLOP vs
    2.1 TFLOP per denoise step), with stock torch ops on the parameter tensors.
    """
    pr = seq[0]
    B, n, D = y.shape
    xp = y + pr.pos_emb.weight[:n]
    lat = pr.latents.unsqueeze(0).expand(B, -1, -1)

    def gln(x, g):
        var = x.var(dim=-1, unbiased=False, keepdim=True)
        return (x - x.mean(dim=-1, keepdim=True)) * (var + 1e-5).rsqrt() * g
    if pr.to_latents_from_mean_pooled_seq is not None:
        mp = pr.to_latents_from_mean_pooled_seq
        pooled = F.linear(gln(y.mean(dim=1), mp[0].g), mp[1].weight, mp[1].bias)
        lat = torch.cat([pooled.reshape(B, -1, D), lat], dim=1)
    for attn, ff in pr.layers:
        h = attn.heads
        xn = F.layer_norm(xp, (D,), attn.norm.weight, attn.norm.bias)
        ln = F.layer_norm(lat, (D,), attn.norm_latents.weight, attn.norm_latents.bias)
        q = F.linear(ln, attn.to_q.weight)
        k, v = F.linear(torch.cat([xn, ln], dim=1), attn.to_kv.weight).chunk(2, dim=-1)
        sp = lambda t: t.reshape(B, t.shape[1], h, -1).permute(0, 2, 1, 3)
        q, k, v = sp(q), sp(k), sp(v)
        q = F.normalize(q, dim=-1) * attn.q_scale
        k = F.normalize(k, dim=-1) * attn.k_scale
        att = (torch.einsum("bhid,bhjd->bhij", q, k) * attn.scale).softmax(dim=-1)
        o = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).reshape(B, lat.shape[1], -1)
        o = F.layer_norm(F.linear(o, attn.to_out[0].weight), (D,), attn.to_out[1].weight, attn.to_out[1].bias)
        lat = o + lat
        hdn = F.linear(gln(lat, ff[0].g), ff[1].weight)
        lat = F.linear(gln(F.gelu(hdn), ff[3].g), ff[4].weight) + lat
    return F.linear(lat, seq[1].weight, seq[1].bias).mean(dim=1)


def visual_core_stock(core, x: torch.Tensor) -> torch.Tensor:
    """`VisualCore` through its parameter-holding stock modules (ResNet18-GN -> SpatialSoftmax -> Flatten ->
    Linear; dp/common/vision_nets.py:65-177) -- what `V2A_ENCODER=torch` used to run."""
    return core.nets(x)


def global_cond_stock(policy, obs: dict):
    """`DiffusionUnetImagePolicy._global_cond` with every encoder on its stock modules (CPU capable)."""
    nobs = policy.normalizer.normalize_d(obs)
    B = next(iter(nobs.values())).shape[0]
    To = policy.n_obs_steps
    enc = policy.obs_encoder
    feats = []
    for key in enc.rgb_keys + enc.low_dim_keys:
        v = nobs[key][:, :To, ...].reshape(-1, *nobs[key].shape[2:])
        if key in enc.key_model_map:
            core = enc.key_model_map[key]
            if core.training:      # the reference draws randn_like(keypoints) * noise_std even for noise_std == 0
                torch.randn((v.shape[0], core.pool._num_kp, 2))
            feats.append(core.nets(v))
        else:
            feats.append(v)
    return torch.cat(feats, dim=-1).reshape(B, -1), B
