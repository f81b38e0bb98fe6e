"""Goal-conditioned diffusion policy behind the reference's call surface.

Mirrors ``DiffusionUnetImagePolicy`` (diffuser/diffusion_policy/diffusion_unet_image_policy.py:15-283):
same constructor arguments (the observation encoder and the two noise schedulers are injected, as
``Init_Diffusion_Policy`` does at diffuser/diffusion_policy/get_dp.py:27-89), ``compute_loss(batch)``,
``predict_action(obs_dict, use_ddim)``, ``conditional_sample`` and ``.to()`` moving the normaliser.

What runs where:
  * P2-P5  ``self.model`` = v2a_b200 ``ConditionalUnet1D``: planned CUDA forward + backward (the hot path);
  * P7     ``add_noise`` and the scheduler steps: restated here — ``diffusers`` is a third-party
           dependency the reference leaves unpinned (requirements.txt:4) and is absent offline;
  * P6     the observation encoder (2x ResNet18-GroupNorm + SpatialSoftmax, 80 % of compute_loss FLOPs): the
           modules below hold its parameters under the reference's ``state_dict`` names; ``VisualCore.forward``
           runs the planned CUDA engine of ``obs_encoder.py`` (forward and backward) and nothing else.

RNG order of ``compute_loss`` follows the reference: SpatialSoftmax draws (goal, then
obs encoder, training mode only), ``randn(trajectory.shape)``, ``randint(0, T, (B,))``.
"""
from __future__ import annotations

import copy
import math
import os
import weakref
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, obs_encoder, ops, packing
from .policy_unet1d import ConditionalUnet1D


# ---------------------------------------------------------------------------
# noise draws (module-level so tests can replay a CPU stream)
# ---------------------------------------------------------------------------
def _randn(shape, device, dtype=torch.float32, generator=None):
    return torch.randn(shape, device=device, dtype=dtype, generator=generator)


def _randint(high, shape, device):
    return torch.randint(0, high, shape, device=device).long()


# ---------------------------------------------------------------------------
# schedulers (diffusers DDPMScheduler / DDIMScheduler restated for the yaml's settings:
# squaredcos_cap_v2, epsilon | sample prediction, clip_sample, fixed_small variance, eta = 0)
# config/diff_policy/lb_train_diffusion_unet_image_orn10.yaml:45-53,101-113
# ---------------------------------------------------------------------------
def squaredcos_cap_v2_betas(num_train_timesteps: int, max_beta: float = 0.999) -> torch.Tensor:
    """betas_for_alpha_bar with alpha_bar(s) = cos^2((s + .008) / 1.008 * pi / 2); in-repo twin of the
    formula: flowdiffusion/.../guided_diffusion/gaussian_diffusion.py:45-62."""
    bar = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    T = num_train_timesteps
    return torch.tensor([min(1 - bar((i + 1) / T) / bar(i / T), max_beta) for i in range(T)], dtype=torch.float32)


class _SchedulerBase:
    def __init__(self, num_train_timesteps=100, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="squaredcos_cap_v2", clip_sample=True, prediction_type="epsilon", **extra):
        if beta_schedule != "squaredcos_cap_v2":
            raise NotImplementedError("only the squaredcos_cap_v2 schedule of the Libero policy yaml is restated")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      prediction_type=prediction_type, **extra)
        self.betas = squaredcos_cap_v2_betas(num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def add_noise(self, original_samples, noise, timesteps):
        """sqrt(acp_t) x0 + sqrt(1 - acp_t) eps, coefficients broadcast over trailing dims."""
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = timesteps.to(original_samples.device)
        a, b = acp[t] ** 0.5, (1 - acp[t]) ** 0.5
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise

    def _x0(self, model_output, sample, a_t):
        pt = self.config.prediction_type
        if pt == "epsilon":
            x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        elif pt == "sample":
            x0 = model_output
        else:
            raise ValueError(f"Unsupported prediction type {pt}")
        return x0.clamp(-1.0, 1.0) if self.config.clip_sample else x0


class DDPMScheduler(_SchedulerBase):
    def __init__(self, variance_type="fixed_small", **kw):
        super().__init__(variance_type=variance_type, **kw)

    def set_timesteps(self, num_inference_steps: int):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps
        self.timesteps = torch.from_numpy((np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64))

    def step(self, model_output, timestep, sample, generator=None, **kw):
        t = int(timestep)
        n = self.num_inference_steps or self.config.num_train_timesteps
        prev = t - self.config.num_train_timesteps // n
        acp = self.alphas_cumprod
        a_t = acp[t]
        a_prev = acp[prev] if prev >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        x0 = self._x0(model_output, sample, a_t)
        mean = (a_prev ** 0.5 * cur_b / b_t) * x0 + (cur_a ** 0.5 * b_prev / b_t) * sample
        if t > 0:
            var = (b_prev / b_t * cur_b).clamp(min=1e-20)   # fixed_small
            mean = mean + var ** 0.5 * _randn(model_output.shape, model_output.device, model_output.dtype, generator)
        return SimpleNamespace(prev_sample=mean, pred_original_sample=x0)


class DDIMScheduler(_SchedulerBase):
    def __init__(self, set_alpha_to_one=True, steps_offset=0, **kw):
        super().__init__(set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, **kw)
        self.final_alpha_cumprod = self.one if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps            # "leading" spacing: [84, 72, ..., 0] for 8 of 100
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset)

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, **kw):
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        x0 = self._x0(model_output, sample, a_t)
        eps = model_output if self.config.prediction_type == "epsilon" else \
            (sample - a_t ** 0.5 * x0) / (1 - a_t) ** 0.5
        var = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        std = eta * var ** 0.5
        out = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * eps
        if eta > 0:
            out = out + std * _randn(model_output.shape, model_output.device, model_output.dtype, generator)
        return SimpleNamespace(prev_sample=out, pred_original_sample=x0)


# ---------------------------------------------------------------------------
# constant [min, max] -> [-1, 1] normalisers (diffuser/diffusion_policy/normalizer.py:11-147)
# ---------------------------------------------------------------------------
class LimitsConstNormalizer:
    def __init__(self, min_max, in_shape):
        self.mins = torch.as_tensor(np.asarray(min_max[0], dtype=np.float32)).reshape(*in_shape)
        self.maxs = torch.as_tensor(np.asarray(min_max[1], dtype=np.float32)).reshape(*in_shape)

    def __call__(self, x):
        return self.normalize(x)

    def normalize(self, x):
        return 2 * ((x - self.mins) / (self.maxs - self.mins)) - 1

    def unnormalize(self, x, eps=0):
        if x.max() > 1 + eps or x.min() < -1 - eps:
            x = torch.clamp(x, -1, 1)
        return (x + 1) / 2.0 * (self.maxs - self.mins) + self.mins


class ConstNormalizerGroup:
    """One normaliser per obs key plus 'action'; shape_meta[...]['minmax_shape'] = (min, max, shape)."""

    def __init__(self, normalizer, shape_meta, n_obs_steps):
        if isinstance(normalizer, str):
            normalizer = eval(normalizer)
        self.normalizers = {}
        entries = dict(shape_meta["obs"])
        entries["action"] = shape_meta["action"]
        for key, attr in entries.items():
            mn, mx, shp = attr["minmax_shape"]
            shp = list(shp)
            steps = 1  # horizon axis broadcasts for the action; n_obs_steps == 1 on this path
            self.normalizers[key] = normalizer((mn, mx), (shp[0], steps, *shp[1:]))

    def __call__(self, x, key):
        return self.normalize(x, key)

    def __getitem__(self, key):
        return self.normalizers[key]

    def normalize(self, x, key):
        return self.normalizers[key].normalize(x)

    def unnormalize(self, x, key):
        return self.normalizers[key].unnormalize(x)

    def normalize_d(self, obs_dict):
        return {k: self.normalize(v, k) for k, v in obs_dict.items()}

    def to_device(self, *args, **kwargs):
        for n in self.normalizers.values():
            n.mins, n.maxs = n.mins.to(*args, **kwargs), n.maxs.to(*args, **kwargs)

    @property
    def device(self):
        return next(iter(self.normalizers.values())).mins.device


# ---------------------------------------------------------------------------
# observation encoder (row P6 / N1): parameter holders with the reference's structure and state_dict names;
# the math runs in obs_encoder.py
# ---------------------------------------------------------------------------
class _AttrMixin(nn.Module):
    """common/module_attr_mixin.py:3-15 — an empty parameter pins .device/.dtype (and is in state_dict)."""

    def __init__(self):
        super().__init__()
        self._dummy_variable = nn.Parameter()

    @property
    def device(self):
        return next(iter(self.parameters())).device

    @property
    def dtype(self):
        return next(iter(self.parameters())).dtype


class ResNet18Conv(nn.Module):
    """torchvision ResNet18 trunk without avgpool/fc (common/vision_nets.py:9-39)."""

    def __init__(self, input_channel=3, pretrained=False, input_coord_conv=False):
        super().__init__()
        from torchvision import models as vision_models
        if input_coord_conv:
            raise NotImplementedError("CoordConv2d input layer is not used by the Libero policy yaml")
        net = vision_models.resnet18(weights=None)
        if input_channel != 3:
            net.conv1 = nn.Conv2d(input_channel, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.nets = nn.Sequential(*(list(net.children())[:-2]))

    def output_shape(self, input_shape):
        return [512, int(math.ceil(input_shape[1] / 32.0)), int(math.ceil(input_shape[2] / 32.0))]

    def forward(self, x):
        return self.nets(x)


class SpatialSoftmax(nn.Module):
    """1x1 conv to K keypoint maps, softmax over pixels, expected (x, y) (common/base_nets.py:153-285)."""

    def __init__(self, input_shape, num_kp=None, temperature=1.0, learnable_temperature=False,
                 output_variance=False, noise_std=0.0):
        super().__init__()
        if output_variance or learnable_temperature:
            raise NotImplementedError("the Libero yaml uses a constant temperature and no variance output")
        self._in_c, self._in_h, self._in_w = input_shape
        self.nets = nn.Conv2d(self._in_c, num_kp, kernel_size=1) if num_kp is not None else None
        self._num_kp = num_kp if num_kp is not None else self._in_c
        self.noise_std = noise_std
        self.register_buffer("temperature", torch.ones(1) * temperature)
        px, py = np.meshgrid(np.linspace(-1.0, 1.0, self._in_w), np.linspace(-1.0, 1.0, self._in_h))
        self.register_buffer("pos_x", torch.from_numpy(px.reshape(1, -1)).float())
        self.register_buffer("pos_y", torch.from_numpy(py.reshape(1, -1)).float())

    def output_shape(self, input_shape):
        return [self._num_kp, 2]

    def forward(self, feature):
        if self.nets is not None:
            feature = self.nets(feature)
        att = F.softmax(feature.reshape(-1, self._in_h * self._in_w) / self.temperature, dim=-1)
        ex = torch.sum(self.pos_x * att, dim=1, keepdim=True)
        ey = torch.sum(self.pos_y * att, dim=1, keepdim=True)
        kp = torch.cat([ex, ey], 1).view(-1, self._num_kp, 2)
        if self.training:  # drawn even when noise_std == 0 (keeps the RNG stream of the reference)
            kp = kp + _randn(kp.shape, kp.device, kp.dtype) * self.noise_std
        return kp


class VisualCore(nn.Module):
    """backbone -> pool -> flatten -> Linear (common/vision_nets.py:65-177)."""

    def __init__(self, input_shape, backbone_class="ResNet18Conv", backbone_kwargs=None, pool_class="SpatialSoftmax",
                 pool_kwargs=None, flatten=True, feature_dimension=None, **kwargs):
        super().__init__()
        if backbone_class != "ResNet18Conv" or pool_class != "SpatialSoftmax" or not flatten:
            raise NotImplementedError("VisualCore restates the ResNet18Conv + SpatialSoftmax configuration")
        self.input_shape = tuple(input_shape)
        bk = {k: v for k, v in dict(backbone_kwargs or {}).items() if k in ("pretrained", "input_coord_conv")}
        self.backbone = ResNet18Conv(input_channel=input_shape[0], pretrained=bool(bk.get("pretrained")),
                                     input_coord_conv=bool(bk.get("input_coord_conv", False)))
        feat = self.backbone.output_shape(input_shape)
        pk = {k: v for k, v in dict(pool_kwargs or {}).items()
              if k in ("num_kp", "temperature", "learnable_temperature", "output_variance", "noise_std")}
        self.pool = SpatialSoftmax(feat, **pk)
        feat = self.pool.output_shape(feat)
        layers = [self.backbone, self.pool, nn.Flatten(start_dim=1, end_dim=-1)]
        self.feature_dimension = feature_dimension
        if feature_dimension is not None:
            layers.append(nn.Linear(int(np.prod(feat)), feature_dimension))
        self.nets = nn.Sequential(*layers)
        self._out = [feature_dimension] if feature_dimension is not None else [int(np.prod(feat))]

    def output_shape(self, input_shape=None):
        return list(self._out)

    def forward(self, x):
        """Planned CUDA engine (obs_encoder.py: tcgen05 convs + GroupNorm/ReLU/MaxPool/SpatialSoftmax kernels,
        forward and backward).  There is no other path: ``self.nets`` only holds the parameters under the
        reference's names (tests and tools call ``core.nets(x)`` themselves when they want the stock-op twin)."""
        assert tuple(x.shape[-3:]) == self.input_shape
        if self.pool.noise_std != 0.0:
            raise NotImplementedError("the CUDA VisualCore covers noise_std == 0 (the Libero yaml)")
        if self.training:   # the reference draws randn_like(keypoints) * noise_std even for noise_std == 0
            _randn((x.shape[0], self.pool._num_kp, 2), x.device, torch.float32)
        return obs_encoder.visual_core_forward(self, x)


def _bn_to_gn(root: nn.Module) -> nn.Module:
    """BatchNorm2d(C) -> GroupNorm(C // 16, C) everywhere (model/multi_image_obs_encoder.py:67-74)."""
    for name, child in list(root.named_children()):
        if isinstance(child, nn.BatchNorm2d):
            setattr(root, name, nn.GroupNorm(child.num_features // 16, child.num_features))
        else:
            _bn_to_gn(child)
    return root


class MultiImageObsEncoder(_AttrMixin):
    """One independent VisualCore per rgb key, features concatenated in SORTED key order
    (model/multi_image_obs_encoder.py:11-196)."""

    def __init__(self, shape_meta: dict, rgb_model, resize_shape=None, crop_shape=None, random_crop=True,
                 use_group_norm=False, share_rgb_model=False, imagenet_norm=False, _target_=None):
        super().__init__()
        if resize_shape is not None or crop_shape is not None or share_rgb_model or imagenet_norm or not use_group_norm:
            raise NotImplementedError("MultiImageObsEncoder restates the Libero yaml setting: per-key models, "
                                      "GroupNorm, no resize / crop / imagenet normalisation")
        self.key_model_map = nn.ModuleDict()
        self.key_transform_map = nn.ModuleDict()
        self.key_shape_map, rgb_keys, low_dim_keys = {}, [], []
        for key, attr in shape_meta["obs"].items():
            self.key_shape_map[key] = tuple(attr["shape"])
            kind = attr.get("type", "low_dim")
            if kind == "rgb":
                rgb_keys.append(key)
                model = rgb_model[key] if isinstance(rgb_model, dict) else copy.deepcopy(rgb_model)
                self.key_model_map[key] = _bn_to_gn(model)
                self.key_transform_map[key] = nn.Sequential(nn.Identity(), nn.Identity(), nn.Identity())
            elif kind == "low_dim":
                low_dim_keys.append(key)
            else:
                raise RuntimeError(f"Unsupported obs type: {kind}")
        self.shape_meta = shape_meta
        self.share_rgb_model = False
        self.rgb_keys, self.low_dim_keys = sorted(rgb_keys), sorted(low_dim_keys)

    def forward(self, obs_dict):
        feats, bs = [], None
        keys = self.rgb_keys + self.low_dim_keys
        cur = torch.cuda.current_stream() if torch.cuda.is_available() else None
        joined = []
        for i, key in enumerate(keys):
            x = obs_dict[key]
            bs = x.shape[0] if bs is None else bs
            assert x.shape[0] == bs and tuple(x.shape[1:]) == self.key_shape_map[key]
            if key not in self.key_model_map:
                feats.append(x)
            elif x.is_cuda and i + 1 < len(self.rgb_keys):
                # independent encoders overlap: all but the last run on side streams (the RNG draws and the
                # Python-side order stay the reference's; autograd replays the backward on the same streams)
                s = obs_encoder.side_stream(x.device, i)
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    f = self.key_model_map[key](x)
                f.record_stream(cur)
                joined.append(s)
                feats.append(f)
            else:
                feats.append(self.key_model_map[key](x))
        for s in joined:
            cur.wait_stream(s)
        return torch.cat(feats, dim=-1)

    @torch.no_grad()
    def output_shape(self):
        """Feature width of the concatenated output (the reference probes this with a zero forward pass,
        model/multi_image_obs_encoder.py:198-212; here it is read off the modules)."""
        n = 0
        for key in self.rgb_keys + self.low_dim_keys:
            n += self.key_model_map[key].output_shape()[0] if key in self.key_model_map else self.key_shape_map[key][0]
        return torch.Size([n])


# ---------------------------------------------------------------------------
# the policy
# ---------------------------------------------------------------------------
class DiffusionUnetImagePolicy(_AttrMixin):
    def __init__(self, shape_meta: dict, noise_scheduler, noise_scheduler_ddim, obs_encoder, horizon,
                 n_action_steps, n_obs_steps, num_inference_steps=None, num_inference_steps_ddim=8,
                 obs_as_global_cond=True, diffusion_step_embed_dim=256, down_dims=(256, 512, 1024), kernel_size=5,
                 n_groups=8, cond_predict_scale=True, _target_=None, cond_unet1d_config={}, **kwargs):
        super().__init__()
        if not obs_as_global_cond:
            raise NotImplementedError("the reference asserts obs_as_global_cond on this path")
        action_shape = shape_meta["action"]["shape"]
        assert len(action_shape) == 1
        action_dim = action_shape[0]
        obs_feature_dim = obs_encoder.output_shape()[0]
        self.obs_encoder = obs_encoder
        self.model = ConditionalUnet1D(input_dim=action_dim, local_cond_dim=None,
                                       global_cond_dim=obs_feature_dim * n_obs_steps,
                                       diffusion_step_embed_dim=diffusion_step_embed_dim, down_dims=list(down_dims),
                                       kernel_size=kernel_size, n_groups=n_groups,
                                       cond_predict_scale=cond_predict_scale, cond_unet1d_config=cond_unet1d_config)
        self.noise_scheduler, self.noise_scheduler_ddim = noise_scheduler, noise_scheduler_ddim
        self.ddpm_var_temp = 1.0
        self.cond_unet1d_config = cond_unet1d_config
        self.normalizer = ConstNormalizerGroup(LimitsConstNormalizer, shape_meta, n_obs_steps)
        self.horizon, self.obs_feature_dim, self.action_dim = horizon, obs_feature_dim, action_dim
        self.n_action_steps, self.n_obs_steps, self.obs_as_global_cond = n_action_steps, n_obs_steps, True
        self.kwargs = kwargs
        if num_inference_steps is None:
            num_inference_steps = noise_scheduler.config.num_train_timesteps
        self.num_inference_steps, self.num_inference_steps_ddim = num_inference_steps, num_inference_steps_ddim

    # ---- shared: normalise + encode -> global_cond [B, Do * To] (:215-239, :148-170) ----
    def _global_cond(self, obs: Dict[str, torch.Tensor]):
        nobs = self.normalizer.normalize_d(obs)
        B = next(iter(nobs.values())).shape[0]
        To = self.n_obs_steps
        this = {k: v[:, :To, ...].reshape(-1, *v.shape[2:]) for k, v in nobs.items()}
        return self.obs_encoder(this).reshape(B, -1), B

    def conditional_sample(self, condition_data, condition_mask, local_cond=None, global_cond=None, generator=None,
                           use_ddim=False, **kwargs):
        assert (condition_mask == False).all(), "no given condition"  # noqa: E712
        trajectory = _randn(condition_data.shape, condition_data.device, condition_data.dtype, generator)
        scheduler = self.noise_scheduler_ddim if use_ddim else self.noise_scheduler
        scheduler.set_timesteps(self.num_inference_steps_ddim if use_ddim else self.num_inference_steps)
        for t in scheduler.timesteps:
            out = self.model(trajectory, t, local_cond=local_cond, global_cond=global_cond)
            trajectory = scheduler.step(out, t, trajectory, generator=generator, **kwargs).prev_sample
        return trajectory

    def predict_action(self, obs_dict: Dict[str, torch.Tensor], use_ddim=False) -> Dict[str, torch.Tensor]:
        assert "past_action" not in obs_dict
        with packing.one_content_check():       # each engine verifies its packed weights once per call
            return self._predict_action(obs_dict, use_ddim)

    def _predict_action(self, obs_dict, use_ddim):
        plan = _predict_plan_for(self, obs_dict, use_ddim)
        if plan is not None:                     # the latency path: encoders + 8 DDIM steps as ONE CUDA graph
            nsample = plan.run(obs_dict)
            action_pred = self.normalizer["action"].unnormalize(nsample[..., :self.action_dim]).detach()
            start = self.n_obs_steps - 1
            return {"action": action_pred[:, start:start + self.n_action_steps], "action_pred": action_pred}
        global_cond, B = self._global_cond(obs_dict)
        cond = torch.zeros((B, self.horizon, self.action_dim), device=self.device, dtype=self.dtype)
        nsample = self.conditional_sample(cond, torch.zeros_like(cond, dtype=torch.bool), global_cond=global_cond,
                                          use_ddim=use_ddim, **self.kwargs)
        action_pred = self.normalizer["action"].unnormalize(nsample[..., :self.action_dim]).detach()
        start = self.n_obs_steps - 1
        return {"action": action_pred[:, start:start + self.n_action_steps], "action_pred": action_pred}

    def compute_loss(self, batch: dict) -> torch.Tensor:
        assert "valid_mask" not in batch
        assert self.n_obs_steps == 1, "temporally"
        assert batch["action"].shape[-1] == self.action_dim
        global_cond, B = self._global_cond(batch["obs"])
        trajectory = self.normalizer["action"].normalize(batch["action"])
        noise = _randn(trajectory.shape, trajectory.device)
        timesteps = _randint(self.noise_scheduler.config.num_train_timesteps, (B,), trajectory.device)
        noisy = self.noise_scheduler.add_noise(trajectory, noise, timesteps)
        pred = self.model(noisy, timesteps, local_cond=None, global_cond=global_cond)
        pt = self.noise_scheduler.config.prediction_type
        if pt == "epsilon":
            target = noise
        elif pt == "sample":
            target = trajectory
        else:
            raise ValueError(f"Unsupported prediction type {pt}")
        loss = F.mse_loss(pred, target, reduction="none")
        return loss.reshape(loss.shape[0], -1).mean(dim=1).mean()

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.normalizer.to_device(*args, **kwargs)
        return self


# ---------------------------------------------------------------------------
# predict_action(use_ddim=True) as ONE CUDA graph
# ---------------------------------------------------------------------------
# The reference calls this 28-42x per exploration rollout and 75x per evaluation episode, between simulator steps
# (lb_online_trainer_v7.py:1060-1079, lb_eval_helper.py:238-298): a pure latency path.  Unfused it is 2 encoder
# graph replays + 8 x (UNet1D graph replay + ~15 elementwise torch kernels of the scheduler) + the host work
# between them.  Here the whole chain -- image normalisation, both encoders (on two streams), 8 x [timestep fill,
# UNet1D forward, fused DDIM update] -- is captured once per (policy, batch) into one graph over static buffers; a
# call copies the observation and the initial noise in, replays, and reads the trajectory out.
_PREDICT_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _predict_plan_for(policy, obs_dict, use_ddim):
    """The graph plan for this call, or None when the call is outside what it covers (DDPM sampling, eta > 0 or other
    scheduler kwargs, low-dim observation keys, encoders in training mode -- they draw SpatialSoftmax noise there --
    or V2A_NO_GRAPH): those take the general loop below, on the same kernels."""
    enc = policy.obs_encoder
    if not use_ddim or policy.kwargs or enc.low_dim_keys or os.environ.get("V2A_NO_GRAPH", "0") == "1":
        return None
    if any(enc.key_model_map[k].training for k in enc.rgb_keys):
        return None
    x = obs_dict[enc.rgb_keys[0]]
    if not x.is_cuda:
        return None              # the general path raises the "CUDA only" error
    sched = policy.noise_scheduler_ddim
    sig = (x.shape[0], str(x.device), policy.num_inference_steps_ddim, sched.config.prediction_type,
           bool(sched.config.clip_sample), policy.horizon, policy.n_obs_steps)
    per = _PREDICT_PLANS.setdefault(policy, {})
    plan = per.get(sig)
    if plan is None:
        if len(per) >= 3:
            per.pop(next(iter(per)))
        plan = per[sig] = _PredictPlan(policy, x.shape[0], x.device)
    return plan


class _PredictPlan:
    def __init__(self, policy, B: int, device):
        from . import policy_unet1d
        self.policy = weakref.ref(policy)
        enc = policy.obs_encoder
        self.B, self.device = B, device
        self.keys = list(enc.rgb_keys)
        To = policy.n_obs_steps
        self.obs = {k: torch.zeros(B, To, *enc.key_shape_map[k], dtype=torch.float32, device=device) for k in self.keys}
        self.enc = [obs_encoder.encoder_engine(enc.key_model_map[k], B * To, device) for k in self.keys]
        self.unet = policy_unet1d._policy_engine(policy.model, B, policy.horizon, device)
        sched = policy.noise_scheduler_ddim
        sched.set_timesteps(policy.num_inference_steps_ddim)
        self.steps = []          # (t, sqrt(1 - a_t), sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev)) as the scheduler's fp32 scalars
        T = sched.config.num_train_timesteps
        for t in sched.timesteps.tolist():
            prev = t - T // sched.num_inference_steps
            a_t = sched.alphas_cumprod[t]
            a_prev = sched.alphas_cumprod[prev] if prev >= 0 else sched.final_alpha_cumprod
            self.steps.append((int(t), float((1 - a_t) ** 0.5), float(a_t ** 0.5), float(a_prev ** 0.5),
                               float((1 - a_prev) ** 0.5)))
        self.pred_sample = int(sched.config.prediction_type == "sample")
        self.clip = int(bool(sched.config.clip_sample))
        self.out = torch.zeros(B, policy.horizon, policy.action_dim, dtype=torch.float32, device=device)
        # the timestep MLP (embedding -> Linear -> Mish -> Linear: 5 launches per forward) sees only t and the weights:
        # its 8 outputs live in a table refreshed with the packed weights, each forward starts after it
        self.fwd_tail = self.unet.fwd.tail(self.unet.n_temb_steps)
        self.temb_table = torch.zeros(len(self.steps), *self.unet.temb_out.shape, dtype=torch.float32, device=device)
        self.graph = None
        self._warm = False
        self._params = [p for eng in (*self.enc, self.unet) for p in eng.params if p.numel()]
        self._key = None
        self._fp = None
        self._fp_stream = torch.cuda.Stream(device=device)
        self._fp_event = torch.cuda.Event()
        self._entry_event = torch.cuda.Event()
        self._fp_host = torch.zeros(1, dtype=torch.int64).pin_memory()

    def _body(self):
        policy = self.policy()
        To, unet = policy.n_obs_steps, self.unet
        cur = torch.cuda.current_stream()
        nobs = policy.normalizer.normalize_d(self.obs)
        off, joined = 0, []
        for i, (k, eng) in enumerate(zip(self.keys, self.enc)):
            x = nobs[k][:, :To].reshape(-1, *nobs[k].shape[2:])
            width = eng.feat.shape[1]
            dst = unet.gc_in[:, off:off + width]
            off += width

            def run(eng=eng, x=x, dst=dst):
                eng.x_in.copy_(x)
                eng._run("fwd", eng.fwd, pre=(eng.stats_arena,), force_eager=True)
                eng.fwd_token += 1
                dst.copy_(eng.feat.reshape(self.B, -1))
            if i + 1 < len(self.keys):       # independent encoders overlap (fork / join inside the capture)
                s = obs_encoder.side_stream(self.device, i)
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    run()
                joined.append(s)
            else:
                run()
        for s in joined:
            cur.wait_stream(s)
        assert off == unet.gc_in.shape[1], "global_cond width disagrees with the encoders' features"
        din, rows = unet.x0.C, self.B * unet.T
        for s, (t, c1, c2, c3, c4) in enumerate(self.steps):
            unet.temb_out.copy_(self.temb_table[s])
            unet._run("fwd_tail", self.fwd_tail, force_eager=True)
            _lib.check(_lib.load().v2a_policy_ddim_step(unet.x_in_base.data_ptr(), unet.x_in_base.stride(0),
                                                        unet.out16.data_ptr(), unet.out16.stride(0), rows, din,
                                                        c1, c2, c3, c4, self.pred_sample, self.clip, ops._stream()),
                       "policy_ddim_step")
        unet.fwd_token += 1
        self.out.copy_(unet.x_in.reshape(self.B, unet.T, din))

    def _refresh(self, fp):
        unet = self.unet
        for eng in (*self.enc, unet):
            eng._wkey = None
            eng.refresh_weights()
        for s, step in enumerate(self.steps):          # timestep-MLP table (eager; only when the weights moved)
            unet.t_buf.fill_(step[0])
            for fn in list(unet.fwd)[:unet.n_temb_steps]:
                fn()
            self.temb_table[s].copy_(unet.temb_out)
        self._fp = fp

    def _host_key(self):
        ptrs = tuple(p.data_ptr() for p in self._params)
        return ptrs, (ptrs, tuple(p._version for p in self._params))

    def run(self, obs_dict):
        policy = self.policy()
        # ONE content check for the three engines: (data_ptr, _version) of every parameter on the host + one
        # fingerprint launch over all of them (a `.data` write bumps no version, packing.content_key); the engines
        # re-pack only when that key moves or one of them was invalidated explicitly (fused optimiser step).
        cur = torch.cuda.current_stream()
        speculate = self.graph is not None and self._fp is not None and \
            not any(eng._wkey is None for eng in (*self.enc, self.unet))
        if speculate:
            self._entry_event.record(cur)             # parameter writes enqueued before this call end here
        else:                                         # first calls / explicit invalidation: check first, synchronously
            ptrs, hostkey = self._host_key()
            fp = ops.params_fingerprint(self._params, key=ptrs)
            if hostkey != self._key or fp != self._fp or any(eng._wkey is None for eng in (*self.enc, self.unet)):
                self._refresh(fp)
            self._key = hostkey
        for k in self.keys:
            self.obs[k].copy_(obs_dict[k][:, :policy.n_obs_steps])
        # RNG order of the reference: the encoders draw nothing in eval mode, then ONE randn for the trajectory
        noise = _randn((self.B, policy.horizon, policy.action_dim), self.device, torch.float32)
        self.unet.x_in.copy_(noise.reshape(self.B * self.unet.T, -1))
        if not self._warm:            # first call: eager (lazy module loads / plan-time launches must not be captured)
            self._body()
            self._warm = True
        else:
            if self.graph is None:
                saved = self.unet.x_in_base.clone()
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        self._body()
                cur.wait_stream(side)
                self.unet.x_in_base.copy_(saved)      # capture does not execute, but keep the inputs explicit
                self.graph = g
            self.graph.replay()
            if speculate:
                # Steady state: the whole content check -- the host walk over the parameters, the fingerprint kernel
                # (0.08 ms) and its read-back -- is off the critical path: issued AFTER the graph, the kernel on a side
                # stream that only waits for the work queued before this call, so it overlaps the replay, which ran
                # on the weights packed last time.  Only if the key differs (a `.data` update or an in-place write
                # since the last call) are the weights re-packed and the graph replayed again; the result is
                # returned after the check, so a stale answer is never handed out.
                ptrs, hostkey = self._host_key()
                self._fp_stream.wait_event(self._entry_event)
                with torch.cuda.stream(self._fp_stream):
                    dev_fp = ops.params_fingerprint(self._params, key=ptrs, launch_only=True)
                    self._fp_host.copy_(dev_fp, non_blocking=True)
                    self._fp_event.record(self._fp_stream)
                self._fp_event.synchronize()
                fp = int(self._fp_host.item())
                if fp != self._fp or hostkey != self._key:
                    self._refresh(fp)
                    self._key = hostkey
                    self.unet.x_in.copy_(noise.reshape(self.B * self.unet.T, -1))
                    self.graph.replay()
        return self.out.clone()


# ---------------------------------------------------------------------------
# the shipped Libero policy (config/diff_policy/lb_train_diffusion_unet_image_orn10.yaml) without OmegaConf
# ---------------------------------------------------------------------------
def libero_shape_meta() -> dict:
    img = (np.zeros(3, np.float32), np.ones(3, np.float32), [1, 3, 1, 1])           # image_minmax_01
    act = (-np.ones(7, np.float32), np.ones(7, np.float32), [1, 7])                 # lb_action_minmax (+-1)
    return {"obs": {"img_obs_1": {"shape": [3, 128, 128], "minmax_shape": img, "type": "rgb"},
                    "img_goal_1": {"shape": [3, 128, 128], "minmax_shape": img, "type": "rgb"}},
            "action": {"shape": [7], "minmax_shape": act}}


def build_libero_policy() -> DiffusionUnetImagePolicy:
    meta = libero_shape_meta()
    core = VisualCore(input_shape=[3, 128, 128], backbone_class="ResNet18Conv",
                      backbone_kwargs=dict(pretrained=None, input_coord_conv=False), pool_class="SpatialSoftmax",
                      pool_kwargs=dict(num_kp=32, learnable_temperature=False, temperature=1.0, noise_std=0.0,
                                       output_variance=False), flatten=True, feature_dimension=64)
    enc = MultiImageObsEncoder(meta, core, use_group_norm=True)
    sched = dict(num_train_timesteps=100, beta_start=0.0001, beta_end=0.02, beta_schedule="squaredcos_cap_v2",
                 clip_sample=True, prediction_type="epsilon")
    return DiffusionUnetImagePolicy(meta, DDPMScheduler(variance_type="fixed_small", **sched),
                                    DDIMScheduler(set_alpha_to_one=True, steps_offset=0, **sched), enc, horizon=16,
                                    n_action_steps=8, n_obs_steps=1, num_inference_steps=100,
                                    num_inference_steps_ddim=8, obs_as_global_cond=True,
                                    diffusion_step_embed_dim=128, down_dims=[256, 512, 1024], kernel_size=5,
                                    n_groups=8, cond_predict_scale=True)
