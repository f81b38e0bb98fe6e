"""Shared seeded inputs / configs of the golden fixtures (used by make_golden.py and the tests)."""
import torch

# small UNetModel with every block kind: skip convs, identity skips, both attention
# resolutions, down/up-sampling, concat seams that straddle GroupNorm groups (128+64=192 -> 6/group)
TINY_UNET = dict(image_size=(16, 16), in_channels=6, model_channels=64, out_channels=3, num_res_blocks=2,
                 attention_resolutions=(2, 4), dropout=0, channel_mult=(1, 2, 3), conv_resample=True, dims=3,
                 num_classes=None, task_tokens=True, task_token_channels=512, use_checkpoint=False,
                 use_fp16=False, num_head_channels=32)


def tiny_inputs():
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 9, 16, 16, generator=g)
    t = torch.tensor([3, 1], dtype=torch.long)
    x_cond = torch.rand(2, 3, 16, 16, generator=g)
    te = torch.randn(2, 6, 512, generator=g)
    return x, t, x_cond, te


def p_losses_inputs():
    """img in [0, 1] and noise for the denoising-loss golden (tests/golden/make_p_losses_golden.py)."""
    g = torch.Generator().manual_seed(4321)
    img = torch.rand(2, 9, 16, 16, generator=g)
    noise = torch.randn(2, 9, 16, 16, generator=g)
    return img, noise


def config1_inputs():
    """BASELINE.json configs[0]: Unet_Libero, 64x64, 4 frames, batch 1."""
    g = torch.Generator().manual_seed(123)
    x = torch.randn(1, 12, 64, 64, generator=g)
    t = torch.tensor([57], dtype=torch.long)
    x_cond = torch.rand(1, 3, 64, 64, generator=g)
    te = torch.randn(1, 8, 512, generator=g)
    return x, t, x_cond, te


# ---- policy (ConditionalUnet1D) ----
POLICY_LIBERO = dict(input_dim=7, local_cond_dim=None, global_cond_dim=128, diffusion_step_embed_dim=128,
                     down_dims=[256, 512, 1024], kernel_size=5, n_groups=8, cond_predict_scale=True)
POLICY_TINY = dict(input_dim=7, local_cond_dim=None, global_cond_dim=32, diffusion_step_embed_dim=32,
                   down_dims=[64, 128], kernel_size=5, n_groups=8, cond_predict_scale=True)


def policy_inputs(B, cfg, seed=0, T=16):
    g = torch.Generator().manual_seed(seed)
    traj = torch.rand(B, T, cfg["input_dim"], generator=g) * 2 - 1
    noise = torch.randn(B, T, cfg["input_dim"], generator=g)
    t = torch.randint(0, 100, (B,), generator=g)
    gc = torch.randn(B, cfg["global_cond_dim"], generator=g)
    return traj, noise, t, gc


def grad_fingerprint(name, grad):
    """(L2 norm, projection on a seeded random direction) of one gradient tensor."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    r = torch.randn(grad.shape, generator=g, dtype=torch.float64)
    gd = grad.detach().double().cpu()
    return [gd.norm().item(), (gd * r).sum().item()]


def policy_loss_batch(B, seed):
    """compute_loss batch in the trainer's format (lb_online_trainer_v7.py:1296-1310)."""
    g = torch.Generator().manual_seed(seed)
    return {"obs": {"img_obs_1": torch.rand(B, 1, 3, 128, 128, generator=g),
                    "img_goal_1": torch.rand(B, 1, 3, 128, 128, generator=g)},
            "action": torch.rand(B, 16, 7, generator=g) * 2 - 1}


def encoder_inputs(B, seed):
    """One VisualCore's input (already normalised to [-1, 1]) and the weights of the scalar its gradient is taken of."""
    g = torch.Generator().manual_seed(seed + 7)
    return torch.rand(B, 3, 128, 128, generator=g) * 2 - 1, torch.randn(B, 64, generator=g)


# ---- replay buffer (row N4) ----
REPLAY = dict(max_num_unitBufs=4, max_len_uB=24, min_len_uB=5, act_seq_len=4, H=16, W=20, A=7,
              episode_lens=(9, 30, 12, 24, 7, 18), batch=32, np_seed=7, py_seed=11)


def replay_episodes():
    """Seeded synthetic episodes: (task, cam, env_idx, uint8 frames [T, H, W, 3], fp32 actions [T - 1, A]).
    Six episodes into a deque of four (the two oldest are evicted), one longer than max_len_uB (truncated)."""
    import numpy as np
    rng = np.random.default_rng(2024)
    c = REPLAY
    eps = []
    for i, T in enumerate(c["episode_lens"]):
        frames = rng.integers(0, 256, size=(T, c["H"], c["W"], 3), dtype=np.uint8)
        acts = rng.uniform(-1, 1, size=(T - 1, c["A"])).astype(np.float32)
        eps.append((f"task_{i % 3}", f"cam_{i % 2}", 100 + i, frames, acts))
    return eps
