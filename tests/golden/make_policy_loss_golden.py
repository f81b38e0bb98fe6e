"""Golden vectors for DiffusionUnetImagePolicy.compute_loss from the UNMODIFIED reference policy class,
observation encoder and normaliser (run in the build container only:
python tests/golden/make_policy_loss_golden.py).  The diffusers schedulers are absent offline; the
reference class receives the restated scheduler objects of v2a_b200.diffusion_policy (third-party
boundary: parity unpinned by reference tests, see oracle/policy_oracle.py)."""
import importlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import as R  # noqa: E402
from oracle import policy_oracle as PO  # noqa: E402
from tests.golden.configs import grad_fingerprint, policy_loss_batch  # noqa: E402


_E = "obs_encoder.key_model_map."
FULL_GRADS = [_E + "img_obs_1.backbone.nets.0.weight", _E + "img_obs_1.backbone.nets.4.0.conv1.weight",
              _E + "img_obs_1.backbone.nets.5.0.conv1.weight", _E + "img_obs_1.backbone.nets.7.1.conv2.weight",
              _E + "img_obs_1.pool.nets.weight",
              _E + "img_obs_1.nets.3.weight", _E + "img_goal_1.backbone.nets.0.weight",
              "model.down_modules.0.0.blocks.0.block.0.weight", "model.final_conv.1.weight"]


FULL_GRAD_ROWS = 16


build_reference_policy = R.build_reference_policy


def main():
    B, seed = 2, 21
    policy = build_reference_policy()
    layout = {k: list(v.shape) for k, v in policy.state_dict().items()}
    sd = policy.state_dict()
    sd.update(PO.seeded_full_policy_state_dict(layout, seed))
    policy.load_state_dict(sd, strict=True)
    policy.train()
    batch = policy_loss_batch(B, seed)
    torch.manual_seed(seed)                       # the RNG stream compute_loss consumes
    loss = policy.compute_loss(batch)
    loss.backward()
    fps = {n: grad_fingerprint(n, p.grad) for n, p in policy.named_parameters() if p.grad is not None}
    # inference entry: 8-step DDIM from a seeded stream, eval mode
    policy.eval()
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        act = policy.predict_action(batch["obs"], use_ddim=True)
    out = {"loss": loss.detach(), "action": act["action"], "action_pred": act["action_pred"]}
    # whole gradient tensors of a few parameters (a fingerprint = norm + one projection would not catch a
    # wrong-but-same-norm gradient): the stem, an early and a late ResNet conv, the keypoint conv and the
    # Linear of one encoder, the stem of the other, and two UNet1D weights
    grads = dict(policy.named_parameters())
    for name in FULL_GRADS:
        out["grad." + name] = grads[name].grad.detach()[:FULL_GRAD_ROWS].clone()   # leading output channels (file size)
    torch.save(out, os.path.join(HERE, "policy_loss_golden.pt"))
    with open(os.path.join(HERE, "policy_loss_golden_meta.json"), "w") as f:
        json.dump({"layout": layout, "grad_fingerprints": fps, "B": B, "seed": seed}, f)
    print("loss", loss.item(), "params with grad", len(fps), "action", tuple(act["action"].shape))


if __name__ == "__main__":
    main()
