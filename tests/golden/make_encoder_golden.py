"""Golden vectors for one VisualCore (ResNet18-GN + SpatialSoftmax + Linear) from the UNMODIFIED reference classes
(run in the build container only: python tests/golden/make_encoder_golden.py).  Output features and
per-parameter gradient fingerprints of sum(features * w) on seeded inputs / weights."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import policy_oracle as PO  # noqa: E402
from tests.golden.configs import encoder_inputs, grad_fingerprint  # noqa: E402
from tests.golden.make_policy_loss_golden import build_reference_policy  # noqa: E402

KEY = "obs_encoder.key_model_map.img_obs_1."


def main():
    B, seed = 2, 33
    policy = build_reference_policy()
    layout = {k: list(v.shape) for k, v in policy.state_dict().items()}
    sd = policy.state_dict()
    sd.update(PO.seeded_full_policy_state_dict(layout, seed))
    policy.load_state_dict(sd, strict=True)
    policy.train()
    core = policy.obs_encoder.key_model_map["img_obs_1"]          # the reference's own VisualCore
    x, w = encoder_inputs(B, seed)
    torch.manual_seed(seed)                                        # SpatialSoftmax draws randn_like (noise_std = 0)
    feat = core(x)
    (feat * w).sum().backward()
    fps = {n: grad_fingerprint(KEY + n, p.grad) for n, p in core.named_parameters() if p.grad is not None}
    torch.save({"feat": feat.detach()}, os.path.join(HERE, "encoder_golden.pt"))
    with open(os.path.join(HERE, "encoder_golden_meta.json"), "w") as f:
        json.dump({"B": B, "seed": seed, "key": KEY, "grad_fingerprints": fps}, f)
    print("feat", tuple(feat.shape), float(feat.abs().mean()), "params with grad", len(fps))


if __name__ == "__main__":
    main()
