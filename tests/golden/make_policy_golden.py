"""Golden vectors for the policy path from the UNMODIFIED reference ConditionalUnet1D
(run in the build container only: python tests/golden/make_policy_golden.py)."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import as R  # noqa: E402
from oracle import policy_oracle as PO  # noqa: E402
from tests.golden.configs import POLICY_LIBERO, POLICY_TINY, grad_fingerprint, policy_inputs  # noqa: E402


def run(cfg, B, seed):
    net = R.ConditionalUnet1D()(**cfg)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(PO.seeded_policy_state_dict(shapes, seed))
    traj, noise, t, gc = policy_inputs(B, cfg, seed)
    gc = gc.clone().requires_grad_(True)
    acp = PO.ddpm_alphas_cumprod(100)
    noisy = PO.add_noise(acp, traj, noise, t).requires_grad_(True)
    pred = net(noisy, t, local_cond=None, global_cond=gc)
    loss = torch.nn.functional.mse_loss(pred, noise, reduction="none").reshape(B, -1).mean(1).mean()
    loss.backward()
    fp = {n: grad_fingerprint(n, p.grad) for n, p in net.named_parameters()}
    return shapes, dict(pred=pred.detach(), loss=loss.detach(), d_global_cond=gc.grad.clone(),
                        d_sample=noisy.grad.clone()), fp


def main():
    out, meta = {}, {}
    for name, cfg, B, seed in (("tiny", POLICY_TINY, 3, 11), ("libero", POLICY_LIBERO, 4, 12)):
        shapes, tensors, fp = run(cfg, B, seed)
        for k, v in tensors.items():
            out[f"{name}.{k}"] = v
        meta[name] = {"layout": {k: list(v) for k, v in shapes.items()}, "grad_fingerprints": fp,
                      "B": B, "seed": seed}
        print(name, "loss", tensors["loss"].item(), "params", sum(torch.tensor(s).prod().item() for s in shapes.values()))
    torch.save(out, os.path.join(HERE, "policy_golden.pt"))
    with open(os.path.join(HERE, "policy_golden_meta.json"), "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    main()
