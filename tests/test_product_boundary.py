"""The oracle is test infrastructure: nothing the product ships may import, call or execute it
(the C-ABI export check lives in tests/test_video_host.py)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "video-to-action-release_b200")


def _sources(d, exts):
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                yield os.path.join(base, f)


def test_product_never_touches_the_oracle_or_the_reference_checkout():
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|/root/reference|ref_import", re.M)
    for path in list(_sources(PKG, (".py", ".cu", ".cuh", ".h"))) + [os.path.join(ROOT, "v2a_b200", "__init__.py")]:
        with open(path) as f:
            m = pat.search(f.read())
        assert m is None, f"{path}: product code references test infrastructure ({m.group(0)!r})"


def test_bench_uses_the_oracle_only_in_the_cpu_baseline_and_reference_arm():
    with open(os.path.join(ROOT, "bench.py")) as f:
        src = f.read()
    # the helpers of the reference arms (`--impl reference`, cpu or cuda) and of the `cpu_baseline` leg; our arm's
    # measured path (run_ours / run_policy / measure_kernel_roofline) must stay free of them
    allowed = {"_reference_modules", "_port_video_step", "_policy_step_fn"}
    for m in re.finditer(r"^\s+from oracle\b.*$", src, re.M):
        head = src[:m.start()]
        fn = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert fn in allowed, f"bench.py imports the oracle inside {fn}()"
    assert not re.search(r"^(from|import) oracle", src, re.M)
    for fn in ("run_ours", "run_policy", "measure_kernel_roofline"):
        body = src[src.index(f"def {fn}("):]
        body = body[:body.index("\ndef ", 1)]
        assert "oracle" not in body and "_video_step_fn" not in body and "_policy_step_fn" not in body, fn
