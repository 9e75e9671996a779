"""-m gpu: barrier protocol of the node projection kernel under a stalled role + the kernel watchdog.

Round 1's two-context mode dead-locked (bounded spin -> __trap -> dead CUDA context).  Root cause: the drain warps of
k_tc_proj waited on AFULL, a barrier whose producers do not wait for them, with a 1-bit phase parity; a drain warp delayed by
more than one tile found the barrier two phases ahead and waited forever (csrc/tc.cu, protocol comment).  These tests
  (a) reproduce that dead-lock deterministically with the round-1 protocol (test-only build libgnb200_oldproj.so) by stalling
      the drain warps, and show that the watchdog turns it into GNB_ERR_TIMEOUT with the CUDA context still usable (since the
      round-2 kernels changed the relative speed of the roles, the old protocol dead-locks even WITHOUT the stall - it was a
      race, not a corner case - so the usability check of that mode runs on the fp32 path);
  (b) show that the product protocol computes bit-identical results under the same stall."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe(variant, B, delay_ns, watchdog_ms):
    env = dict(os.environ)
    env.pop("GNB_DEBUG_PROJ_DRAIN_DELAY_NS", None)
    env["GNB_WATCHDOG_MS"] = str(watchdog_ms)
    if variant:
        env["GNB_LIB_VARIANT"] = variant
    else:
        env.pop("GNB_LIB_VARIANT", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "watchdog_probe.py"), str(B), str(delay_ns)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_product_protocol_survives_stalled_drain():
    # 2048 graphs: 1024 node tiles = 7 per CTA, enough for a stage to be refilled twice while the drain warps sleep
    res = _probe("", 2048, 40000, 20000)
    assert res == {"timeout": False, "equal": True, "usable": True}, res


def test_round1_protocol_deadlock_is_reproduced_and_caught():
    if not os.path.exists(os.path.join(ROOT, "graphnets.jl_b200", "libgnb200_oldproj.so")):
        pytest.skip("test-only variant library not built")
    res = _probe("oldproj", 2048, 40000, 1500)
    assert res["timeout"], "the round-1 protocol was expected to dead-lock under a stalled drain: %s" % res
    assert res["usable"], "the CUDA context must stay usable after a fired watchdog: %s" % res
