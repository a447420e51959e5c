"""The opt-in paths stay parity-green: programmatic dependent launch (MCG_PDL=1), dynamic work distribution of the
persistent fprop/dgrad kernels (MCG_TC_DYN=1) and split-K fprop (MCG_TC_SPLITK=1).  Both are read once per process, so each case runs in a child
pytest process with the flag set: the BASELINE config-2 layer sizes of the tcgen05 kernels against the float64
host convolution (only those are large enough for the dynamic distribution to switch on) and one whole update_core step
against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEP = ["tests/test_step_gpu.py::test_step_bf16_tcgen05_mug_normal"]
SELECT = {
    # every kernel of a step launched with the programmatic-serialization attribute, eager and multi-stream
    "MCG_PDL": STEP + ["tests/test_step_gpu.py::test_step_fp32_strict_infogan"],
    # the dynamic distribution only switches on when a CTA has >= 4 tile steps: the full-size layers
    "MCG_TC_DYN": ["tests/test_kernels_gpu.py::test_conv_tc_full_size_vs_float64"] + STEP,
    # split-K fprop (fp32 red.add of partial tiles + finish pass): only Dv.dc4 is large and box-poor enough to take it
    "MCG_TC_SPLITK": ["tests/test_kernels_gpu.py::test_conv_tc_full_size_vs_float64[Dv.dc4]"],
}


@pytest.mark.parametrize("flag", sorted(SELECT))
def test_parity_with_flag(flag):
    env = dict(os.environ)
    env[flag] = "1"
    out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu"] + SELECT[flag], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
