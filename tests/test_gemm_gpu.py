"""-m gpu: tcgen05 GEMM parity through the C-ABI op hook against torch fp32 matmul of the same bf16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    import tools.gpu_check_gemm as G
    return [c[0] for c in G.CASES]


@pytest.mark.parametrize("name", _cases())
def test_gemm_case(name, built_lib):
    import tools.gpu_check_gemm as G
    assert torch.cuda.is_available()
    res = G.run_case(name)
    assert res["ok"], res


@pytest.mark.parametrize("name", ["pair_kk", "pair_mnAB", "pair_ragged", "pair_odd_blocks", "pair_batched", "pair_epi_resid",
                                  "pair_epi_dropout"])
def test_gemm_cta_pair_mode(name, built_lib):
    """the opt-in cta_group::2 (cluster of two CTAs, 256 x 256 tile) variant: the switch is read once per process, so
    every case runs in its own interpreter with P5_GEMM_PAIR=1"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, P5_GEMM_PAIR="1")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_check_gemm.py"), "--case", name], capture_output=True,
                       text=True, timeout=300, env=env)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, (p.stdout[-500:], p.stderr[-500:])
    res = json.loads(line[-1][7:])
    assert res["ok"], res
