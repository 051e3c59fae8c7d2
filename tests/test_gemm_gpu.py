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
