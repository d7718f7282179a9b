"""Teacher-forced per-kernel parity on the GPU: every kernel is fed identical (bf16-rounded) inputs as a torch-CPU /
numpy-oracle evaluation of the same operator and must agree within the bf16 tolerance (rel 1e-2 of the tensor's max).
This is the strict gate for the training path: end-to-end training gradients of a BatchNorm network are chaotic under
1-ulp bf16 perturbations ('conditioning'), so kernel correctness is established here, per kernel."""
import pytest

pytestmark = pytest.mark.gpu


def _probe_conv():
    from tools import gpu_probe_conv
    return gpu_probe_conv


def _probe_train():
    from tools import gpu_probe_train
    return gpu_probe_train


@pytest.mark.parametrize("idx", range(13))
def test_conv_forward_shapes(cuda_device, idx):
    m = _probe_conv()
    assert m.run_case(*m.CASES[idx])


def test_conv_forward_epilogues(cuda_device):
    m = _probe_conv()
    base = (2, 4, 14, 14, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert m.run_case("affine+res+relu", *base, True, True, True)
    assert m.run_case("stats", *base, False, False, False, True)
    assert m.run_case("block_n=64", *base, block_n=64)
    assert m.run_case("multi-tile persistent", 8, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, True, True, True)
    assert m.run_case("ragged M tail", 1, 3, 7, 9, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), True, False, True, True)


@pytest.mark.parametrize("idx", range(12))
def test_conv_wgrad_dgrad(cuda_device, idx):
    m = _probe_train()
    assert m.conv_case(*m.CONV_CASES[idx])


@pytest.mark.parametrize("idx", range(5))
def test_batchnorm_forward_backward(cuda_device, idx):
    m = _probe_train()
    assert m.bn_case(*m.BN_CASES[idx])
