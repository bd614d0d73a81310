"""Small helpers shared by the shift wrappers (host-side mirror of rubiksnet/utils.py)."""
import torch

__all__ = ["make_tuple", "allocate_output"]


def make_tuple(elem, repeats):
    """3 -> [3, 3, 3]; sequences are validated and converted to ints (rubiksnet/utils.py:4-12)."""
    if isinstance(elem, int):
        return [elem] * repeats
    vals = [int(v) for v in elem]
    assert len(vals) == repeats, "expected %d values, got %r" % (repeats, elem)
    return vals


def allocate_output(output, tensor_like, desired_shape):
    """Returns `output` after checking shape/dtype/device, or a fresh tensor (rubiksnet/utils.py:15-45).

    The reference allocates with new_zeros because its 2D quantize kernels only write in-bounds taps;
    librubiks_b200 writes every element, so a fresh buffer is torch.empty (one memset less per call).
    """
    desired_shape = tuple(int(d) for d in desired_shape)
    if output is None:
        return torch.empty(desired_shape, dtype=tensor_like.dtype, device=tensor_like.device)
    assert torch.is_tensor(output)
    assert tuple(output.size()) == desired_shape, "output tensor has wrong shape {}, which should be {}".format(
        tuple(output.size()), desired_shape)
    assert output.dtype == tensor_like.dtype, "output tensor has wrong dtype {}, which should be {}".format(
        output.dtype, tensor_like.dtype)
    assert output.device == tensor_like.device, "output tensor has wrong device {}, which should be {}".format(
        output.device, tensor_like.device)
    return output
