"""Buffer plumbing: torch is used only as the device allocator / stream carrier."""
import numpy as np
import torch

from ._lib import KmpcError


def require_cuda():
    if not torch.cuda.is_available():
        raise KmpcError("no CUDA device: koopman_online_updated_mpc_b200 has no CPU fallback")


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def to_dev(a, shape=None):
    """Return a contiguous float64 CUDA tensor for `a` (numpy array, CPU tensor or CUDA tensor);
    host inputs are copied host->device on the current stream."""
    require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if not t.is_cuda:
        t = t.to("cuda", non_blocking=True)
    t = t.contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


def like_input(result, original):
    """Give the result back in the same kind of container the caller passed in."""
    if isinstance(original, torch.Tensor):
        return result if original.is_cuda else result.cpu()
    return result.cpu().numpy()


def ptr(t):
    return None if t is None else t.data_ptr()
