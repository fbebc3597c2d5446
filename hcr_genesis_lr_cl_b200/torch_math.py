"""Small torch helpers for the host-side (non-hot-path) plugin methods."""
import torch


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Rotate v by the inverse of q (xyzw); same convention as legged_gym/utils/math_utils.py:63-76."""
    w = q[..., 3:4]
    u = q[..., :3]
    return v * (2.0 * w * w - 1.0) - torch.cross(u, v, dim=-1) * (2.0 * w) + u * (2.0 * (u * v).sum(-1, keepdim=True))
