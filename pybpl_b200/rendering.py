"""Drop-in for pybpl/rendering.py:render_image (and the helpers it exposes), computed by the fused CUDA
kernels.  Same signature and return convention: (pimg [H,W] float32 on strokes' device, ink_off_page), H x W =
ps.imsize (105 x 105 by default: the fused kernels; any other size: the run-time-geometry kernels)."""
import torch

from . import ops
from .parameters import Parameters

__all__ = ['check_bounds', 'space_motor_to_img', 'render_image']


def check_bounds(myt, imsize):
    """Out-of-bounds mask of image-space points (pybpl/rendering.py:10-33); tiny, plain tensor ops."""
    xmin, ymin = 0, 0
    xmax, ymax = imsize[0], imsize[1]
    x, y = myt[..., 0], myt[..., 1]
    return (torch.floor(x) < xmin) | (torch.ceil(x) >= xmax) | (torch.floor(y) < ymin) | (torch.ceil(y) >= ymax)


def space_motor_to_img(pt):
    """(x, y) motor space -> (row, col) image space (pybpl/rendering.py:36-55)."""
    return torch.stack([-pt[..., 1], pt[..., 0]], dim=-1)


def _compute_device(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def render_image(strokes, epsilon, blur_sigma, ps=None):
    """Render stroke trajectories into an image probability map (pybpl/rendering.py:185-229).

    strokes: (m,n,2) tensor or list of (n_i,2) tensors, motor space.  epsilon, blur_sigma: float or
    0-d tensor (gradients flow to them and to the strokes).  Returns (pimg, ink_off_page); ink_off_page
    is a 0-d bool tensor (bool(x) synchronises only when it is read)."""
    if ps is None:
        ps = Parameters()
    src_dev = strokes[0].device
    dev = _compute_device(strokes[0])
    if torch.is_tensor(strokes):
        assert strokes.dim() == 3 and strokes.size(-1) == 2
        m, n = strokes.shape[0], strokes.shape[1]
        pts = strokes.reshape(-1, 2)
        lens = [n] * m
    else:
        for s in strokes:
            assert s.dim() == 2 and s.size(-1) == 2
        lens = [int(s.shape[0]) for s in strokes]
        pts = torch.cat(list(strokes), dim=0)
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    pts = pts.to(dev, torch.float32)
    eps = torch.as_tensor(epsilon, dtype=torch.float32).to(dev).reshape(1)
    sig = torch.as_tensor(blur_sigma, dtype=torch.float32).to(dev).reshape(1)
    batch = ops.StrokeBatch([0, len(lens)], off, pts, eps, sig)
    pimg, _, offp = ops.render_strokes(batch, want_pimg=True, ps=ps)
    return pimg[0].to(src_dev), offp[0].to(src_dev)
