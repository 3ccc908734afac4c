"""Light containers with the field names the hot path reads from the reference's token objects
(pybpl/objects/part.py:188-249 StrokeToken, pybpl/objects/concept.py:217-241 CharacterToken)."""
import torch

from . import ops


def vanilla_to_motor(shapes, invscales, first_pos, neval=200):
    """Fine-motor trajectory of one stroke (pybpl/objects/part.py:290-328).

    shapes (ncpt,2,nsub), invscales (nsub,), first_pos (2,) -> motor (nsub,neval,2), motor_spline (ncpt,2,nsub)."""
    for elt in [shapes, invscales, first_pos]:
        assert elt is not None
        assert isinstance(elt, torch.Tensor)
    assert len(shapes.shape) == 3
    assert shapes.shape[1] == 2
    assert len(invscales.shape) == 1
    assert first_pos.shape == torch.Size([2])
    src = shapes.device
    dev = src if shapes.is_cuda else torch.device("cuda", torch.cuda.current_device())
    nsub = shapes.shape[2]
    ctrl = shapes.permute(2, 0, 1).to(dev, torch.float32).contiguous()
    motor, ms = ops.vanilla_to_motor_batch(ctrl, invscales.to(dev, torch.float32), first_pos.to(dev, torch.float32).view(1, 2),
                                           [0, nsub], neval)
    return motor.to(src), ms.permute(1, 2, 0).to(src)


class StrokeToken:
    def __init__(self, shapes, invscales, position=None):
        self.shapes = shapes
        self.invscales = invscales
        self.position = position

    @property
    def motor(self):
        assert self.position is not None
        return vanilla_to_motor(self.shapes, self.invscales, self.position)[0]

    @property
    def motor_spline(self):
        assert self.position is not None
        return vanilla_to_motor(self.shapes, self.invscales, self.position)[1]

    def parameters(self):
        return [self.shapes, self.invscales]


class CharacterToken:
    def __init__(self, part_tokens, affine=None, epsilon=1e-4, blur_sigma=0.5):
        self.part_tokens = part_tokens
        self.affine = affine
        self.epsilon = epsilon if torch.is_tensor(epsilon) else torch.tensor(float(epsilon))
        self.blur_sigma = blur_sigma if torch.is_tensor(blur_sigma) else torch.tensor(float(blur_sigma))
