"""Differentiable UNet forward (training path).  Filled in by the training milestone."""
from . import _lib


def unet_forward_train(model, x, t, cond, cond_mask):
    raise _lib.DiffsgError("diffsg_b200: the training forward/backward kernels are not built yet; "
                           "call the model under torch.no_grad() for inference")
