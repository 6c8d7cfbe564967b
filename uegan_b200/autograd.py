"""Training path (torch.autograd.Function wrappers over the fprop / dgrad / wgrad kernels).

Not built yet in this round: the forward-only (inference) path is native; calling the Generator or the
Discriminator with gradients enabled raises instead of silently falling back to PyTorch ops."""


def generator_apply(module, x):
    raise NotImplementedError("uegan_b200: the Generator backward (dgrad/wgrad kernels) is not built yet; "
                              "call under torch.no_grad() for inference")


def discriminator_apply(module, x):
    raise NotImplementedError("uegan_b200: the Discriminator backward (dgrad/wgrad kernels) is not built yet; "
                              "call under torch.no_grad()")
