"""Drop-in mirrors of the reference's `models` package (models/__init__.py:6-16): same class names, constructor
arguments, parameter names / shapes (checkpoints interchange) and call signatures; forward and backward run in
libofdmgan."""
from .discriminator import (ConditionalDiscriminator, Discriminator, MiniDiscriminator, compute_gradient_penalty,
                            create_discriminator)
from .generator import ConvBlock, MiniGenerator, UNetGenerator, create_generator

__all__ = ["MiniGenerator", "UNetGenerator", "ConvBlock", "create_generator", "MiniDiscriminator", "Discriminator",
           "ConditionalDiscriminator", "compute_gradient_penalty", "create_discriminator"]
