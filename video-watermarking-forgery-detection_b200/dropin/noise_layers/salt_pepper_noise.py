"""Drop-in for the reference's noise_layers/salt_pepper_noise.py."""
from wmattack.modules import SaltPepper  # noqa: F401
