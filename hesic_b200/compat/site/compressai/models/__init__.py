from .priors import CompressionModel  # noqa: F401
from .utils import conv, deconv, update_registered_buffers  # noqa: F401
