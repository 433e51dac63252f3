from .utils import ImageFolder

__all__ = ["ImageFolder"]
