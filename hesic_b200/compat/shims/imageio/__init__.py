"""Stand-in for ``imageio`` used only when the real package is not installed.  The drivers call ``imageio.mimsave`` in
``tensors_to_gif`` (ywz/mywork/test3real.py:51-54), a debugging helper off the evaluated path; it is implemented on
Pillow, which torchvision already requires."""
import numpy as np

_HESIC_STUB = True
__version__ = "0.0+hesic_b200.shim"


def _pil(a):
    from PIL import Image
    a = np.asarray(a)
    if a.dtype != np.uint8:
        a = (np.clip(a, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8) if a.dtype.kind == "f" else a.astype(np.uint8)
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    return Image.fromarray(a)


def imwrite(uri, im, **kwargs):
    _pil(im).save(uri)


imsave = imwrite


def imread(uri, **kwargs):
    from PIL import Image
    return np.asarray(Image.open(uri))


def mimsave(uri, ims, duration=0.1, **kwargs):
    frames = [_pil(a) for a in ims]
    if not frames:
        raise ValueError("imageio.mimsave: no frames")
    frames[0].save(uri, save_all=True, append_images=frames[1:], duration=int(float(duration) * 1000), loop=0)


mimwrite = mimsave
