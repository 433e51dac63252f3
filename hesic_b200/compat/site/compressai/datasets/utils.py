"""Stereo ImageFolder placeholder.  Data loading (SURF + RANSAC homography, crops;
compressai/datasets/utils.py:30-214) is outside the forward hot path;
the name exists so the drivers' imports resolve.  Benchmarks and tests use hesic_b200.synth."""
from torch.utils.data import Dataset


class ImageFolder(Dataset):
    def __init__(self, root, transform=None, patch_size=(256, 256), split="train", need_file_name=False):
        raise NotImplementedError(
            "hesic_b200 replaces the HSIC forward path only; use the reference's compressai.datasets.ImageFolder "
            "for InStereo2K/KITTI loading, or hesic_b200.synth.stereo_pairs for synthetic pairs")
