"""Device-side input transforms."""
from .gpu_transforms import (GpuEvalTransforms, GpuTrainTransforms, affine_matrix, affine_walk_tables, brightness_contrast_lut, cubic_tables,
                             nearest_table, warp_cubic_table)

__all__ = ["GpuEvalTransforms", "GpuTrainTransforms", "affine_matrix", "affine_walk_tables", "brightness_contrast_lut", "cubic_tables",
           "nearest_table", "warp_cubic_table"]
