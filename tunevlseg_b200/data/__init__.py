"""Device-side input transforms."""
from .data_collator import CustomDataCollatorWithPadding
from .gpu_transforms import (GpuEvalTransforms, GpuTrainTransforms, affine_matrix, affine_walk_tables, brightness_contrast_lut, cubic_tables,
                             nearest_table, warp_cubic_table)

__all__ = ["CustomDataCollatorWithPadding", "GpuEvalTransforms", "GpuTrainTransforms", "affine_matrix", "affine_walk_tables", "brightness_contrast_lut", "cubic_tables",
           "nearest_table", "warp_cubic_table"]
