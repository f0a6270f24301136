"""Device-side input transforms."""
from .gpu_transforms import GpuEvalTransforms, cubic_tables, nearest_table

__all__ = ["GpuEvalTransforms", "cubic_tables", "nearest_table"]
