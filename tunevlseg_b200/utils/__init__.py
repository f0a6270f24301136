from .save_utils import resize_to_png_array, save_predictions  # noqa: F401
