"""/root/reference/textboost/augment/paired_augmentation.py under its own module path."""
from textboost_b200.augment import (PairedAugmentation, adjust_brightness, adjust_scale, crop, grayscale,  # noqa: F401
                                    horizontal_flip, horizontal_translate, jpeg_compression, random_resized_crop,
                                    rotate, square_photo_collage)
