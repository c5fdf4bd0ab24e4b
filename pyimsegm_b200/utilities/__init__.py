"""utilities kept from the reference surface: only what the hot path needs (imsegm/utilities/__init__.py:39)"""


class ImageDimensionError(TypeError):
    """raised when image / segmentation shapes do not fit together (reference: imsegm/utilities/__init__.py:39)"""
