from .. import Layer


class _ConvNd(Layer):  # only subclassed at import time by squeezeformer/conv2d.py (Conv2DValid, unused by the shipped configs)
    def __init__(self, *a, **k):
        raise NotImplementedError("Conv2DValid is not part of the hot path")
