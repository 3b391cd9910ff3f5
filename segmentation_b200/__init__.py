"""segmentation_b200 — B200-native (sm_100a) implementation of the segmentation
hot path of nathanin/segmentation: forward + backward of the convolutional
encoder-decoder models behind the reference's own Python surface.

    from segmentation_b200.models.unet import UNetModel
"""
__version__ = '0.1.0'
