"""B200-native (sm_100a) implementation of the 3D U-Net training / sliding-window-inference hot path of
icrdr/3D-UNet-Renal-Anatomy-Extraction, behind the reference's own Python API.

    from unet3d_b200 import ResUnet3D, DiceLoss, HybirdLoss, Trainer, predict_per_patch
"""
from .network import (ResUnet3D, ResAttrUnet3D, ResAttrUnet3D2, ResAttrBNUnet3D, UNet3D, Unet, ResBlock, ResBlockStack, ConvBlock,
                      ConvBlockStack, MaxPoolBlock, AttBlock, ConvTrans3D, UpConcat, generate_paired_features,
                      generate_paired_features2)
from .loss import DiceLoss, FocalLoss, HybirdLoss, Dice, dice
from .trainer import Trainer, DevicePrefetcher, predict_per_patch, predict_case, cascade_predict_case, evaluate_case, tile_centres, tile_origins, gaussian_window, center_pad_crop, pad_to_patch
from . import parallel
from . import transform
from . import augment
from .transform import (rescale, resize, resample_normalize_case, get_spacing, apply_scale, apply_translate,
                        regions_crop_case, crop_pad_to_bbox, remove_small_region)
from .graph import GraphedTrainStep

__all__ = ["ResUnet3D", "ResAttrUnet3D", "ResAttrUnet3D2", "ResAttrBNUnet3D", "UNet3D", "Unet", "ResBlock", "ResBlockStack", "ConvBlock",
           "ConvBlockStack", "MaxPoolBlock", "AttBlock", "ConvTrans3D", "UpConcat", "generate_paired_features",
           "generate_paired_features2", "DiceLoss", "FocalLoss", "HybirdLoss", "Dice", "dice", "Trainer", "DevicePrefetcher",
           "predict_per_patch", "predict_case", "cascade_predict_case", "evaluate_case", "augment", "regions_crop_case", "crop_pad_to_bbox", "apply_translate", "rescale", "resize", "resample_normalize_case", "get_spacing",
           "apply_scale", "remove_small_region", "transform", "tile_centres", "tile_origins", "gaussian_window", "center_pad_crop", "pad_to_patch", "parallel", "GraphedTrainStep"]
