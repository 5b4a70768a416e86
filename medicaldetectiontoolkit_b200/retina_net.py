"""RetinaNet = Retina U-Net without the segmentation head (the reference's models/retina_net.py differs from retina_unet.py in 5 hunks)."""
from .retina_unet import net as _unet_net, Classifier, BBRegressor, compute_class_loss, compute_bbox_loss, refine_detections, get_results  # noqa: F401


class net(_unet_net):
    has_seg_head = False
