"""Evaluation metrics under the reference's ``spair.metric`` names (metric.py:5-100).
Off the timed path; plain tensor code that runs on whatever device its inputs are on."""
import torch

from . import config as cfg


def intersect(box_a, box_b):
    """Pairwise intersection areas of [Batch, A, 4] and [Batch, B, 4] corner boxes -> [Batch, A, B]."""
    hi = torch.min(box_a[..., None, 2:], box_b[..., None, :, 2:])
    lo = torch.max(box_a[..., None, :2], box_b[..., None, :, :2])
    wh = (hi - lo).clamp(min=0)
    return wh[..., 0] * wh[..., 1]


def batch_jaccard(box_a, box_b):
    """IoU of every pair of corner boxes: [Batch, A, 4] x [Batch, B, 4] -> [Batch, A, B] (metric.py:80-100)."""
    inter = intersect(box_a, box_b)
    area_a = ((box_a[..., 2] - box_a[..., 0]) * (box_a[..., 3] - box_a[..., 1]))[..., None]
    area_b = ((box_b[..., 2] - box_b[..., 0]) * (box_b[..., 3] - box_b[..., 1]))[..., None, :]
    return inter / (area_a + area_b - inter)


def mAP(z_where, z_pres, ground_truth_bbox, truth_bbox_digit_count):
    """IoU-threshold score of metric.py:5-47: best predicted box per label box, averaged over the
    thresholds 0.1..0.9 with a linear ramp, normalised by the label count.  Like the reference it
    scales ``z_where`` and converts ``ground_truth_bbox`` to corners IN PLACE."""
    image_size = cfg.INPUT_IMAGE_SHAPE[-1]
    batch = cfg.BATCH_SIZE
    z_where *= image_size
    boxes = z_where.permute(0, 2, 3, 1).contiguous().view(batch, -1, 4)
    boxes[..., 2:] += boxes[..., :2]
    ground_truth_bbox[..., 2:] += ground_truth_bbox[..., :2]
    iou = batch_jaccard(boxes, ground_truth_bbox)
    best = iou.max(dim=1)[0].unsqueeze(-1).cpu()
    thresholds = torch.arange(0.1, 1.0, 0.1)
    ramp = ((best - thresholds) / (1 - thresholds)).clamp(0, 1)
    per_image = ramp.mean(dim=-1).sum(dim=-1, keepdim=True) / truth_bbox_digit_count.cpu()
    return per_image.mean()


def object_count_accuracy(z_pres: torch.Tensor, truth_bbox_digit_count):
    """Mean signed difference between label count and rounded presence count (metric.py:49-56)."""
    batch = cfg.BATCH_SIZE
    counts = z_pres.permute(0, 2, 3, 1).contiguous().view(batch, -1, 1).round().sum(dim=-2)
    return (truth_bbox_digit_count - counts).mean()
