"""stillleben.losses: the loss helper used next to stillleben.diff (interface of python/stillleben/losses.py:5-23)."""


def neg_iou_loss(predict, target):
    """Negative intersection over union of two soft masks.

    predict, target: B x C x H x W. Returns (scalar loss = 1 - mean IoU over the batch, per-pixel visualisation
    1 - intersection / union as a detached tensor of the inputs' shape)."""
    reduce_dims = tuple(range(1, predict.ndimension()))
    inter = predict * target
    union = predict + target - inter
    iou = inter.sum(reduce_dims) / (union.sum(reduce_dims) + 1e-6)
    loss_img = (1.0 - inter / (union + 1e-6)).detach().clone()
    return 1.0 - iou.sum() / iou.nelement(), loss_img
