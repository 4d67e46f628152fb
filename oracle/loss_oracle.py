"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's PointPillarLoss
(/root/reference/opencood/loss/point_pillar_loss.py:36-116, 119-158, 201-245) in torch, differentiable so that autograd
gives the reference gradients w.r.t. the head outputs.  Pinned by tests/golden/loss_*.npz, produced by the unmodified
reference class (tests/golden/gen_golden_loss.py)."""
import numpy as np
import torch


def limit_period(val, offset=0.5, period=2 * np.pi):                     # common_utils.py:70-79
    return val - torch.floor(val / period + offset) * period


def sigmoid_focal_loss(preds, targets, weights, alpha, gamma):           # :230-245
    ce = torch.clamp(preds, min=0) - preds * targets.type_as(preds)
    ce = ce + torch.log1p(torch.exp(-torch.abs(preds)))
    p = torch.sigmoid(preds)
    p_t = targets * p + (1 - targets) * (1 - p)
    mod = torch.pow(1.0 - p_t, gamma)
    aw = targets * alpha + (1 - targets) * (1 - alpha)
    return mod * aw * ce * weights


def weighted_smooth_l1_loss(preds, targets, sigma, weights):             # :219-227
    diff = preds - targets
    ad = torch.abs(diff)
    lt = torch.le(ad, 1 / (sigma ** 2)).type_as(ad)
    loss = lt * 0.5 * torch.pow(ad * sigma, 2) + (ad - 0.5 / (sigma ** 2)) * (1.0 - lt)
    return loss * weights


def pointpillar_loss(args, cls_preds, reg_preds, dir_preds, pos_equal_one, neg_equal_one, targets):
    """Returns (total, {'reg_loss','cls_loss','dir_loss'}) as tensors; inputs as the reference gets them: NCHW float32
    predictions, (n,H,W,A) / (n,H,W,7A) label tensors (float64 from the reference's collate)."""
    n = pos_equal_one.shape[0]
    cls_labls = pos_equal_one.view(n, -1, 1)
    positives = cls_labls > 0
    negatives = neg_equal_one.view(n, -1, 1) > 0
    pos_normalizer = positives.sum(1, keepdim=True).float()
    cp = cls_preds.permute(0, 2, 3, 1).contiguous().view(n, -1, 1)
    cls_weights = positives * args["pos_cls_weight"] + negatives * 1.0
    cls_weights = cls_weights / torch.clamp(pos_normalizer, min=1.0)
    cls_loss = sigmoid_focal_loss(cp, cls_labls, cls_weights, args["cls"]["alpha"], args["cls"]["gamma"])
    cls_loss = cls_loss.sum() * args["cls"]["weight"] / n
    reg_weights = positives / torch.clamp(pos_normalizer, min=1.0)
    rp = reg_preds.permute(0, 2, 3, 1).contiguous().view(n, -1, 7)
    rt = targets.view(n, -1, 7)
    enc_p = torch.sin(rp[..., 6:7]) * torch.cos(rt[..., 6:7])               # add_sin_difference (:119-131)
    enc_t = torch.cos(rp[..., 6:7]) * torch.sin(rt[..., 6:7])
    rp2 = torch.cat([rp[..., :6], enc_p], dim=-1)
    rt2 = torch.cat([rt[..., :6], enc_t], dim=-1)
    reg_loss = weighted_smooth_l1_loss(rp2, rt2, args["reg"]["sigma"], reg_weights)
    reg_loss = reg_loss.sum() * args["reg"]["weight"] / n
    total = reg_loss + cls_loss
    parts = {"reg_loss": reg_loss, "cls_loss": cls_loss}
    if args.get("dir"):
        da = args["dir"]["args"]
        num_bins, dir_offset = da["num_bins"], da["dir_offset"]
        yaw = torch.from_numpy(np.deg2rad(np.array(da["anchor_yaw"]))).view(1, -1, 1)
        A = yaw.shape[1]
        amap = yaw.repeat(1, rt.shape[1] // A, 1)
        rot_gt = rt[..., -1] + amap[..., -1]
        off = limit_period(rot_gt - dir_offset, 0, 2 * np.pi)
        bins = torch.clamp(torch.floor(off / (2 * np.pi / num_bins)).long(), min=0, max=num_bins - 1)
        logits = dir_preds.permute(0, 2, 3, 1).contiguous().view(n, -1, num_bins).view(-1, num_bins)
        ce = torch.nn.functional.cross_entropy(logits, bins.view(-1), reduction="none")
        dir_loss = (ce.flatten() * reg_weights.flatten()).sum() * args["dir"]["weight"] / n
        total = total + dir_loss
        parts["dir_loss"] = dir_loss
    return total, parts


def loss_and_grads(args, case):
    """case: numpy dict from coalign_b200.synth.loss_case -> (losses dict of floats, grads dict of numpy arrays)."""
    c = torch.from_numpy(case["cls"]).requires_grad_(True)
    r = torch.from_numpy(case["reg"]).requires_grad_(True)
    d = torch.from_numpy(case["dir"]).requires_grad_(True)
    total, parts = pointpillar_loss(args, c, r, d, torch.from_numpy(case["pos"]), torch.from_numpy(case["neg"]),
                                    torch.from_numpy(case["tgt"]))
    total.backward()
    losses = {"total_loss": float(total.detach())}
    losses.update({k: float(v.detach()) for k, v in parts.items()})
    return losses, {"cls": c.grad.numpy(), "reg": r.grad.numpy(), "dir": d.grad.numpy()}
