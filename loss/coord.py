"""Drop-in twin of the reference's `loss.coord` (scene-coordinate regression loss, training entry of the hot path).

Same public functions, argument order and return values as /root/reference/loss/coord.py
(`get_cam_mat` :7-17, `scene_coords_regression_loss` :87-188), called from
/root/reference/train_single_task.py:279-283.  The arithmetic is a handful of element-wise passes over
[B, 3, 5400] tensors (negligible next to the network) and stays in torch so that
autograd flows into the coordinate map; tensors are created on the inputs' device instead of the reference's
hard-wired `.cuda()`.
"""
import os

import torch

try:
    from utils.io import safe_printout
except Exception:   # pragma: no cover - reference utils not on the path
    def safe_printout(words):
        print(words)

_TINY = 1.e-7
# CROSSLOC_B200_LOSS_SYNC_FREE=1 (or loss.coord.SYNC_FREE = True): no device -> host synchronisation inside the loss.  The
# reference reads three scalars back per step (loss/coord.py:132 `.cpu().numpy()`, :168-175 `.item()` for its progress
# line), which stalls the launch queue of a GPU that trains at 300 frames/s.  In this mode the progress line is skipped,
# the "any valid prediction" branch becomes a device-side select and valid_pred_rate is returned as a 0-dim tensor
# (`float(rate)` gives the reference's number when the caller wants it).  Same loss value, bit for bit.
SYNC_FREE = os.environ.get('CROSSLOC_B200_LOSS_SYNC_FREE', '0') == '1'


def get_cam_mat(width, height, focal_length):
    """3x3 intrinsics with the principal point at the image centre (loss/coord.py:7-17)."""
    device = 'cuda' if torch.cuda.is_available() else 'cpu'
    k = torch.eye(3, device=device)
    k[0, 0] = focal_length
    k[1, 1] = focal_length
    k[0, 2] = width / 2
    k[1, 2] = height / 2
    return k


def _valid_labels(coords, nodata_value):
    """[B, N] mask of pixels whose label carries no NODATA entry (utils/learning.py:49-71)."""
    return (coords == nodata_value).sum(dim=1) == 0


def coords_world_to_cam(scene_coords, gt_coords, gt_poses):
    """World -> camera for predictions and labels; gt_poses are camera-to-world (loss/coord.py:20-38)."""
    world_to_cam = gt_poses.inverse()[:, 0:3, :]
    rot, trans = world_to_cam[:, :, 0:3], world_to_cam[:, :, 3:4]
    return torch.bmm(rot, scene_coords) + trans, torch.bmm(rot, gt_coords) + trans


def get_repro_err(camera_coords, cam_mat, pixel_grid_crop, min_depth):
    """Pixel reprojection error with the depth clamped to min_depth (loss/coord.py:41-57)."""
    proj = torch.matmul(cam_mat.to(camera_coords.dtype), camera_coords)
    depth = proj[:, 2:3].clamp(min=min_depth)
    err = proj[:, 0:2] / depth - pixel_grid_crop[None]
    return err.norm(p=2, dim=1).clamp(min=_TINY)


def check_constraints(camera_coords, reproj_error, cam_coords_reg_error, mask_gt_coords_nodata, min_depth,
                      max_reproj_error, max_coords_reg_error):
    """Pixels whose prediction is in front of the camera, reprojects within the hard clamp and lies within the
    initial tolerance of a known label (loss/coord.py:60-84)."""
    too_close = camera_coords[:, 2] < min_depth
    too_far_off = reproj_error > max_reproj_error
    off_label = (cam_coords_reg_error > max_coords_reg_error) & ~mask_gt_coords_nodata
    return ~(too_close | too_far_off | off_label)


def scene_coords_regression_loss(min_depth, soft_clamp, hard_clamp, init_tolerance, uncertainty, pixel_grid,
                                 nodata_value, cam_mat, scene_coords, uncertainty_map, gt_poses, gt_coords,
                                 reduction='mean'):
    """Reprojection + (MLE-weighted) 3-D distance loss of loss/coord.py:87-188.  Returns (loss, valid_pred_rate)."""
    h, w = scene_coords.size(2), scene_coords.size(3)
    grid = pixel_grid[:, 0:h, 0:w].reshape(2, -1).to(scene_coords.device)
    pred = scene_coords.reshape(scene_coords.size(0), 3, -1)
    label = gt_coords.reshape(gt_coords.size(0), 3, -1)

    cam_pred, cam_label = coords_world_to_cam(pred, label, gt_poses)
    dist3d = torch.norm(cam_pred - cam_label, dim=1, p=2)
    reproj = get_repro_err(cam_pred, cam_mat, grid, min_depth)

    has_label = _valid_labels(label[:, :3, :], nodata_value)
    valid = check_constraints(cam_pred, reproj, dist3d, ~has_label, min_depth, hard_clamp, init_tolerance)
    sync_free = SYNC_FREE
    num_valid = valid.sum() if sync_free else valid.sum(dim=1).cpu().numpy()
    pixels_batch = valid.numel()
    pixels_instance = valid[0].numel()

    # soft-clamped L1 of the reprojection error over valid predictions: linear up to soft_clamp, sqrt beyond
    loss_reproj = 0
    if sync_free or num_valid.sum() > 0:
        reproj = reproj * valid
        linear = (reproj * (reproj <= soft_clamp)).clamp(min=_TINY)
        root = (reproj * (reproj > soft_clamp)).clamp(min=_TINY)
        root = torch.sqrt(soft_clamp * root + _TINY).clamp(min=_TINY)
        loss_reproj = linear + root
        if sync_free:   # the reference adds nothing when no prediction is valid
            loss_reproj = torch.where(num_valid > 0, loss_reproj, torch.zeros_like(loss_reproj))

    if uncertainty is None:
        loss = torch.sum(dist3d * has_label + loss_reproj, dim=1)
    elif uncertainty == 'MLE':
        sigma = uncertainty_map.reshape(uncertainty_map.size(0), -1).clamp(min=_TINY)
        sq = dist3d.square().clamp(min=_TINY)
        nll = 3.0 * torch.log(sigma) + sq / (2.0 * sigma.square().clamp(min=_TINY))
        loss = torch.sum(nll * has_label + loss_reproj, dim=1)
        if not sync_free:
            safe_printout('Regression error: coord:  %.2f, reprojection:  %.2f' % (
                torch.sum(dist3d * has_label).item() / max(1, has_label.sum().item()),
                torch.sum(reproj * valid).item() / max(1, valid.sum().item())))
    else:
        raise NotImplementedError

    valid_pred_rate = num_valid / pixels_batch if sync_free else num_valid.sum() / pixels_batch
    if reduction is None:
        loss = loss / pixels_instance
    elif reduction == 'mean':
        loss = loss.sum() / pixels_batch
    else:
        raise NotImplementedError
    return loss, valid_pred_rate
