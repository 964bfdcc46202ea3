"""Host side of the DSAC* solver: tensor plumbing around cl_dsac_forward_rgb / cl_dsac_backward_rgb.

Mirrors the operator interface of the reference extension
(/root/reference/dsacstar/dsacstar.cpp:63-73, :888 -- `dsacstar.forward_rgb`) and adds the batched
form the bench and the multi-GPU evaluation use.  PyTorch is only used for memory and streams.
"""
import os

import numpy as np
import torch

from . import _lib

MAX_HYPOTHESES_TRIES = 1000000   # /root/reference/dsacstar/dsacstar.cpp:48
MAX_REF_STEPS = 100              # /root/reference/dsacstar/dsacstar.cpp:47

_state = {
    'seed': int(os.environ.get('CROSSLOC_B200_SEED', '1305')),   # thread_rand.h default seed
    'image_index': 0,   # the reference's RNG stream continues across calls (thread_rand.cpp:17); so does this index
}


def set_seed(seed, image_index=0):
    """Re-key the sampler: every (seed, image index, hypothesis, try) names one fixed draw."""
    _state['seed'] = int(seed)
    _state['image_index'] = int(image_index)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream_for(t):
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return None


def forward_rgb_batch(coords, out_pose, hyps, thr, focal, cx, cy, alpha, max_reproj, subsample,
                      seed=None, image_base=None, max_tries=MAX_HYPOTHESES_TRIES, refine=True,
                      forced_samples=None, debug=False):
    """Localize a batch.  coords [B,3,Hc,Wc] float32 (CPU or CUDA), out_pose [B,4,4] float32 written in place.

    `focal` is a float or a [B] float32 tensor.  With debug=True returns a dict of host tensors
    (best, scores, hyps_rt, tries, refine_counts, rt), otherwise None.
    """
    lib = _lib.load()
    if coords.dim() != 4 or coords.size(1) != 3:
        raise RuntimeError('scene coordinates must be [B, 3, H, W], got %s' % (tuple(coords.shape),))
    if coords.dtype != torch.float32 or out_pose.dtype != torch.float32:
        raise RuntimeError('scene coordinates and output pose must be float32')
    B, _, Hc, Wc = coords.shape
    if tuple(out_pose.shape) != (B, 4, 4) or not out_pose.is_contiguous():
        raise RuntimeError('output pose must be a contiguous [B, 4, 4] tensor')
    coords_c = coords.contiguous()
    dev = coords_c.device
    if coords_c.is_cuda and not torch.cuda.is_available():
        raise RuntimeError('CUDA tensor without a CUDA device')
    if torch.is_tensor(focal):
        focal_t = focal.to(torch.float32).reshape(-1).contiguous()
        if focal_t.numel() == 1 and B > 1:
            focal_t = focal_t.expand(B).contiguous()
    else:
        focal_t = torch.full((B,), float(focal), dtype=torch.float32)
    if focal_t.numel() != B:
        raise RuntimeError('focal must be a scalar or hold one value per image')
    if seed is None:
        seed = _state['seed']
    if image_base is None:
        image_base = _state['image_index']
        _state['image_index'] += B
    forced_t = None
    if forced_samples is not None:
        forced_t = torch.as_tensor(forced_samples, dtype=torch.int32).reshape(B, hyps, 4, 2).contiguous()

    dbg = None
    if debug:
        dbg = {
            'best': torch.zeros(B, dtype=torch.int32),
            'scores': torch.zeros(B, hyps, dtype=torch.float64),
            'hyps_rt': torch.zeros(B, hyps, 6, dtype=torch.float64),
            'tries': torch.zeros(B, hyps, dtype=torch.int32),
            'refine_counts': torch.zeros(B, MAX_REF_STEPS, dtype=torch.int32),
            'rt': torch.zeros(B, 6, dtype=torch.float64),
        }
    if coords_c.is_cuda:
        torch.cuda.set_device(dev)
    code = lib.cl_dsac_forward_rgb(
        _ptr(coords_c), B, Hc, Wc, _ptr(out_pose), int(hyps), float(thr), _ptr(focal_t), float(cx), float(cy),
        float(alpha), float(max_reproj), int(subsample), int(seed) & 0xFFFFFFFFFFFFFFFF, int(image_base) & 0xFFFFFFFF,
        int(max_tries), 1 if refine else 0, _ptr(forced_t),
        _ptr(dbg['best']) if dbg else None, _ptr(dbg['scores']) if dbg else None,
        _ptr(dbg['hyps_rt']) if dbg else None, _ptr(dbg['tries']) if dbg else None,
        _ptr(dbg['refine_counts']) if dbg else None, _ptr(dbg['rt']) if dbg else None,
        _stream_for(coords_c))
    _lib.check(code)
    return dbg


def forward_rgb(scene_coordinates, out_pose, ransac_hypotheses, inlier_threshold, focal_length, ppoint_x, ppoint_y,
                inlier_alpha, max_reproj, sub_sampling):
    """`dsacstar.forward_rgb` (/root/reference/dsacstar/dsacstar.cpp:63-73): [1,3,Hc,Wc] map -> [4,4] pose in place.

    Also accepts CUDA tensors and the batched [B,3,Hc,Wc] / [B,4,4] form.  Never prints (the reference's
    eleven std::cout lines per call are dropped).
    """
    if out_pose.dim() == 2:
        if scene_coordinates.size(0) != 1:
            raise RuntimeError('a [4, 4] output pose needs a batch of one scene-coordinate map')
        view = out_pose.unsqueeze(0)
        if not view.is_contiguous():
            tmp = torch.zeros(1, 4, 4, dtype=torch.float32, device=out_pose.device)
            forward_rgb_batch(scene_coordinates, tmp, ransac_hypotheses, inlier_threshold, focal_length, ppoint_x,
                              ppoint_y, inlier_alpha, max_reproj, sub_sampling)
            out_pose.copy_(tmp[0])
            return None
        forward_rgb_batch(scene_coordinates, view, ransac_hypotheses, inlier_threshold, focal_length, ppoint_x,
                          ppoint_y, inlier_alpha, max_reproj, sub_sampling)
        return None
    forward_rgb_batch(scene_coordinates, out_pose, ransac_hypotheses, inlier_threshold, focal_length, ppoint_x,
                      ppoint_y, inlier_alpha, max_reproj, sub_sampling)
    return None


def backward_rgb_batch(coords, grad, gt_pose, hyps, thr, focal, cx, cy, w_rot, w_trans, soft_clamp, alpha, max_reproj,
                       subsample, seed=None, image_base=None, max_tries=MAX_HYPOTHESES_TRIES, forced_samples=None,
                       debug=False):
    """Expected pose loss and its gradient for a batch.  coords / grad [B,3,Hc,Wc] float32 on one device (CPU or
    CUDA), grad is accumulated into; gt_pose [B,4,4] float32.  Returns the [B] float64 expected losses (host tensor),
    and with debug=True also a dict of host tensors (probs, losses, hyps_rt, ref_rt, tries, cells)."""
    lib = _lib.load()
    if coords.dim() != 4 or coords.size(1) != 3:
        raise RuntimeError('scene coordinates must be [B, 3, H, W], got %s' % (tuple(coords.shape),))
    if coords.dtype != torch.float32 or grad.dtype != torch.float32:
        raise RuntimeError('scene coordinates and their gradient must be float32')
    if grad.shape != coords.shape or not grad.is_contiguous() or grad.device != coords.device:
        raise RuntimeError('the gradient tensor must be contiguous, of the shape and on the device of the scene coordinates')
    B, _, Hc, Wc = coords.shape
    coords_c = coords.contiguous()
    gt = torch.as_tensor(gt_pose).to(torch.float32).reshape(B, 16).contiguous().cpu()
    if torch.is_tensor(focal):
        focal_t = focal.to(torch.float32).reshape(-1).contiguous().cpu()
        if focal_t.numel() == 1 and B > 1:
            focal_t = focal_t.expand(B).contiguous()
    else:
        focal_t = torch.full((B,), float(focal), dtype=torch.float32)
    if focal_t.numel() != B:
        raise RuntimeError('focal must be a scalar or hold one value per image')
    if seed is None:
        seed = _state['seed']
    if image_base is None:
        image_base = _state['image_index']
        _state['image_index'] += B
    forced_t = None
    if forced_samples is not None:
        forced_t = torch.as_tensor(forced_samples, dtype=torch.int32).reshape(B, hyps, 4, 2).contiguous()
    loss = torch.zeros(B, dtype=torch.float64)
    dbg = None
    if debug:
        dbg = {
            'probs': torch.zeros(B, hyps, dtype=torch.float64),
            'losses': torch.zeros(B, hyps, dtype=torch.float64),
            'hyps_rt': torch.zeros(B, hyps, 6, dtype=torch.float64),
            'ref_rt': torch.zeros(B, hyps, 6, dtype=torch.float64),
            'tries': torch.zeros(B, hyps, dtype=torch.int32),
            'cells': torch.zeros(B, hyps, 4, 2, dtype=torch.int32),
        }
    if coords_c.is_cuda:
        torch.cuda.set_device(coords_c.device)
    code = lib.cl_dsac_backward_rgb(
        _ptr(coords_c), B, Hc, Wc, _ptr(grad), _ptr(gt), int(hyps), float(thr), _ptr(focal_t), float(cx), float(cy),
        float(w_rot), float(w_trans), float(soft_clamp), float(alpha), float(max_reproj), int(subsample),
        int(seed) & 0xFFFFFFFFFFFFFFFF, int(image_base) & 0xFFFFFFFF, int(max_tries), _ptr(forced_t), _ptr(loss),
        _ptr(dbg['probs']) if dbg else None, _ptr(dbg['losses']) if dbg else None,
        _ptr(dbg['hyps_rt']) if dbg else None, _ptr(dbg['ref_rt']) if dbg else None,
        _ptr(dbg['tries']) if dbg else None, _ptr(dbg['cells']) if dbg else None, _stream_for(coords_c))
    _lib.check(code)
    return (loss, dbg) if debug else loss


def backward_rgb(scene_coordinates, out_scene_coordinates_grad, gt_pose, ransac_hypotheses, inlier_threshold,
                 focal_length, ppoint_x, ppoint_y, w_loss_rot, w_loss_trans, soft_clamp, inlier_alpha, max_reproj,
                 sub_sampling, random_seed):
    """`dsacstar.backward_rgb` (/root/reference/dsacstar/dsacstar.cpp:200-215, :889): [1,3,Hc,Wc] map, gradient tensor
    of the same shape (accumulated into) and [4,4] ground-truth pose -> expected pose loss (Python float).

    `random_seed` re-keys the sampler for this call as the reference's ThreadRand::init(randomSeed) does (:217).
    Also accepts CUDA tensors and batches (then returns a [B] float64 tensor).  Never prints."""
    gt = torch.as_tensor(gt_pose)
    B = scene_coordinates.size(0)
    loss = backward_rgb_batch(scene_coordinates, out_scene_coordinates_grad, gt.reshape(B, 4, 4), ransac_hypotheses,
                              inlier_threshold, focal_length, ppoint_x, ppoint_y, w_loss_rot, w_loss_trans, soft_clamp,
                              inlier_alpha, max_reproj, sub_sampling, seed=int(random_seed), image_base=0)
    return float(loss[0]) if B == 1 else loss
